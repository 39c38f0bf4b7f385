"""The Gap2Seq-core compatible command line (gap2seq_b200/lib/Gap2Seq-core) against the golden outputs of the
reference binary: same flags, byte-identical filled FASTA, identical Scaffold:/SubgraphStats:/Filled lines."""
import os
import subprocess

import pytest

import cases
from test_emul_golden import single_gap_inputs

pytestmark = pytest.mark.gpu
CLI = os.path.join(cases.ROOT, "gap2seq_b200", "lib", "Gap2Seq-core")


def strip_nodes(golden_log: str) -> str:
    return "".join(ln + "\n" for ln in golden_log.splitlines() if not ln.startswith("Nodes:"))


@pytest.mark.parametrize("name", ["cfg1_small", "cfg1_multigap", "cfg2_small_unique", "cfg3_small_bestonly",
                                  "cfg1_allupper", "cfg1_fuz0", "cfg5_small"])
def test_cli_scaffold_mode(name, case_inputs, tmp_path):
    case, w, g = case_inputs(name)
    paths = w.write(str(tmp_path))
    out = str(tmp_path / "filled.fa")
    p = subprocess.run([CLI] + cases.ref_cmd_flags(case) + ["-reads", paths["reads"], "-scaffolds", paths["gaps"],
                                                            "-filled", out], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    gfa, glog = cases.golden_paths(name)
    assert open(out).read() == open(gfa).read()
    assert cases.filter_log(p.stdout) == strip_nodes(open(glog).read())
    assert "k-mer size: %d" % w.k in p.stdout and "All paths: %d" % (0 if case.opts()["best_only"] else 1) in p.stdout


def test_cli_single_gap_mode(case_inputs, tmp_path):
    case, w, g = case_inputs("cfg1_small")
    paths = w.write(str(tmp_path))
    gfa, glog = cases.golden_paths("single_gap")
    fa, log = "", ""
    for left, right, length in single_gap_inputs(w):
        out = str(tmp_path / "f.fa")
        p = subprocess.run([CLI] + cases.ref_cmd_flags(case) + ["-reads", paths["reads"], "-filled", out, "-left", left,
                                                                "-right", right, "-length", str(length)],
                           capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stdout + p.stderr
        fa += open(out).read()
        log += cases.filter_log(p.stdout)
    assert fa == open(gfa).read()
    assert log == open(glog).read()


def test_cli_rejects_unknown_flag_and_missing_reads():
    p = subprocess.run([CLI, "-bogus", "1"], capture_output=True, text=True)
    assert p.returncode != 0 and "Unknown parameter" in p.stdout
    p = subprocess.run([CLI, "-k", "31"], capture_output=True, text=True)
    assert p.returncode != 0
