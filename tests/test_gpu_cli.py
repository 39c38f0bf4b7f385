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


def test_whole_flow_cutter_core_merger_equals_the_reference_flow(case_inputs, tmp_path):
    """GapCutter -> Gap2Seq-core -> GapMerger, the flow the reference's wrapper drives (src/Gap2Seq.py), with this
    package's three binaries (the core on the GPU): the merged scaffolds restore the error-free genome and - when
    oracle/_ref is present - equal, file by file, what the reference's three binaries produce from the same inputs."""
    lib = os.path.join(cases.ROOT, "gap2seq_b200", "lib")
    ref = os.path.join(cases.ROOT, "oracle", "_ref", "bin")
    case, w, g = case_inputs("cfg1_small")
    genome = w.genome.tobytes().decode()
    seq = list(genome[:12000])
    for p, L in [(800, 120), (2500, 60), (2500 + 60 + 100, 40), (5000, 300), (7000, 90), (9000, 450), (11000, 33)]:
        seq[p:p + L] = "N" * L
    reads = w.write(str(tmp_path))["reads"]

    def flow(tool_dir, tag):
        d = tmp_path / tag
        d.mkdir()
        (d / "scaf.fa").write_text(">chr1 test\n" + "".join(seq) + "\n>chr2 no gaps\n" + genome[13000:13400] + "\n")
        run = lambda tool, args: subprocess.run([os.path.join(tool_dir, tool)] + args, cwd=str(d), capture_output=True, text=True,
                                                timeout=900, check=True).stdout
        run("GapCutter", ["-scaffolds", "scaf.fa", "-contigs", "c.fa", "-gaps", "g.fa", "-bed", "b.bed"])
        log = run("Gap2Seq-core", ["-k", "31", "-solid", "2", "-nb-cores", "1", "-randseed", "3", "-reads", reads, "-scaffolds", "g.fa",
                                   "-filled", "f.fa"])
        run("GapMerger", ["-scaffolds", "m.fa", "-contigs", "c.fa", "-gaps", "f.fa"])
        return {n: (d / n).read_text() for n in ("c.fa", "g.fa", "b.bed", "f.fa", "m.fa")}, cases.filter_log(log)

    ours, our_log = flow(lib, "b200")
    assert ours["m.fa"].upper() == ">CHR1 TEST\n" + genome[:12000] + "\n>CHR2 NO GAPS\n" + genome[13000:13400] + "\n"
    assert our_log.strip().endswith("Filled 7 gaps out of 7")
    if os.path.exists(os.path.join(ref, "Gap2Seq-core")) and os.path.exists(os.path.join(ref, "GapCutter")):
        theirs, their_log = flow(ref, "ref")
        assert ours == theirs
        assert our_log == their_log
