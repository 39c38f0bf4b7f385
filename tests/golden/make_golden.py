#!/usr/bin/env python
"""Generate the golden outputs under tests/golden/ by running the UNMODIFIED reference binary
(oracle/_ref/bin/Gap2Seq-core-stats, built from /root/reference by oracle/build_ref.sh) on the seeded
cases of tests/cases.py:  `-nb-cores 1 -randseed 1` is the only deterministic mode.

Runs only in the build container (the reference is not present on the GPU box); the outputs are
committed so that the tests can run anywhere.

    python tests/golden/make_golden.py [case ...]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cases  # noqa: E402

REF = os.path.join(cases.ROOT, "oracle", "_ref", "bin", "Gap2Seq-core-stats")
REF_PLAIN = os.path.join(cases.ROOT, "oracle", "_ref", "bin", "Gap2Seq-core")


def run_case(case: cases.Case) -> None:
    w = cases.build(case)
    tmp = tempfile.mkdtemp(prefix="g2s_golden_")
    try:
        paths = w.write(tmp)
        t0 = time.time()
        cmd = [REF] + cases.ref_cmd_flags(case) + ["-reads", paths["reads"], "-scaffolds", paths["gaps"], "-filled",
                                                   os.path.join(tmp, "filled.fa")]
        # one working directory per invocation: the reference writes ./.h5
        p = subprocess.run("ulimit -s unlimited; " + " ".join(cmd), shell=True, cwd=tmp, capture_output=True, text=True,
                           timeout=3600)
        if p.returncode != 0:
            raise RuntimeError(f"{case.name}: reference failed rc={p.returncode}\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}")
        # the -DSTATS build and the plain build must agree on the FASTA (sanity check of the golden itself)
        os.remove(os.path.join(tmp, ".h5")) if os.path.exists(os.path.join(tmp, ".h5")) else None
        cmd2 = [REF_PLAIN] + cases.ref_cmd_flags(case) + ["-reads", paths["reads"], "-scaffolds", paths["gaps"],
                                                         "-filled", os.path.join(tmp, "filled2.fa")]
        p2 = subprocess.run("ulimit -s unlimited; " + " ".join(cmd2), shell=True, cwd=tmp, capture_output=True,
                            text=True, timeout=3600)
        assert p2.returncode == 0
        a = open(os.path.join(tmp, "filled.fa")).read()
        b = open(os.path.join(tmp, "filled2.fa")).read()
        assert a == b, f"{case.name}: reference is not deterministic?"
        fa, log = cases.golden_paths(case.name)
        with open(fa, "w") as f:
            f.write(a)
        with open(log, "w") as f:
            f.write(cases.filter_log(p.stdout))
        print(f"{case.name}: {len(w.gaps)} records, {time.time() - t0:.1f} s, "
              f"{cases.filter_log(p.stdout).strip().splitlines()[-1]}")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_single_gap() -> None:
    """Single-gap mode (-left/-right/-length, src/Gap2Seq.cpp:227-283) with flanks LONGER than k+fuz: the right pass then
    starts from the end of the right flank while the targets sit at its beginning."""
    case = cases.CASE_BY_NAME["cfg1_small"]
    w = cases.build(case)
    tmp = tempfile.mkdtemp(prefix="g2s_golden_")
    try:
        paths = w.write(tmp)
        out_fa, out_log = [], []
        for gi, (fl, fr) in enumerate([(41, 41), (60, 55), (35, 48)]):
            rec = w.gaps[gi]
            i = rec.seq.index("N")
            left, right = rec.seq[i - fl:i], rec.seq[i + rec.gap_len:i + rec.gap_len + fr]
            cmd = [REF_PLAIN] + cases.ref_cmd_flags(case) + ["-reads", paths["reads"], "-filled", os.path.join(tmp, "f.fa"),
                                                           "-left", left, "-right", right, "-length", str(rec.gap_len)]
            p = subprocess.run(" ".join(cmd), shell=True, cwd=tmp, capture_output=True, text=True, timeout=600)
            assert p.returncode == 0, p.stdout[-2000:]
            out_fa.append(open(os.path.join(tmp, "f.fa")).read())
            out_log.append(cases.filter_log(p.stdout))
        fa, log = cases.golden_paths("single_gap")
        open(fa, "w").write("".join(out_fa))
        open(log, "w").write("".join(out_log))
        print("single_gap:", [len(x) for x in out_fa], out_log[1].strip()[:120])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_h5_fixture() -> None:
    """The graph file the reference writes for cfg1_small (./.h5 of the working directory), gzip'ed: the
    fixture of the HDF5 reader (gap2seq_b200/csrc/g2s_h5.h).  Its dsk/solid partitions hold the solid canonical k-mers."""
    import gzip
    case = cases.CASE_BY_NAME["cfg1_small"]
    w = cases.build(case)
    tmp = tempfile.mkdtemp(prefix="g2s_golden_")
    try:
        paths = w.write(tmp)
        cmd = [REF_PLAIN] + cases.ref_cmd_flags(case) + ["-reads", paths["reads"], "-scaffolds", paths["gaps"], "-filled",
                                                        os.path.join(tmp, "filled.fa")]
        p = subprocess.run("ulimit -s unlimited; " + " ".join(cmd), shell=True, cwd=tmp, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stdout[-2000:]
        raw = open(os.path.join(tmp, ".h5"), "rb").read()
        out = os.path.join(HERE, "cfg1_small.h5.gz")
        with open(out, "wb") as f:
            f.write(gzip.compress(raw, 9, mtime=0))
        print(f"h5: {len(raw)} bytes -> {os.path.getsize(out)} gzip'ed")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main() -> None:
    if not os.path.exists(REF):
        sys.exit("oracle/_ref is not built: run oracle/build_ref.sh in the container that has /root/reference")
    names = sys.argv[1:] or [c.name for c in cases.CASES] + ["single_gap", "h5"]
    for n in names:
        if n == "single_gap":
            run_single_gap()
        elif n == "h5":
            run_h5_fixture()
        else:
            run_case(cases.CASE_BY_NAME[n])


if __name__ == "__main__":
    main()
