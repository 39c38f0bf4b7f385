"""Gaps the reference skips for exceeding -max-mem are listed separately, not counted as mismatches (north_star; reference
src/Gap2Seq.cpp:140-147, 454-476, 1525-1526).  The unmodified reference binary is run here with a -max-mem so small that
its byte accounting gives up on some gaps of a repeat-rich workload; the engine's driver (host emulation of the device
pass) fills the same scaffolds without that limit; tools/compare_with_reference.py must list exactly the reference's
`Memory limit exceeded` gaps as skipped, call every gap in front of the first of them identical, and report no mismatch."""
import os
import subprocess
import sys

import pytest

import cases
import emul_lib
import oracle_lib

sys.path.insert(0, os.path.join(cases.ROOT, "tools"))
import compare_with_reference as cmp  # noqa: E402

REF = os.path.join(cases.ROOT, "oracle", "_ref", "bin", "Gap2Seq-core")


@pytest.mark.skipif(not os.path.exists(REF), reason="reference binary not built")
@pytest.mark.parametrize("name,max_mem_gb", [("cfg3_small", 0.0006), ("cfg3_small", 0.0008), ("cfg3_seed33", 0.0003)])
def test_reference_max_mem_skips_are_listed_not_mismatched(name, max_mem_gb, case_inputs, tmp_path):
    case, w, g = case_inputs(name)
    paths = w.write(str(tmp_path))
    out = str(tmp_path / "ref.fa")
    p = subprocess.run([REF] + cases.ref_cmd_flags(case) + ["-max-mem", str(max_mem_gb), "-reads", paths["reads"],
                                                            "-scaffolds", paths["gaps"], "-filled", out],
                       capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr[-2000:]
    ref_log = cases.filter_log(p.stdout)
    n_skipped = ref_log.count("Memory limit exceeded")
    assert n_skipped > 0, "the probe -max-mem did not make the reference skip any gap"
    o = case.opts()
    oracle_lib.srand(o["randseed"])
    em = emul_lib.EmulEngine(w.k, g.kmers())
    filled, log, gc, fc, res = em.fill_records([x.seq for x in w.gaps], [x.name for x in w.gaps], dist_error=o["dist_error"],
                                               fuz=o["fuz"], skip_confident=o["all_upper"], unique=o["unique"],
                                               all_paths=not o["best_only"])
    our_fa = "".join(f">{x.name}\n{s}\n" for x, s in zip(w.gaps, filled))
    r = cmp.compare(log, our_fa, ref_log, open(out).read())
    assert len(r["reference_skipped_max_mem"]) == n_skipped
    assert not r["mismatch"] and not r["records_mismatch"], r
    first = r["reference_skipped_max_mem"][0]["gap"]
    assert r["identical"] >= first               # everything in front of the first skipped gap is byte-identical
    assert r["identical"] + n_skipped + len(r["after_reference_skip"]) == r["gaps"]
