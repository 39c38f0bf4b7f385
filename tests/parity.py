"""Shared comparison logic: a run of the engine's gap driver (real CUDA engine or its host emulation) against the
golden outputs of the unmodified reference (tests/golden/)."""
from __future__ import annotations

import cases
import oracle_lib


def golden_split(name):
    """golden log -> (per-gap [(nonzeros, scaffold line, subgraph line or None)], final line); golden FASTA -> records"""
    gfa, glog = cases.golden_paths(name)
    lines = open(glog).read().splitlines()
    gaps, cur = [], None
    for ln in lines:
        if ln.startswith("Nodes:"):
            cur = {"nonzeros": int(ln.split("Nonzeros:")[1]), "lines": []}
            gaps.append(cur)
        elif ln.startswith(("Scaffold:", "SubgraphStats:")):
            cur["lines"].append(ln)
    final = lines[-1]
    fa = open(gfa).read().splitlines()
    recs = [(fa[i][1:], fa[i + 1]) for i in range(0, len(fa), 2)]
    return gaps, final, recs


def check_against_golden(name, case, w, run_records, allow_order_dependent=True):
    """run_records(seqs, comments, **flags) -> (filled, log, gapcount, filledcount, results).
    Seeds libc rand() like Gap2Seq-core -randseed does (src/Gap2Seq.cpp:178) - the engine and the oracle library
    share the process-wide libc state."""
    o = case.opts()
    oracle_lib.srand(o["randseed"])
    seqs = [g.seq for g in w.gaps]
    comments = [g.name for g in w.gaps]
    filled, log, gc, fc, results = run_records(seqs, comments, dist_error=o["dist_error"], fuz=o["fuz"],
                                               skip_confident=o["all_upper"], unique=o["unique"],
                                               all_paths=not o["best_only"])
    ggaps, gfinal, grecs = golden_split(name)
    flagged = [r for r in results if r.status & 1]
    if flagged and not allow_order_dependent:
        raise AssertionError(f"{name}: unexpected ORDER_DEPENDENT gaps")
    assert len(results) == len(ggaps), (len(results), len(ggaps))
    # DP cell counts == the reference's -DSTATS Nonzeros, gap by gap
    for r, gg in zip(results, ggaps):
        if r.status & 1:
            continue
        assert r.cells_right + r.cells_left == gg["nonzeros"], (name, r.cells_right, r.cells_left, gg["nonzeros"])
    if not flagged:
        got_log = log.splitlines()
        want_log = [ln for gg in ggaps for ln in gg["lines"]]
        assert got_log == want_log
        assert f"Filled {fc} gaps out of {gc}" == gfinal
        assert [(c, f) for c, f in zip(comments, filled)] == grecs
    else:
        # order-dependent gaps are fenced: everything before the first flagged gap must still agree
        # (after it the rand() stream may have diverged)
        first = next(i for i, r in enumerate(results) if r.status & 1)
        got_log = log.splitlines()
        want = [ln for gg in ggaps[:first] for ln in gg["lines"]]
        assert got_log[:len(want)] == want
    return results
