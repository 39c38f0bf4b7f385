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


GAP_FIELDS = ("count", "left_fuz", "right_fuz", "fill", "forward_count", "n_path_lengths", "rand_calls", "cells_right",
              "cells_left")


def compare_with_oracle(fill_batch, og, specs, seed=13, order_dependent_bit=1):
    """Every gap of `specs` through the engine (fill_batch(specs) -> (results, fills), libc rand() consumed in list
    order) against the oracle's fill_gap, gap by gap: count, fuz, fill string with case, forward count, path lengths,
    rand() calls, DP cells and - where Phase B ran - the six subgraph statistics.

    A gap the oracle flags as order dependent (the reference's own result depends on libstdc++ hashtable
    order) must carry the engine's ORDER_DEPENDENT bit; its result is not compared, and because the two rand() streams
    may have diverged there, BOTH sides are re-seeded and the comparison goes on with the gaps behind it - no gap is
    left untested.  Returns {"compared", "order_dependent", "passes"}."""
    start, compared, od, passes = 0, 0, [], 0
    while start < len(specs):
        part = specs[start:]
        oracle_lib.srand(seed + start)
        res, fills = fill_batch(part)
        passes += 1
        oracle_lib.srand(seed + start)
        nxt = len(specs)
        for j, (s, r, f) in enumerate(zip(part, res, fills)):
            o, of = og.fill_gap(s.left, s.right, s.gap_len, s.gap_err, s.left_max_fuz, s.right_max_fuz,
                                skip_confident=s.skip_confident, all_paths=s.all_paths)
            if o.order_dependent:
                assert r.status & order_dependent_bit, f"gap {start + j}: the oracle is order dependent, the engine did not flag it"
                od.append(start + j)
                nxt = start + j + 1
                break
            got = (r.count, r.left_fuz, r.right_fuz, f, r.forward_count, r.n_path_lengths, r.rand_calls, r.cells_right,
                   r.cells_left)
            want = (o.count, o.left_fuz, o.right_fuz, of, o.forward_count, o.n_path_lengths, o.rand_calls, o.cells_right,
                    o.cells_left)
            if got != want:
                diff = [n for n, a, b in zip(GAP_FIELDS, got, want) if a != b]
                raise AssertionError(f"gap {start + j} (len {s.gap_len}): {diff} differ: engine {got[:3] + got[4:]} oracle {want[:3] + want[4:]}")
            if o.substats_valid:
                assert r.substats.astuple() == tuple(o.substats), f"gap {start + j}: subgraph statistics differ"
            compared += 1
        start = nxt
    return {"compared": compared, "order_dependent": od, "passes": passes}
