"""The named BASELINE.json workloads at (or near) FULL size through the CUDA engine (C ABI) against the CPU oracle,
gap by gap - the same comparison tools/gpu_check_oracle.py made by hand in round 1, now part of `-m gpu`:

  cfg2 x1.0   the 170 gaps of configs[1] (thin tier: chain splitting, merged chains, uncertain tracebacks), 3 modes
  cfg3 x1.0   the 60 repeat gaps of configs[2] (wide tier, general subgraphs, saturated counts), 3 modes
  cfg4 x0.3   3 000 gaps of configs[3] on a 30 Mbp genome (persistent queue deeper than the resident CTAs)
  cfg5 x0.5   2 500 insertion records of configs[4]

plus the device-memory budget: a forced small `max_mem_bytes` drives gaps through tiers 1 and 2, into a full pack
arena and past the last tier (count -1, "Memory limit exceeded").  Order-dependent gaps do not end a
comparison: both sides are re-seeded and the gaps behind them are still compared (parity.compare_with_oracle).
Bit-exact: integer / byte work only."""
import numpy as np
import pytest

import oracle_lib
import parity
from gap2seq_b200 import engine, synth
from gap2seq_b200.scaffold import record_gaps

pytestmark = pytest.mark.gpu

MODES = {"all_paths": dict(skip_confident=False, all_paths=True), "best_only": dict(skip_confident=False, all_paths=False),
         "all_upper": dict(skip_confident=True, all_paths=True)}


@pytest.fixture(scope="module")
def full_workload():
    """(workload, engine with the graph built on the device, oracle graph over the SAME k-mer set)."""
    cache = {}

    def get(name, scale):
        key = (name, scale)
        if key not in cache:
            for old in list(cache):            # one large graph at a time
                cache.pop(old)[1].close()
            w = synth.make_workload(name, scale)
            e = engine.Engine(k=w.k, max_fuz=w.fuz)
            e.build_from_reads(w.reads, w.solid)
            keys, _ = e.export_kmers()
            og = oracle_lib.OracleGraph.from_kmers(w.k, keys)
            w.reads = None
            cache[key] = (w, e, og)
        return cache[key]

    yield get
    for _, e, _ in cache.values():
        e.close()


def specs_of(w, **kw):
    return [s for rec in w.gaps for s in record_gaps(rec.seq, w.k, w.fuz, w.dist_error, **kw)]


@pytest.mark.parametrize("mode", list(MODES))
def test_cfg2_full_size_matches_oracle(mode, full_workload):
    """BASELINE configs[1] at full size: 2.9 Mbp, 50x reads with 1 % errors, 170 gaps."""
    w, e, og = full_workload("cfg2", 1.0)
    specs = specs_of(w, **MODES[mode])
    assert len(specs) == 170
    st = parity.compare_with_oracle(e.fill_batch, og, specs, seed=101)
    assert st["compared"] + len(st["order_dependent"]) == len(specs)
    assert not st["order_dependent"]          # random genome: no inverted repeats


@pytest.mark.parametrize("mode", list(MODES))
def test_cfg3_full_size_matches_oracle(mode, full_workload):
    """BASELINE configs[2] at full size: repeat-rich 5 Mbp genome, 60 gaps next to repeat copies; every gap outgrows the
    thin tier, most path subgraphs are general graphs (Tarjan + branch counter), counts saturate at MAX_PATHS."""
    w, e, og = full_workload("cfg3", 1.0)
    specs = specs_of(w, **MODES[mode])
    assert len(specs) == 60
    st = parity.compare_with_oracle(e.fill_batch, og, specs, seed=202)
    assert st["compared"] + len(st["order_dependent"]) == len(specs)
    res, _ = e.fill_batch(specs)
    assert sum(1 for r in res if r.status & engine.ST_BIG_TIER) >= 30
    if mode == "all_paths":
        assert any(r.status & engine.ST_GENERAL_SUBGRAPH for r in res)
        assert any(r.forward_count == 1073741822 for r in res), "no saturated count: the workload lost its repeats"


@pytest.mark.parametrize("cluster", [1, 4, 8])
def test_cfg3_every_wide_tier_shape_matches_oracle(cluster, full_workload, monkeypatch):
    """The wide tier runs a gap on one CTA (`k_fill_cta`) or on a thread-block cluster of 2 / 4 / 8 CTAs (`k_fill_cluster`);
    the engine picks the largest cluster that keeps the pass resident (2 for these 60 gaps, covered above).  The other
    shapes are forced here (G2S_CLUSTER is read when a pass is planned); with 4 or 8 CTAs per gap the 58 gaps no longer
    fit at once, so the work queue hands a cluster several gaps in a row."""
    monkeypatch.setenv("G2S_CLUSTER", str(cluster))
    w, e, og = full_workload("cfg3", 1.0)
    specs = specs_of(w, **MODES["all_paths"])
    st = parity.compare_with_oracle(e.fill_batch, og, specs, seed=404)
    assert st["compared"] + len(st["order_dependent"]) == len(specs)
    assert e.timing().wide_cluster == cluster


def test_cfg4_30pct_matches_oracle(full_workload):
    """BASELINE configs[3] at 0.3 of its size: 30 Mbp genome, 3 000 gaps of 100 bp - 5 kbp (three times the CTAs that are
    resident at once: the device work queue is exercised)."""
    w, e, og = full_workload("cfg4", 0.3)
    specs = specs_of(w)
    assert len(specs) == 3000
    st = parity.compare_with_oracle(e.fill_batch, og, specs, seed=303)
    assert st["compared"] == len(specs)
    # the driver call on the same records: byte-identical with the oracle's restatement of execute() for a sample
    seqs, names = [g.seq for g in w.gaps], [g.name for g in w.gaps]
    oracle_lib.srand(1)
    filled, log, gc, fc, _ = e.fill_records(seqs, names, dist_error=w.dist_error, fuz=w.fuz)
    assert gc == len(seqs)
    oracle_lib.srand(1)
    want_log = []
    for s, n, f in zip(seqs, names, filled):
        of, olog, _, _ = og.fill_record(s, n, d_err=w.dist_error, fuz=w.fuz)
        assert f == of
        want_log.append(olog)
    assert log == "".join(want_log)


def test_cfg5_half_size_matches_oracle(full_workload):
    """BASELINE configs[4] at half size: 10 Mbp reference, 2 500 insertions of 50-2000 bp as gap records."""
    w, e, og = full_workload("cfg5", 0.5)
    specs = specs_of(w)
    assert len(specs) >= 2400
    st = parity.compare_with_oracle(e.fill_batch, og, specs, seed=404)
    assert st["compared"] == len(specs)


@pytest.fixture(scope="module")
def heavy_repeats():
    """cfg3 with repeat_load 12 (interspersed copies cover ~20 % of a 500 kbp genome, the regime of the reference
    probe in SURVEY section 6: 10^6 - 10^7 DP cells per gap) plus 40 gaps placed anywhere on the same genome."""
    w = synth.make_workload("cfg3", 0.1, repeat_load=12.0)
    rng = np.random.default_rng(99)
    w.gaps = w.gaps + synth.place_gaps(w.genome, rng.integers(100, 701, size=40), rng, flank=100)
    og = oracle_lib.OracleGraph.from_reads(w.k, w.solid, w.reads)
    return w, og


def test_heavy_repeat_gaps_reach_tier_3_and_match_oracle(heavy_repeats):
    """Gaps of several million DP cells (tier 3 = 256 x the thin tier's slice) against the oracle."""
    w, og = heavy_repeats
    e = engine.Engine(k=w.k, max_fuz=w.fuz)
    e.import_kmers(og.kmers())
    specs = specs_of(w)
    res, fills = e.fill_batch(specs)
    t = e.timing()
    assert t.max_tier_run >= 3 and t.gaps_mem_limit == 0 and all(r.count >= 0 for r in res)
    cells = np.array([r.cells_left + r.cells_right for r in res])
    assert cells.max() > 5_000_000
    # the oracle needs ~10 s per 10^7 cells: compare the unique-sequence gaps and the three lightest heavy gaps
    order = np.argsort(cells)
    pick = sorted(int(i) for i in order[:len(order) - 3])
    st = parity.compare_with_oracle(e.fill_batch, og, [specs[i] for i in pick], seed=505)
    assert st["compared"] + len(st["order_dependent"]) == len(pick)
    e.close()


def test_small_memory_budget_walks_the_tiers_to_minus_one(heavy_repeats, monkeypatch):
    """-max-mem analogue (src/Gap2Seq.cpp:140-147, 1525-1526).  With a device budget that holds tier 2 but not tier 3,
    gaps are re-run tier after tier and those that outgrow tier 2 return -1 with G2S_ST_MEM_LIMIT and are printed as
    "Memory limit exceeded"; with a budget so small that the pack arena overflows (GS_OUT_FULL) the same happens one
    tier earlier.  Every gap that is NOT given up still equals the oracle."""
    w, og = heavy_repeats
    keys = og.kmers()
    specs = specs_of(w)
    full = engine.Engine(k=w.k, max_fuz=w.fuz)
    full.import_kmers(keys)
    ref_res, _ = full.fill_batch(specs)
    full.close()
    cells = np.array([r.cells_left for r in ref_res])
    # slice of tier t (Dmax 1220): 0.6 MB, 8.4 MB, 117 MB, 1.9 GB.  Third round: default budget, tiny pack arena
    for budget, want_tier, pack_cap in ((1 << 30, 2, None), (4 << 20, 0, None), (0, None, 20000)):
        if pack_cap:
            monkeypatch.setenv("G2S_PACK_CAP_RECORDS", str(pack_cap))
        small = engine.Engine(k=w.k, max_fuz=w.fuz, max_mem_bytes=budget)
        small.import_kmers(keys)
        res, fills = small.fill_batch(specs)
        t = small.timing()
        lost = [i for i, r in enumerate(res) if r.count == -1]
        kept = [i for i in range(len(specs)) if i not in lost]
        assert t.gaps_mem_limit == len(lost) and lost and kept, (budget, pack_cap, len(lost), len(kept))
        assert t.gaps_rerun_scratch + t.gaps_rerun_out_full >= len(lost)
        if pack_cap:
            assert t.gaps_rerun_out_full > 0, "the small pack arena never overflowed"
        else:
            assert t.max_tier_run == want_tier
        for i in lost:
            assert res[i].status & engine.ST_MEM_LIMIT and fills[i] == "" and res[i].fill_len == 0
        if not pack_cap:
            assert cells[lost].min() >= cells[kept].max() // 64   # the gaps given up are the big ones
        st = parity.compare_with_oracle(small.fill_batch, og, [specs[i] for i in kept if cells[i] < 2_000_000], seed=606)
        assert st["compared"] > 0
        # the driver prints the reference's line for such gaps and leaves their Ns in place
        seqs, names = [w.gaps[i].seq for i in lost[:2]], [w.gaps[i].name for i in lost[:2]]
        filled, log, gc, fc, _ = small.fill_records(seqs, names, dist_error=w.dist_error, fuz=w.fuz)
        assert fc == 0 and gc == len(seqs) and filled == seqs
        assert log.count("Memory limit exceeded") == len(seqs)
        small.close()
