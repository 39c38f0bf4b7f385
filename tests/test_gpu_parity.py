"""Parity of the CUDA engine (through the C ABI, gap2seq_b200/lib/libg2s_gpu.so) with the reference:
  * golden outputs of the unmodified reference binary (tests/golden/, byte-identical FASTA + log lines + DP cell counts)
  * the CPU oracle gap by gap (count, fuz, fill string, subgraph statistics, cells)
  * size-independent properties at larger sizes (error-free random genome: every fill equals the genome).
Bit-exact: integer / byte work only."""
import ctypes as C

import numpy as np
import pytest

import cases
import oracle_lib
import parity
from gap2seq_b200 import engine, synth
from gap2seq_b200.scaffold import record_gaps

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu_engine(case_inputs):
    cache = {}

    def get(name):
        if name not in cache:
            case, w, g = case_inputs(name)
            e = engine.Engine(k=w.k, max_fuz=max(w.fuz, 1))
            e.build_from_reads(w.reads, w.solid)
            cache[name] = e
        return cache[name]

    yield get
    for e in cache.values():
        e.close()


@pytest.mark.parametrize("name", ["cfg1_small", "cfg2_small", "cfg3_small", "cfg1_k21"])
def test_device_table_equals_oracle_graph(name, case_inputs, gpu_engine):
    case, w, g = case_inputs(name)
    keys, adj = gpu_engine(name).export_kmers()
    okeys, oadj = g.kmers(with_adjacency=True)
    assert np.array_equal(keys, okeys)
    assert np.array_equal(adj, oadj)


@pytest.mark.parametrize("name,limit", [("cfg2_small", 60000), ("cfg2_small_solid3", 20000), ("cfg1_k21", 1024)])
def test_partitioned_counting_equals_oracle_graph(name, limit, case_inputs, monkeypatch):
    """Inputs with more k-mer occurrences than one sort pass holds (2^30; cub's run-length API counts in int) are counted in
    hash partitions of the canonical k-mer.  A tiny limit forces 2 - 256 partitions on a small read set: same solid k-mers,
    same adjacency bytes, same hints-derived fills as the one-pass path."""
    case, w, g = case_inputs(name)
    monkeypatch.setenv("G2S_COUNT_PART_LIMIT", str(limit))
    e = engine.Engine(k=w.k, max_fuz=max(w.fuz, 1))
    e.build_from_reads(w.reads, w.solid)
    keys, adj = e.export_kmers()
    okeys, oadj = g.kmers(with_adjacency=True)
    assert np.array_equal(keys, okeys)
    assert np.array_equal(adj, oadj)
    e.close()


def test_import_kmers_equals_build_from_reads(case_inputs, gpu_engine):
    case, w, g = case_inputs("cfg2_small")
    e = engine.Engine(k=w.k)
    e.import_kmers(g.kmers())
    k1, a1 = e.export_kmers()
    k2, a2 = gpu_engine("cfg2_small").export_kmers()
    assert np.array_equal(k1, k2) and np.array_equal(a1, a2)
    rng = np.random.default_rng(0)
    probe = ["".join(rng.choice(list("ACGT"), size=w.k)) for _ in range(64)]
    probe += [w.genome[i:i + w.k].tobytes().decode() for i in range(0, 6400, 100)]
    got = e.contains(probe)
    assert got[64:].mean() > 0.9 and got[:64].sum() == 0
    e.close()


@pytest.mark.parametrize("name", [c.name for c in cases.CASES])
def test_engine_matches_reference_golden(name, case_inputs, gpu_engine):
    case, w, g = case_inputs(name)
    parity.check_against_golden(name, case, w, gpu_engine(name).fill_records)


@pytest.mark.parametrize("name", ["cfg1_full", "cfg2_small", "cfg3_small", "cfg3_seed33", "cfg5_small"])
@pytest.mark.parametrize("mode", ["all_paths", "best_only", "all_upper"])
def test_fill_batch_matches_oracle_gap_by_gap(name, mode, case_inputs, gpu_engine):
    case, w, g = case_inputs(name)
    eng = gpu_engine(name)
    kw = dict(skip_confident=(mode == "all_upper"), all_paths=(mode != "best_only"))
    specs = []
    for rec in w.gaps:
        specs += record_gaps(rec.seq, w.k, w.fuz, 500, **kw)
    st = parity.compare_with_oracle(eng.fill_batch, g, specs, seed=7)
    assert st["compared"] + len(st["order_dependent"]) == len(specs)


def test_order_dependent_gap_is_flagged_and_later_gaps_still_compared():
    """SURVEY Q3 fence, driven positively: a constructed gap whose border receives both strands of one k-mer on one
    level must carry G2S_ST_ORDER_DEPENDENT; the gaps before AND behind it equal the oracle."""
    w = synth.make_order_dependent_workload()
    e = engine.Engine(k=w.k, max_fuz=w.fuz)
    e.build_from_reads(w.reads, w.solid)
    keys, _ = e.export_kmers()
    og = oracle_lib.OracleGraph.from_kmers(w.k, keys)
    specs = [s for rec in w.gaps for s in record_gaps(rec.seq, w.k, w.fuz, w.dist_error)]
    st = parity.compare_with_oracle(e.fill_batch, og, specs, seed=17)
    assert st["order_dependent"] == [3] and st["compared"] == len(specs) - 1 and st["passes"] == 2
    # the driver call reports the flag in the per-gap results
    _, _, gc, _, results = e.fill_records([g.seq for g in w.gaps], [g.name for g in w.gaps], dist_error=w.dist_error, fuz=w.fuz)
    assert gc == len(specs) and [i for i, r in enumerate(results) if r.status & engine.ST_ORDER_DEPENDENT] == [3]
    e.close()


def test_single_fill_gap_dropin_and_resident_split(case_inputs, gpu_engine):
    case, w, g = case_inputs("cfg1_small")
    eng = gpu_engine("cfg1_small")
    specs = [s for rec in w.gaps for s in record_gaps(rec.seq, w.k, w.fuz, 500)]
    oracle_lib.srand(3)
    singles = [eng.fill_gap(s) for s in specs]
    oracle_lib.srand(3)
    eng.upload(specs)
    eng.run_resident()
    eng.run_resident()   # the device pass is repeatable
    res, fills = eng.finish_resident()
    for s, (cnt, lf, rf, fill, ss), r, f, rec in zip(specs, singles, res, fills, w.gaps):
        assert (cnt, lf, rf, fill) == (r.count, r.left_fuz, r.right_fuz, f)
        assert ss == r.substats.astuple()
        assert f[:-w.k] == rec.truth
    t = eng.timing()
    assert t.launches >= 1 and t.kernel_ms > 0


def test_wide_tier_is_used_and_agrees(case_inputs, gpu_engine):
    """cfg3 gaps outgrow the warp tier's slice (4*(Dmax+2)+1024 cells) and are re-run one CTA per gap."""
    case, w, g = case_inputs("cfg3_small")
    eng = gpu_engine("cfg3_small")
    specs = [s for rec in w.gaps for s in record_gaps(rec.seq, w.k, w.fuz, 500)]
    res, fills = eng.fill_batch(specs)
    assert any(r.status & engine.ST_BIG_TIER for r in res)
    assert eng.timing().gaps_big_tier >= 1


def test_errors_are_loud():
    e = engine.Engine(k=31)
    with pytest.raises(engine.G2SError):
        e.fill_batch([engine.GapSpec("A" * 41, "C" * 41, 10)])   # no graph imported
    e.import_kmers(np.zeros(0, dtype=np.uint64))
    res, fills = e.fill_batch([engine.GapSpec("A" * 41, "C" * 41, 10)])
    assert res[0].count == 0 and fills[0] == ""
    assert e.fill_batch([]) == ([], [])
    with pytest.raises(engine.G2SError):
        e.fill_batch([engine.GapSpec("A" * 71, "C" * 41, 10, left_max_fuz=40)])   # fuz beyond the configured bound
    e.close()


def test_property_error_free_genome_fills_equal_truth():
    """BASELINE configs[3] shape at reduced size (2 Mbp, 400 gaps up to 5 kbp): on an error-free random genome every
    gap has exactly one path, so fill == the removed bases, all upper case, count == 1, and the DP cell count has the
    closed form of SURVEY 3.2."""
    w = synth.make_workload("cfg4", 0.02, seed=44)
    eng = engine.Engine(k=w.k, max_fuz=w.fuz)
    eng.build_from_reads(w.reads, w.solid)
    specs = [s for rec in w.gaps for s in record_gaps(rec.seq, w.k, w.fuz, 500)]
    res, fills = eng.fill_batch(specs)
    n_ok = 0
    for s, r, f, rec in zip(specs, res, fills, w.gaps):
        if r.count != 1:
            continue   # a coverage hole (30x leaves a few) - must then be unfilled in the oracle too; checked below
        n_ok += 1
        assert f[:-w.k] == rec.truth and f.isupper()
        gl = s.gap_len
        assert r.cells_right == 10 + -(-(gl + 500) // 2) + 1
        assert r.cells_left == max(gl // 2 + 250 + 10, 10 + gl + 31 + 10 + 1)
    assert n_ok >= 0.9 * len(specs)
    # spot-check the unfilled / multi-path ones against the oracle
    keys, _ = eng.export_kmers()
    og = oracle_lib.OracleGraph.from_kmers(w.k, keys)
    odd = [i for i, r in enumerate(res) if r.count != 1][:20]
    oracle_lib.srand(5)
    r2, f2 = eng.fill_batch([specs[i] for i in odd])
    oracle_lib.srand(5)
    for i, r, f in zip(odd, r2, f2):
        s = specs[i]
        o, of = og.fill_gap(s.left, s.right, s.gap_len, s.gap_err, s.left_max_fuz, s.right_max_fuz)
        assert (r.count, f) == (o.count, of)
    eng.close()


def test_table_replication_roundtrip(case_inputs, gpu_engine):
    """g2s_table_raw / g2s_table_adopt: the device-to-device replication path bench.py uses under torchrun."""
    import torch
    from gap2seq_b200.shard import DeviceBytes
    case, w, g = case_inputs("cfg1_small")
    src = gpu_engine("cfg1_small")
    ptr, cap, n, nbytes = src.table_raw()
    dst = engine.Engine(k=w.k, max_fuz=w.fuz)
    dptr = dst.table_adopt(cap, n)
    torch.as_tensor(DeviceBytes(dptr, nbytes), device="cuda").copy_(torch.as_tensor(DeviceBytes(ptr, nbytes), device="cuda"))
    torch.cuda.synchronize()
    k1, a1 = src.export_kmers()
    k2, a2 = dst.export_kmers()
    assert np.array_equal(k1, k2) and np.array_equal(a1, a2)
    specs = [s for rec in w.gaps for s in record_gaps(rec.seq, w.k, w.fuz, 500)]
    oracle_lib.srand(1)
    r1, f1 = src.fill_batch(specs)
    oracle_lib.srand(1)
    r2, f2 = dst.fill_batch(specs)
    assert f1 == f2 and all(f[:-w.k] == rec.truth for f, rec in zip(f1, w.gaps))
    dst.close()


@pytest.mark.parametrize("shape", [0, 1, 2])
def test_every_thin_cta_shape_matches_oracle(shape, case_inputs, monkeypatch):
    """The thin tier runs a gap on 2, 6 or 12 warps (picked per pass from the batch size); G2S_THIN_SHAPE forces one.
    cfg2_small has read errors (bubbles: chain splitting, merged chains, uncertain tracebacks)."""
    monkeypatch.setenv("G2S_THIN_SHAPE", str(shape))
    for name in ["cfg2_small", "cfg1_small"]:
        case, w, g = case_inputs(name)
        e = engine.Engine(k=w.k, max_fuz=w.fuz)
        e.import_kmers(g.kmers())
        specs = [s for rec in w.gaps for s in record_gaps(rec.seq, w.k, w.fuz, 500)]
        oracle_lib.srand(21)
        res, fills = e.fill_batch(specs)
        oracle_lib.srand(21)
        for s, r, f in zip(specs, res, fills):
            o, of = g.fill_gap(s.left, s.right, s.gap_len, s.gap_err, s.left_max_fuz, s.right_max_fuz)
            assert not o.order_dependent
            assert (r.count, r.left_fuz, r.right_fuz, f, r.rand_calls) == (o.count, o.left_fuz, o.right_fuz, of, o.rand_calls)
            assert (r.cells_right, r.cells_left) == (o.cells_right, o.cells_left)
        e.close()


def test_device_traceback_equals_host_traceback_and_rand_stream(case_inputs, monkeypatch):
    """Gaps whose trellis is a single chain are traced by the kernel (G2S_ST_DEVICE_TRACED): same fills, counts and
    rand() consumption as the host walk over pack records, and libc's generator ends in the same state."""
    libc = C.CDLL(None)
    w = synth.make_workload("cfg4", 0.01, seed=45)
    specs = [s for rec in w.gaps for s in record_gaps(rec.seq, w.k, w.fuz, 500)]
    case, w2, g2 = case_inputs("cfg2_small")
    specs2 = [s for rec in w2.gaps for s in record_gaps(rec.seq, w2.k, w2.fuz, 500)]
    outs = []
    for host_only in (False, True):
        if host_only:
            monkeypatch.setenv("G2S_NO_DEVICE_TRACE", "1")
        row = []
        for ww, sp in ((w, specs), (w2, specs2)):
            e = engine.Engine(k=ww.k, max_fuz=ww.fuz)
            e.build_from_reads(ww.reads, ww.solid)
            oracle_lib.srand(5)
            res, fills = e.fill_batch(sp)
            after = [libc.rand() for _ in range(4)]
            row.append(([(r.count, r.left_fuz, r.right_fuz, r.rand_calls, r.fill_len, r.substats.astuple()) for r in res], fills,
                        after, sum(1 for r in res if r.status & engine.ST_DEVICE_TRACED)))
            e.close()
        outs.append(row)
    for dev, host in zip(outs[0], outs[1]):
        assert dev[0] == host[0] and dev[1] == host[1] and dev[2] == host[2]
        assert host[3] == 0
    assert outs[0][0][3] >= 0.9 * len(specs), "error-free gaps must be traced on the device"
    assert all(f[:-w.k] == rec.truth for f, rec, r in zip(outs[0][0][1], w.gaps, outs[0][0][0]) if r[0] == 1)


@pytest.mark.parametrize("name,scale", [("cfg2", 0.3), ("cfg3", 0.1), ("cfg5", 0.05)])
def test_named_workloads_match_oracle_gap_by_gap(name, scale):
    """Every gap of a named workload at moderate size (50-250 gaps) through the engine against the oracle: thin tier with
    chain splitting and device tracebacks (cfg2, cfg5), wide tier with general subgraphs (cfg3)."""
    w = synth.make_workload(name, scale)
    e = engine.Engine(k=w.k, max_fuz=w.fuz)
    e.build_from_reads(w.reads, w.solid)
    keys, _ = e.export_kmers()
    og = oracle_lib.OracleGraph.from_kmers(w.k, keys)
    specs = [s for rec in w.gaps for s in record_gaps(rec.seq, w.k, w.fuz, w.dist_error, all_paths=not w.best_only)]
    st = parity.compare_with_oracle(e.fill_batch, og, specs, seed=13)
    assert st["compared"] + len(st["order_dependent"]) == len(specs)
    e.close()
