"""SURVEY Q3 (same-level strand collision) on the CPU: the oracle's detector and the engine's detector (host emulation of
the per-gap device algorithm) agree on a constructed hairpin gap, and a comparison does not stop at such a gap."""
import emul_lib
import oracle_lib
import parity
from gap2seq_b200 import synth
from gap2seq_b200.scaffold import record_gaps


def test_constructed_hairpin_gap_is_order_dependent_in_oracle_and_engine():
    w = synth.make_order_dependent_workload()
    og = oracle_lib.OracleGraph.from_reads(w.k, w.solid, w.reads)
    em = emul_lib.EmulEngine(w.k, og.kmers())
    for mode in (dict(), dict(all_paths=False), dict(skip_confident=True)):
        specs = [s for rec in w.gaps for s in record_gaps(rec.seq, w.k, w.fuz, w.dist_error, **mode)]
        flags = [og.fill_gap(s.left, s.right, s.gap_len, s.gap_err, s.left_max_fuz, s.right_max_fuz, **mode)[0].order_dependent
                 for s in specs]
        assert flags == [0, 0, 0, 1, 0, 0, 0]
        st = parity.compare_with_oracle(em.fill_batch, og, specs, seed=3)
        assert st == {"compared": 6, "order_dependent": [3], "passes": 2}


def test_plain_gaps_of_that_workload_fill_to_the_genome():
    w = synth.make_order_dependent_workload()
    og = oracle_lib.OracleGraph.from_reads(w.k, w.solid, w.reads)
    em = emul_lib.EmulEngine(w.k, og.kmers())
    specs = [s for rec in w.gaps for s in record_gaps(rec.seq, w.k, w.fuz, w.dist_error)]
    res, fills = em.fill_batch(specs)
    for i, (r, f, rec) in enumerate(zip(res, fills, w.gaps)):
        if i != 3:
            assert r.count == 1 and f[:-w.k] == rec.truth
