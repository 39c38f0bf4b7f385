"""Every gap of a named workload through the CUDA engine (C ABI) against the CPU oracle: count, fuz, fill, cells, rand calls."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib
from gap2seq_b200 import engine, synth
from gap2seq_b200.scaffold import record_gaps
name, scale = sys.argv[1], float(sys.argv[2])
limit = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
w = synth.make_workload(name, scale)
eng = engine.Engine(k=w.k, max_fuz=w.fuz)
eng.build_from_reads(w.reads, w.solid)
keys, _ = eng.export_kmers()
og = oracle_lib.OracleGraph.from_kmers(w.k, keys)
specs = [s for rec in w.gaps for s in record_gaps(rec.seq, w.k, w.fuz, w.dist_error, all_paths=not w.best_only)][:limit]
for seed in (1, 9):
    oracle_lib.srand(seed)
    res, fills = eng.fill_batch(specs)
    oracle_lib.srand(seed)
    bad = skipped = 0
    t0 = time.time()
    for i, (s, r, f) in enumerate(zip(specs, res, fills)):
        o, of = og.fill_gap(s.left, s.right, s.gap_len, s.gap_err, s.left_max_fuz, s.right_max_fuz, all_paths=s.all_paths)
        if o.order_dependent:
            skipped = len(specs) - i
            break
        got = (r.count, r.left_fuz, r.right_fuz, f, r.forward_count, r.rand_calls, r.cells_right, r.cells_left)
        want = (o.count, o.left_fuz, o.right_fuz, of, o.forward_count, o.rand_calls, o.cells_right, o.cells_left)
        if got != want or (o.substats_valid and r.substats.astuple() != tuple(o.substats)):
            bad += 1
            if bad <= 3:
                print("MISMATCH gap", i, "len", s.gap_len, [a == b for a, b in zip(got, want)], got[:3], want[:3], got[4:], want[4:])
    print(f"check {name} x{scale} seed {seed}: {len(specs) - skipped} gaps compared, {bad} mismatches, {skipped} skipped after an order-dependent gap, oracle {time.time() - t0:.1f} s")
    assert bad == 0
