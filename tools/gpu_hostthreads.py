"""d2h / host time of g2s_fill_records as a function of g2s_config.host_threads (cfg2 by default)."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gap2seq_b200 import engine, synth
name, scale = (sys.argv[1], float(sys.argv[2])) if len(sys.argv) > 2 else ("cfg2", 1.0)
w = synth.make_workload(name, scale)
libc = ctypes.CDLL(None)
seqs = [g.seq for g in w.gaps]; names = [g.name for g in w.gaps]
for th in (1, 2, 4, 8, 16):
    eng = engine.Engine(k=w.k, max_fuz=w.fuz, host_threads=th)
    eng.build_from_reads(w.reads, w.solid)
    prep = eng.prepare_records(seqs, names)
    rows = []
    for it in range(8):
        libc.srand(1)
        t0 = time.perf_counter()
        out = eng.fill_records_c(prep, dist_error=w.dist_error, fuz=w.fuz, all_paths=not w.best_only)
        wall = 1e3 * (time.perf_counter() - t0)
        out[4]()
        t = eng.timing()
        rows.append((wall, t.kernel_ms, t.d2h_ms, t.host_ms))
    best = min(rows[2:])
    print(f"host_threads={th:2d}: wall {best[0]:.3f} ms kernel {best[1]:.3f} d2h {best[2]:.3f} host {best[3]:.3f}  (all walls: {' '.join('%.2f' % r[0] for r in rows)})")
    eng.close()
