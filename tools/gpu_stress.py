"""Repeats the device pass of a workload many times under a watchdog (a hung pass prints the Python stack and exits)."""
import faulthandler, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gap2seq_b200 import engine, synth
from gap2seq_b200.scaffold import record_gaps
name, scale, reps = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
w = synth.make_workload(name, scale)
eng = engine.Engine(k=w.k, max_fuz=w.fuz)
eng.build_from_reads(w.reads, w.solid)
specs = [s for rec in w.gaps for s in record_gaps(rec.seq, w.k, w.fuz, w.dist_error, all_paths=not w.best_only)]
eng.upload(specs)
faulthandler.dump_traceback_later(60, exit=True)
t0 = time.time()
ms = []
for i in range(reps):
    eng.run_resident()
    ms.append(eng.timing().kernel_ms)
    faulthandler.cancel_dump_traceback_later(); faulthandler.dump_traceback_later(60, exit=True)
res, fills = eng.finish_resident()
print(f"stress {name} x{scale}: {reps} passes ok, kernel ms min/max {min(ms):.2f}/{max(ms):.2f}, filled {sum(r.count > 0 for r in res)}/{len(res)}, {time.time() - t0:.1f} s")
