"""Development aid: time the HOST stages of g2s_fill_records (scanner, analysis, traceback, splice, log) on a workload,
with the device pass replaced by the host emulation (whose own time is subtracted).  No GPU needed."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import emul_lib, oracle_lib
from gap2seq_b200 import synth
name, scale = sys.argv[1], float(sys.argv[2])
w = synth.make_workload(name, scale)
og = oracle_lib.OracleGraph.from_reads(w.k, w.solid, w.reads)
em = emul_lib.EmulEngine(w.k, og.kmers())
L = emul_lib.lib()
L.emul_times.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
seqs = [g.seq for g in w.gaps]; names = [g.name for g in w.gaps]
for it in range(3):
    a, b = C.c_double(), C.c_double()
    L.emul_times(a, b, 1)
    oracle_lib.srand(1)
    t0 = time.perf_counter()
    out = em.fill_records(seqs, names, dist_error=w.dist_error, fuz=w.fuz)
    t1 = time.perf_counter()
    L.emul_times(a, b, 1)
    print(f"total {1e3*(t1-t0):.2f} ms  emulated device pass {1e3*(a.value-b.value):.2f} ms  analyze {1e3*b.value:.2f} ms  "
          f"rest of host (scan, trace, splice, log, ctypes) {1e3*(t1-t0-a.value):.2f} ms")
