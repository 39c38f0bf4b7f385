"""Development aid: run the host emulation of the per-gap algorithm with -DG2S_EMUL_STATS on the longest gaps of a
workload and print node-table probe lengths and the level-width mix of the backward sweep (no GPU needed)."""
import ctypes as C, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import emul_lib, oracle_lib
from gap2seq_b200 import synth, engine as E

so = "/tmp/libg2s_emul_stats.so"
subprocess.run(["g++", "-O2", "-std=c++17", "-DG2S_HOST_EMUL", "-DG2S_EMUL_STATS", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                os.path.join(ROOT, "tests/emul/g2s_emul.cpp"), "-o", so, "-lpthread"], check=True)
emul_lib._SO = so
emul_lib._SRC = [so]
name, scale, top = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
t0 = time.time()
w = synth.make_workload(name, scale)
og = oracle_lib.OracleGraph.from_reads(w.k, w.solid, w.reads)
keys = og.kmers()
print("graph", keys.size, "kmers", time.time() - t0, "s")
em = emul_lib.EmulEngine(w.k, keys)
L = emul_lib.lib()
L.emul_stats_get.argtypes = [C.c_void_p, C.c_int]
gaps = sorted(w.gaps, key=lambda g: -g.gap_len)[:top]
buf = (C.c_ulonglong * 15)()
for g in gaps:
    L.emul_stats_get(buf, 1)
    out = em.fill_records([g.seq], [g.name], dist_error=w.dist_error, fuz=w.fuz)
    L.emul_stats_get(buf, 1)
    v = list(buf)
    print(f"gap {g.gap_len}: finds {v[0]} probes/find {v[1]/max(v[0],1):.2f}  inserts {v[2]} probes/insert {v[3]/max(v[2],1):.2f}  "
          f"width_hist(back general) {v[5:13]} general_levels {v[13]} chain_levels {v[14]}")
