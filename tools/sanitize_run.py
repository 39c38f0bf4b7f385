import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from gap2seq_b200 import engine, synth
import numpy as np
w = synth.make_workload(sys.argv[1], float(sys.argv[2]))
e = engine.Engine(k=w.k, max_fuz=w.fuz)
e.build_from_reads(w.reads, w.solid)
filled, log, gc, fc, res = e.fill_records([g.seq for g in w.gaps], [g.name for g in w.gaps], dist_error=w.dist_error, fuz=w.fuz)
t = e.timing()
print(sys.argv[1], 'gaps', gc, 'filled', fc, 'launches', t.launches, 'wide gaps', t.gaps_big_tier, 'cluster', t.wide_cluster)
sets = [w.reads_near([rec], reach=150) for rec in w.gaps[:8]]
e.build_graphs_from_reads(sets, w.solid)
print('own graphs', [e.graph_num_kmers(g) for g in range(len(sets))][:4])
e.close()
