#!/usr/bin/env python
"""Per-gap filtered-reads mode, measured: the graphs of all gaps of a workload built in one device
pass + all gaps filled in one pass (g2s_build_graphs_from_reads, g2s_fill_batch), beside the reference's way - one
Gap2Seq-core process and one GATB graph per gap (src/Gap2Seq.py:132-220) - timed on a few gaps with the unmodified
binary (oracle/_ref, when present).  Prints one JSON line.   usage: tools/own_graphs_bench.py [workload] [scale]"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gap2seq_b200 import engine                      # noqa: E402
from gap2seq_b200.scaffold import record_gaps        # noqa: E402
from gap2seq_b200.synth import make_workload, write_reads_fasta   # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    w = make_workload(name, scale)
    sets = [w.reads_near([rec], reach=150) for rec in w.gaps]
    specs = []
    for gi, rec in enumerate(w.gaps):
        (g,) = record_gaps(rec.seq, k=w.k, fuz=w.fuz, dist_error=w.dist_error)
        g.graph = gi
        specs.append(g)
    out = {"workload": name, "scale": scale, "gaps": len(specs), "reads_total": int(sum(s.shape[0] for s in sets))}
    with engine.Engine(k=w.k) as e:
        e.build_graphs_from_reads(sets, w.solid)      # warm-up (allocator, module load)
        e.fill_batch(specs)
        tb, tf = [], []
        for _ in range(5):
            t0 = time.perf_counter()
            e.build_graphs_from_reads(sets, w.solid)
            t1 = time.perf_counter()
            res, _ = e.fill_batch(specs)
            t2 = time.perf_counter()
            tb.append(t1 - t0); tf.append(t2 - t1)
        out["solid_kmers_total"] = int(sum(e.graph_num_kmers(g) for g in range(len(sets))))
    out["filled"] = sum(1 for r in res if r.count > 0)
    out["build_graphs_s"] = min(tb)
    out["fill_s"] = min(tf)
    out["gaps_per_s_build_and_fill"] = len(specs) / (min(tb) + min(tf))
    ref = os.path.join(ROOT, "oracle", "_ref", "bin", "Gap2Seq-core")
    if os.path.exists(ref):
        n_ref = min(8, len(specs))
        t_ref = 0.0
        with tempfile.TemporaryDirectory() as tmp:
            for gi in range(n_ref):
                rp = os.path.join(tmp, f"r{gi}.fa")
                write_reads_fasta(rp, sets[gi])
                wd = os.path.join(tmp, f"w{gi}")
                os.mkdir(wd)
                s = specs[gi]
                t0 = time.perf_counter()
                subprocess.run([ref, "-k", str(w.k), "-solid", str(w.solid), "-nb-cores", "1", "-randseed", "1", "-reads", rp, "-left", s.left,
                                "-right", s.right, "-length", str(s.gap_len), "-filled", os.path.join(wd, "f.fa")], cwd=wd, capture_output=True,
                               check=True, timeout=600)
                t_ref += time.perf_counter() - t0
        out["reference_one_process_per_gap"] = {"gaps_timed": n_ref, "s_per_gap": t_ref / n_ref, "gaps_per_s": n_ref / t_ref,
                                                "note": "unmodified Gap2Seq-core, single-gap mode, graph construction included (that is the mode's cost)"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
