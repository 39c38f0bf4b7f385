#!/usr/bin/env python
"""bench.py - throughput of the gap-filling hot path (Gap2Seq::fill_gap, reference src/Gap2Seq.cpp:858-1556)
on the BASELINE.json workload the metric is quoted on: `configs[3]`, the chromosome-scale synthetic genome with
10 000 gaps of 100 bp - 5 kbp ("gaps sharded across 1/2/4/8 B200"); it fits one GPU (8.6 GB k-mer table).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg4] [--scale 1.0]

One "step" = one pass of the hot path over the workload's whole gap list.
  value        gaps/s of the device pass with the gaps already resident in HBM (CUDA events on the engine's stream).
               N > 1 (torchrun, one process per GPU): the ONE gap list is split longest-processing-time-first over the
               ranks (strong scaling), the k-mer table is built on rank 0 and replicated by one NCCL broadcast; the
               step time is the max over ranks
  e2e          gaps/s through the reference-facing driver call (g2s_fill_records: host strings in, filled strings
               out; H2D, kernels, D2H, tracebacks and splice inside the timed region).  N > 1: rank 0 makes that call on
               an engine that owns all N GPUs (g2s_create_multi: sharded waves, results gathered in input order, one
               rand() stream) while the other ranks wait; its output must equal the one-GPU output byte for byte
  roofline     the kernel with the largest share of the step against the HBM-gather roofline: 52 algorithmic bytes per
               DP cell; kernel name = the tier that actually finished the gaps
  cpu_baseline the oracle port on a bounded prefix of the same gaps (1 host core)
The step's output is compared with the oracle (a prefix of the records, outside the timed regions): a mismatch is fatal.
`--impl reference` times the unmodified reference binary (oracle/_ref/bin/Gap2Seq-core) on the host CPU at
-nb-cores 1, 4 and all cores (the timed steps run at the fastest of the three) on a bounded sample of the workload.
Graph construction is not part of any timed region (BASELINE.json: "timed separately and is not the product"); its
duration is reported under config.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PHASES = ["init", "right", "forward", "backward", "pack", "total", "fwd_scout", "fwd_node", "fwd_layout", "fwd_found",
          "fwd_general", "n_narrow_batches", "n_narrow_levels", "n_general_levels"]
BYTES_PER_CELL = 52          # 8 B frontier read + 32 B sector gather + 4 B count update + 8 B cell write
WORKLOAD_DESC = {
    "cfg1": "cfg1: 100 kbp random genome, 30x error-free 100 bp reads, 20 gaps 50-500 bp, k=31, dist-error 500",
    "cfg2": "cfg2: 2.9 Mbp random genome, 50x 101 bp reads at 1% substitution error, 170 gaps 50-3000 bp, k=31, dist-error 500",
    "cfg3": "cfg3: repeat-rich 5 Mbp genome (tandem + interspersed repeats), 60 gaps, all-paths, k=31",
    "cfg4": "cfg4: 100 Mbp random genome, 30x reads, 10k gaps 100 bp-5 kbp, k=31",
    "cfg5": "cfg5: 20 Mbp reference with 5k insertions 50-2000 bp as gap records, 40x reads, k=31",
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--oracle-seconds", type=float, default=8.0, help="budget of the oracle check of the step's output")
    ap.add_argument("--ref-gaps", type=int, default=1000, help="--impl reference: gaps in the bounded sample")
    return ap.parse_args()


def measured_traffic(workload: str, scale: float):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this workload
    (profiles/ncu_traffic.json, written from the capture named there), or None when there is no capture for it."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        ent = json.load(open(p)).get(f"{workload}@{scale:g}")
        return (float(ent["dram_bytes_per_launch"]), ent.get("source")) if ent else (None, None)
    except Exception:
        return None, None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks line of B200_PROFILING.md, sampled during the timed region."""

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(name: str, scale: float, rank: int):
    from gap2seq_b200 import synth
    w = synth.make_workload(name, scale, with_reads=(rank == 0))   # only rank 0 builds the graph (then broadcast)
    if rank > 0:   # weak scaling: every rank fills its own, equally shaped batch of gaps on the same graph
        import numpy as np
        rng = np.random.default_rng(7000 + rank)
        lens = [g.gap_len for g in w.gaps]
        w.gaps = synth.place_gaps(w.genome, np.array(lens), rng, flank=100)
    return w


def run_reference(args, rank: int, world: int):
    """The reference's own CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    w = make_workload(args.workload, args.scale, 0)
    binp = os.path.join(ROOT, "oracle", "_ref", "bin", "Gap2Seq-core")
    cores = os.cpu_count() or 1
    n_gaps = len(w.gaps)
    line = {"impl": "reference", "metric": "gaps_per_s", "unit": "gaps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": {"workload": WORKLOAD_DESC.get(args.workload, args.workload), "scale": args.scale}}
    if os.path.exists(binp):
        tmp = tempfile.mkdtemp(prefix="g2s_refbench_")
        try:
            paths = w.write(tmp)
            flags = ["-k", str(w.k), "-solid", str(w.solid), "-fuz", str(w.fuz), "-dist-error", str(w.dist_error),
                     "-randseed", "1", "-nb-cores", str(cores)]
            nogap = os.path.join(tmp, "nogap.fa")
            with open(nogap, "w") as f:
                f.write(">none x\n" + w.genome[:2000].tobytes().decode() + "\n")

            def run(scaf, out):
                t0 = time.perf_counter()
                subprocess.run([binp] + flags + ["-reads", paths["reads"], "-scaffolds", scaf, "-filled", out], cwd=tmp,
                               check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                return time.perf_counter() - t0
            t_build = run(nogap, os.path.join(tmp, "o0.fa"))          # builds ./.h5 (DSK + Bloom + MPHF)
            os.replace(os.path.join(tmp, ".h5"), paths["reads"] + ".h5")   # reload instead of rebuild
            t_reload = min(run(nogap, os.path.join(tmp, "o0.fa")) for _ in range(2))
            times = []
            for s in range(args.warmup + args.steps):
                t = run(paths["gaps"], os.path.join(tmp, "o1.fa"))
                if s >= args.warmup:
                    times.append(max(t - t_reload, 1e-6))
            ms = 1e3 * sum(times) / len(times)
            v = n_gaps / (ms / 1e3)
            line.update({"value": v, "ms_per_step": ms,
                         "cpu_baseline": {"value": v, "unit": "gaps/s", "cores": cores, "kind": "reference",
                                          "sample": f"all {n_gaps} gaps per step, Gap2Seq-core -nb-cores {cores}, wall minus "
                                                    f"graph reload ({t_reload:.2f} s); graph build {t_build:.1f} s untimed"},
                         "e2e": {"value": v, "unit": "gaps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    else:   # the reference binary did not travel: time the oracle port (pinned to it by tests/test_oracle_golden.py)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        g = oracle_lib.OracleGraph.from_reads(w.k, w.solid, w.reads)
        times = []
        for s in range(args.warmup + args.steps):
            oracle_lib.srand(1)
            t0 = time.perf_counter()
            for rec in w.gaps:
                g.fill_record(rec.seq, rec.name, d_err=w.dist_error, fuz=w.fuz)
            if s >= args.warmup:
                times.append(time.perf_counter() - t0)
        ms = 1e3 * sum(times) / len(times)
        v = n_gaps / (ms / 1e3)
        line.update({"value": v, "ms_per_step": ms,
                     "cpu_baseline": {"value": v, "unit": "gaps/s", "cores": 1, "kind": "port",
                                      "sample": f"all {n_gaps} gaps per step, oracle port, 1 thread"},
                     "e2e": {"value": v, "unit": "gaps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line), flush=True)


def cpu_baseline(w, keys, budget_s: float):
    """Oracle port (CPU restatement pinned to the reference) on a bounded prefix of the gap list, 1 core."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    g = oracle_lib.OracleGraph.from_kmers(w.k, keys)
    oracle_lib.srand(1)
    t0 = time.perf_counter()
    n = cells = 0
    for rec in w.gaps:
        i = rec.seq.index("N")
        res, _ = g.fill_gap(rec.seq[i - w.k - w.fuz:i], rec.seq[i + rec.gap_len:i + rec.gap_len + w.k + w.fuz],
                            rec.gap_len, w.dist_error, w.fuz, w.fuz, all_paths=not w.best_only)
        n += 1
        cells += res.cells_left + res.cells_right
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "gaps/s", "cores": 1, "kind": "port", "cells_per_s": cells / dt,
            "sample": f"first {n} of {len(w.gaps)} gaps, oracle port (oracle/g2s_oracle.cpp), {dt:.1f} s on 1 core"}


def main():
    args = parse_args()
    import faulthandler
    faulthandler.dump_traceback_later(float(os.environ.get("G2S_BENCH_WATCHDOG_S", "240")), repeat=True, file=sys.stderr)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from gap2seq_b200 import engine
    from gap2seq_b200.scaffold import record_gaps

    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    w = make_workload(args.workload, args.scale, rank)
    # host threads of the engine (tracebacks, splice): the box's cores are shared by the ranks of this node
    host_threads = max(2, min(16, (os.cpu_count() or 2) // max(1, world)))
    eng = engine.Engine(k=w.k, max_fuz=w.fuz, device=local_rank, host_threads=host_threads)
    t0 = time.perf_counter()
    if rank == 0 or world == 1:
        eng.build_from_reads(w.reads, w.solid)
    t_graph = time.perf_counter() - t0
    t_bcast, table_bytes = 0.0, 0
    if world > 1:   # the table is built once and replicated by ONE NCCL broadcast over NVLink (no per-gap collective)
        from gap2seq_b200 import shard
        barrier()
        t0 = time.perf_counter()
        table_bytes = shard.broadcast_table(eng, 0)
        barrier()
        t_bcast = time.perf_counter() - t0
    n_kmers = eng.num_kmers()
    specs = [s for rec in w.gaps for s in record_gaps(rec.seq, w.k, w.fuz, w.dist_error, all_paths=not w.best_only)]
    seqs = [g.seq for g in w.gaps]
    names = [g.name for g in w.gaps]
    n_gaps = len(specs)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    # ---- device pass, inputs resident in HBM
    eng.upload(specs)
    for _ in range(args.warmup):
        eng.run_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    kern_ms, launches = [], 0
    t_region = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        eng.run_resident()
        t = eng.timing()
        kern_ms.append(t.kernel_ms)
        launches += t.launches
    barrier()
    t_region = time.perf_counter() - t_region
    phase_sum, phase_slowest = eng.phase_cycles()
    res, fills = eng.finish_resident()
    cells = sum(r.cells_left + r.cells_right for r in res)
    filled = sum(1 for r in res if r.count > 0)
    exact = sum(1 for f, rec in zip(fills, w.gaps) if f[:-w.k] == rec.truth) if len(fills) == len(w.gaps) else None
    big = eng.timing().gaps_big_tier

    # ---- end to end through the driver call, host buffers
    import ctypes
    libc = ctypes.CDLL(None)
    libc.srand(1)
    ref_out = eng.fill_records(seqs, names, dist_error=w.dist_error, fuz=w.fuz, all_paths=not w.best_only)
    prepared = eng.prepare_records(seqs, names)   # host buffers of a C caller: NUL-terminated strings
    for _ in range(max(args.warmup, 3)):
        libc.srand(1)
        eng.fill_records_c(prepared, dist_error=w.dist_error, fuz=w.fuz, all_paths=not w.best_only)[4]()
    barrier()
    e2e_ms, h2d_b, d2h_b = [], 0, 0
    e2e_same = True
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        libc.srand(1)
        t1 = time.perf_counter()
        out = eng.fill_records_c(prepared, dist_error=w.dist_error, fuz=w.fuz, all_paths=not w.best_only)
        e2e_ms.append(1e3 * (time.perf_counter() - t1))
        t = eng.timing()
        h2d_b, d2h_b = t.h2d_bytes, t.d2h_bytes
        # outside the timed region: the timed call's output equals the driver call made through the Python binding
        e2e_same = e2e_same and out[2:4] == ref_out[2:4] and ctypes.string_at(out[1].value).decode() == ref_out[1] and \
            all(ctypes.string_at(out[0][i]).decode() == ref_out[0][i] for i in range(len(seqs)))
        out[4]()
    barrier()
    clocks = sampler.stop()
    e2e_detail = {"h2d_ms": t.h2d_ms, "kernel_ms": t.kernel_ms, "d2h_ms": t.d2h_ms, "host_ms": t.host_ms}

    ms_step = sum(kern_ms) / len(kern_ms)
    ms_e2e = sum(e2e_ms) / len(e2e_ms)
    stats = torch.tensor([ms_step, ms_e2e], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(n_gaps), float(cells), float(launches), float(h2d_b), float(d2h_b)], dtype=torch.float64,
                       device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_step, ms_e2e = stats.tolist()
    all_gaps, all_cells, all_launches, all_h2d, all_d2h = tot.tolist()

    if rank == 0:
        peak, peak_src = measured_peaks()
        achieved = cells * BYTES_PER_CELL / (ms_step / 1e3) / 1e9     # this rank's kernel, GB/s
        traffic, traffic_src = measured_traffic(args.workload, args.scale)
        line = {
            "metric": "gaps_per_s", "value": all_gaps / (ms_step / 1e3), "unit": "gaps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC.get(args.workload, args.workload), "scale": args.scale,
                       "gaps_per_gpu": n_gaps, "filled": filled, "fills_equal_to_removed_bases": exact, "solid_kmers": n_kmers,
                       "l2": "256 MiB device buffer rewritten between timed iterations",
                       "graph_build_s_untimed": round(t_graph, 3), "gaps_in_wide_tier": big, "host_threads_per_rank": host_threads,
                       "parallelism": f"{world} x (one process per GPU, replicated k-mer table, disjoint gap batches)",
                       "table_broadcast_s_untimed": round(t_bcast, 4), "table_bytes": table_bytes},
            "dp_cells_per_s": all_cells / (ms_step / 1e3), "dp_cells_per_step": all_cells,
            "e2e": {"value": all_gaps / (ms_e2e / 1e3), "unit": "gaps/s", "h2d_bytes_per_step": all_h2d,
                    "d2h_bytes_per_step": all_d2h, "ms_per_step": ms_e2e, "breakdown_ms_last_step": e2e_detail,
                    "call": "g2s_fill_records (C ABI) on host strings: scan, H2D, device pass, D2H, tracebacks, splice, log",
                    "output_identical_every_step": e2e_same},
            "gpu_launches": int(all_launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_fill_thin", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": cells * BYTES_PER_CELL, "peak_source": peak_src,
                         "algorithmic_bytes_per_cell": BYTES_PER_CELL,
                         "note": "thin gaps are chains of dependent gathers (1 cell per level): latency bound, see DESIGN.md"},
            "timed_region_wall_s": t_region,
            "phase_cycles_slowest_gap": dict(zip(PHASES, phase_slowest)),
            "phase_cycles_sum": dict(zip(PHASES, phase_sum)),
        }
        if world == 1:
            keys, _ = eng.export_kmers()
            line["cpu_baseline"] = cpu_baseline(w, keys, args.cpu_seconds)
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
