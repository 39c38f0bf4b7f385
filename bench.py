#!/usr/bin/env python
"""bench.py - throughput of the gap-filling hot path (Gap2Seq::fill_gap, reference src/Gap2Seq.cpp:858-1556)
on the BASELINE.json workload the metric is quoted on: `configs[3]`, the chromosome-scale synthetic genome with
10 000 gaps of 100 bp - 5 kbp ("gaps sharded across 1/2/4/8 B200"); it fits one GPU (8.6 GB k-mer table).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg4] [--scale 1.0]

One "step" = one pass of the hot path over the workload's whole gap list.
  value        gaps/s of the device pass with the gaps already resident in HBM (CUDA events on the engine's stream).
               N > 1 (torchrun, one process per GPU): ONE gap list of N x the workload's gaps (the workload's own list plus
               N - 1 lists of the same shape on the same genome) is split longest-processing-time-first over the ranks -
               weak scaling through the product's sharding, not N replicas; the k-mer table is built on rank 0 and
               replicated by one NCCL broadcast; the step time is the max over ranks.  The strong-scaling split of the
               workload's own 10 000 gaps is timed in the same run and reported under run.strong_scaling_named_workload
               (it cannot speed up: the pass is bound by the longest gap's chain of dependent gathers)
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
L2_NOTE = "GPU arm: a 256 MiB device buffer is rewritten between timed iterations (L2 is 126 MB)"
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
    ap.add_argument("--ref-gaps", type=int, default=600, help="--impl reference: gaps in the bounded sample")
    ap.add_argument("--host-threads", type=int, default=0, help="host threads of the engine (0: min(16, cores / ranks))")
    return ap.parse_args()


def measured_traffic(workload: str, scale: float) -> dict:
    """DRAM bytes per launch and L1 sectors per request of the dominant kernel from the committed `ncu --set full`
    capture of this workload (profiles/ncu_traffic.json, written from the capture named there); {} without a capture."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p)).get(f"{workload}@{scale:g}") or {}
    except Exception:
        return {}


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


def make_workload(name: str, scale: float, rank: int, lists: int = 1):
    """Every rank generates the SAME workload (seeded); only rank 0 simulates the reads (it builds the graph).
    lists > 1 (N GPUs): the gap list is `lists` gap lists of the workload's shape on the same genome, concatenated into
    ONE list - list 0 is the named workload itself, the others are fresh placements of the same gap lengths (cfg4) or
    renamed copies (workloads whose gaps are tied to their reads) - so that the work per GPU stays that of the named
    workload while the sharding, the gather in input order and the single rand() stream are those of one job."""
    import copy
    import numpy as np
    from gap2seq_b200 import synth
    w = synth.make_workload(name, scale, with_reads=(rank == 0))
    base = list(w.gaps)
    for r in range(1, lists):
        if name == "cfg4":
            rng = np.random.default_rng(7000 + r)
            extra = synth.place_gaps(w.genome, np.array([g.gap_len for g in base]), rng, flank=100)
        else:
            extra = [copy.copy(g) for g in base]
        for g in extra:
            g.name = f"list{r}_{g.name}"
        w.gaps = w.gaps + extra
    return w, len(base)


def ref_sample(w, n_gaps: int):
    """Bounded sample of the workload for the CPU arms: a prefix of the gap list and (large workloads) only the reads
    that overlap those gaps' DP windows - they fill to the same records as with all reads (tests/test_synth_sample.py)."""
    sub = w.gaps[:n_gaps]
    if len(sub) < len(w.gaps) and w.read_starts is not None and w.reads.shape[0] > 4_000_000:
        return sub, w.reads_near(sub), f"first {len(sub)} of {len(w.gaps)} gaps, reads restricted to those gaps' DP windows"
    return sub, w.reads, (f"first {len(sub)} of {len(w.gaps)} gaps, all reads" if len(sub) < len(w.gaps) else f"all {len(sub)} gaps, all reads")


def run_reference(args, rank: int, world: int, emit):
    """The reference's own CPU implementation of the path on the host cores (rank 0 only): the unmodified binary at
    -nb-cores 1, 4 and all; the timed steps run at the fastest of the three."""
    if rank != 0:
        return
    from gap2seq_b200 import synth
    w, _ = make_workload(args.workload, args.scale, 0)
    binp = os.path.join(ROOT, "oracle", "_ref", "bin", "Gap2Seq-core")
    cores = os.cpu_count() or 1
    sub, reads, sample = ref_sample(w, args.ref_gaps)
    n_gaps = len(sub)
    line = {"impl": "reference", "metric": "gaps_per_s", "unit": "gaps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": {"workload": WORKLOAD_DESC.get(args.workload, args.workload), "scale": args.scale, "l2": L2_NOTE}}
    if os.path.exists(binp):
        tmp = tempfile.mkdtemp(prefix="g2s_refbench_")
        try:
            rp, gp = os.path.join(tmp, "reads.fa"), os.path.join(tmp, "gaps.fa")
            synth.write_reads_fasta(rp, reads)
            with open(gp, "w") as f:
                for g in sub:
                    f.write(f">{g.name}\n{g.seq}\n")
            flags = ["-k", str(w.k), "-solid", str(w.solid), "-fuz", str(w.fuz), "-dist-error", str(w.dist_error), "-randseed", "1"]
            if w.best_only:
                flags.append("-best-only")
            nogap = os.path.join(tmp, "nogap.fa")
            with open(nogap, "w") as f:
                f.write(">none x\n" + w.genome[:2000].tobytes().decode() + "\n")

            def run(scaf, out, nb):
                t0 = time.perf_counter()
                subprocess.run([binp] + flags + ["-nb-cores", str(nb), "-reads", rp, "-scaffolds", scaf, "-filled", out], cwd=tmp,
                               check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                return time.perf_counter() - t0
            t_build = run(nogap, os.path.join(tmp, "o0.fa"), cores)          # builds ./.h5 (DSK + Bloom + MPHF)
            os.replace(os.path.join(tmp, ".h5"), rp + ".h5")                   # reload instead of rebuild
            t_reload = min(run(nogap, os.path.join(tmp, "o0.fa"), cores) for _ in range(2))
            by_cores = {}
            for nb in sorted({1, min(4, cores), cores}):
                by_cores[nb] = n_gaps / max(run(gp, os.path.join(tmp, "o1.fa"), nb) - t_reload, 1e-6)
            best = max(by_cores, key=by_cores.get)
            times = []
            for s in range(args.warmup + args.steps):
                t = run(gp, os.path.join(tmp, "o1.fa"), best)
                if s >= args.warmup:
                    times.append(max(t - t_reload, 1e-6))
            ms = 1e3 * sum(times) / len(times)
            v = n_gaps / (ms / 1e3)
            line.update({"value": v, "ms_per_step": ms,
                         "cpu_baseline": {"value": v, "unit": "gaps/s", "cores": best, "kind": "reference",
                                          "gaps_per_s_by_nb_cores": {str(k): round(x, 2) for k, x in by_cores.items()},
                                          "host_cores": cores,
                                          "sample": f"{sample}; one step = Gap2Seq-core -nb-cores {best} over the sample, wall minus "
                                                    f"graph reload ({t_reload:.2f} s); graph build {t_build:.1f} s untimed"},
                         "e2e": {"value": v, "unit": "gaps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    else:   # the reference binary did not travel: time the oracle port (pinned to it by tests/test_oracle_golden.py)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        g = oracle_lib.OracleGraph.from_reads(w.k, w.solid, reads)
        times = []
        for s in range(args.warmup + args.steps):
            oracle_lib.srand(1)
            t0 = time.perf_counter()
            for rec in sub:
                g.fill_record(rec.seq, rec.name, d_err=w.dist_error, fuz=w.fuz, all_paths=not w.best_only)
            if s >= args.warmup:
                times.append(time.perf_counter() - t0)
        ms = 1e3 * sum(times) / len(times)
        v = n_gaps / (ms / 1e3)
        line.update({"value": v, "ms_per_step": ms,
                     "cpu_baseline": {"value": v, "unit": "gaps/s", "cores": 1, "kind": "port",
                                      "sample": f"{sample}; oracle port, 1 thread"},
                     "e2e": {"value": v, "unit": "gaps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    emit(line)


def cpu_baseline_and_check(w, keys, engine_out, budget_s: float):
    """Oracle port (CPU restatement pinned to the reference) on a bounded prefix of the records, 1 core - and, outside
    every timed region of the engine, the check of the engine's step output against it: record by record, the same
    rand() stream (srand(1), input order).  A mismatch is fatal."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    g = oracle_lib.OracleGraph.from_kmers(w.k, keys)
    oracle_lib.srand(1)
    t0 = time.perf_counter()
    n = 0
    for rec, ours in zip(w.gaps, engine_out):
        out = g.fill_record(rec.seq, rec.name, d_err=w.dist_error, fuz=w.fuz, all_paths=not w.best_only)
        if out[0] != ours:
            sys.exit(f"bench.py: record {rec.name!r} filled by the engine differs from the oracle")
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "gaps/s", "cores": 1, "kind": "port",
            "sample": f"first {n} of {len(w.gaps)} records, oracle port (oracle/g2s_oracle.cpp), {dt:.1f} s on 1 core"}, n


def main():
    args = parse_args()
    # the ONE JSON line goes to the real stdout; whatever libraries print (NCCL's version banner ...) goes to stderr
    out_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict):
        os.write(out_fd, (json.dumps(line) + "\n").encode())
    import faulthandler
    faulthandler.dump_traceback_later(float(os.environ.get("G2S_BENCH_WATCHDOG_S", "240")), repeat=True, file=sys.stderr)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world, emit)
        return

    import ctypes
    import numpy as np
    # the e2e leg receives megabytes of malloc'ed strings per call: this PROCESS keeps freed memory instead of returning
    # it to the kernel (M_TRIM_THRESHOLD = -1, M_MMAP_THRESHOLD = -3) - an application decision, the library leaves it alone
    ctypes.CDLL(None).mallopt(-1, 1 << 30)
    ctypes.CDLL(None).mallopt(-3, 1 << 28)
    import torch
    import torch.distributed as dist
    from gap2seq_b200 import engine, shard
    from gap2seq_b200.scaffold import record_gaps

    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")   # waits of idle ranks must not spin on their GPUs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    w, n_base = make_workload(args.workload, args.scale, rank, lists=world)
    # host threads of the engine (subgraph analysis, tracebacks, splice)
    cores = os.cpu_count() or 2
    host_threads = args.host_threads if args.host_threads > 0 else max(2, min(16, cores // max(1, world)))
    eng = engine.Engine(k=w.k, max_fuz=w.fuz, device=local_rank, host_threads=host_threads)
    t0 = time.perf_counter()
    if rank == 0:
        eng.build_from_reads(w.reads, w.solid)
    t_graph = time.perf_counter() - t0
    t_bcast, table_bytes = 0.0, 0
    if world > 1:   # the table is built once and replicated by ONE NCCL broadcast over NVLink (no per-gap collective)
        barrier()
        t0 = time.perf_counter()
        table_bytes = shard.broadcast_table(eng, 0)
        barrier()
        t_bcast = time.perf_counter() - t0
    n_kmers = eng.num_kmers()
    per_rec = [record_gaps(rec.seq, w.k, w.fuz, w.dist_error, all_paths=not w.best_only) for rec in w.gaps]
    specs = [s for r in per_rec for s in r]
    seqs = [g.seq for g in w.gaps]
    names = [g.name for g in w.gaps]
    n_all = len(specs)
    # ONE gap list, longest-processing-time-first over the ranks.  Weak scaling: the list holds N times the
    # named workload's gaps (make_workload); the strong-scaling split of the named workload's own gaps is timed below too
    costs = [shard.estimate_cost(s.gap_len, s.gap_err, w.fuz) for s in specs]
    shards = shard.lpt_shards(costs, world)
    mine = [specs[i] for i in shards[rank]]
    n_base_specs = sum(len(r) for r in per_rec[:n_base])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    # ---- device pass, inputs resident in HBM
    eng.upload(mine)
    for _ in range(args.warmup):
        eng.run_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    kern_ms, launches = [], 0
    tier_ms = {"chain": 0.0, "thin": 0.0, "wide": 0.0}
    t_region = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        eng.run_resident()
        t = eng.timing()
        kern_ms.append(t.kernel_ms)
        launches += t.launches
        tier_ms["chain"] += t.kernel_ms_chain; tier_ms["thin"] += t.kernel_ms_thin; tier_ms["wide"] += t.kernel_ms_wide
    barrier()
    t_region = time.perf_counter() - t_region
    t_dev = eng.timing()
    phase_sum, phase_slowest = eng.phase_cycles()
    res, fills = eng.finish_resident()
    cells = sum(r.cells_left + r.cells_right for r in res)
    filled = sum(1 for r in res if r.count > 0)
    exact = None
    if world == 1 and len(fills) == len(w.gaps):
        exact = sum(1 for f, rec in zip(fills, w.gaps) if f[:-w.k] == rec.truth)

    # ---- supplementary: strong scaling of the named workload's own gap list over the N GPUs (device pass only)
    strong_ms = None
    if world > 1:
        sh0 = shard.lpt_shards(costs[:n_base_specs], world)
        eng.upload([specs[i] for i in sh0[rank]])
        for _ in range(args.warmup):
            eng.run_resident()
        barrier()
        acc = []
        for _ in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            eng.run_resident()
            acc.append(eng.timing().kernel_ms)
        barrier()
        sm = torch.tensor([sum(acc) / len(acc)], dtype=torch.float64, device="cuda")
        dist.all_reduce(sm, op=dist.ReduceOp.MAX)
        strong_ms = sm.item()

    # ---- end to end through the driver call, host buffers.  N = 1: this rank's engine.  N > 1: ONE engine on rank 0 owns
    # all N GPUs (g2s_create_multi); the other ranks free their GPUs and wait on the host.
    libc = ctypes.CDLL(None)
    ref_out = None
    if rank == 0:
        libc.srand(1)
        ref_out = eng.fill_records(seqs, names, dist_error=w.dist_error, fuz=w.fuz, all_paths=not w.best_only)
    keys = None
    if rank == 0:
        keys, _ = eng.export_kmers()
    e2e_eng = eng
    multi_note = None
    if world > 1:
        eng.close()
        dist.barrier(group=cpu_group)
        if rank == 0:
            e2e_eng = engine.Engine(k=w.k, max_fuz=w.fuz, host_threads=min(32, cores), devices=list(range(world)))
            t0 = time.perf_counter()
            e2e_eng.import_kmers(keys)
            multi_note = f"g2s_create_multi over GPUs 0..{world - 1}; table import + replication {time.perf_counter() - t0:.2f} s untimed"
    e2e_ms, h2d_b, d2h_b, e2e_same, e2e_detail = [0.0], 0, 0, True, {}
    if rank == 0:
        prepared = e2e_eng.prepare_records(seqs, names)   # host buffers of a C caller: NUL-terminated strings
        kw = dict(dist_error=w.dist_error, fuz=w.fuz, all_paths=not w.best_only)
        for _ in range(max(args.warmup, 3)):
            libc.srand(1)
            e2e_eng.fill_records_c(prepared, **kw)[4]()
        e2e_ms = []
        for _ in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            libc.srand(1)
            t1 = time.perf_counter()
            out = e2e_eng.fill_records_c(prepared, **kw)
            e2e_ms.append(1e3 * (time.perf_counter() - t1))
            t = e2e_eng.timing()
            h2d_b, d2h_b = t.h2d_bytes, t.d2h_bytes
            # outside the timed region: the timed call's output equals the one-GPU driver call made through the Python binding
            e2e_same = e2e_same and out[2:4] == ref_out[2:4] and ctypes.string_at(out[1].value).decode() == ref_out[1] and \
                all(ctypes.string_at(out[0][i]).decode() == ref_out[0][i] for i in range(len(seqs)))
            out[4]()
        e2e_detail = {"h2d_ms": t.h2d_ms, "kernel_ms": t.kernel_ms, "d2h_ms": t.d2h_ms, "host_ms": t.host_ms}
        if not e2e_same:
            sys.exit("bench.py: the timed driver call's output differs from the one-GPU output")
    if world > 1:
        dist.barrier(group=cpu_group)
    clocks = sampler.stop()

    ms_step = sum(kern_ms) / len(kern_ms)
    ms_e2e = sum(e2e_ms) / len(e2e_ms)
    stats = torch.tensor([ms_step], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(len(mine)), float(cells), float(launches), float(filled), float(t_dev.table_lookups), float(t_dev.node_ops)],
                       dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_max = stats.tolist()[0]
    all_gaps, all_cells, all_launches, all_filled, all_lookups, all_nodeops = tot.tolist()

    if rank == 0:
        peak, peak_src = measured_peaks()
        # the kernel with the largest share of this rank's step, the cells of the gaps it finished, its mean duration
        dom = max(tier_ms, key=tier_ms.get)
        dom_ms = tier_ms[dom] / args.steps
        dom_cells = {"chain": t_dev.cells_chain, "thin": t_dev.cells_thin, "wide": t_dev.cells_wide}[dom]
        shapes = {0: "k_fill_thin<1,1>", 1: "k_fill_thin<2,4>", 2: "k_fill_thin<4,8>"}
        dom_name = {"chain": "k_chain_walk + k_chain_finish", "thin": shapes.get(t_dev.thin_shape, "k_fill_thin"), "wide": "k_fill_cta" if t_dev.wide_cluster <= 1 else f"k_fill_cluster (cluster of {t_dev.wide_cluster} CTAs per gap)"}[dom]
        achieved = dom_cells * BYTES_PER_CELL / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else 0.0
        traffic = measured_traffic(args.workload, args.scale)
        cpu, n_checked = cpu_baseline_and_check(w, keys, ref_out[0], args.cpu_seconds) if world == 1 else (None, 0)
        line = {
            "metric": "gaps_per_s", "value": all_gaps / (ms_max / 1e3), "unit": "gaps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC.get(args.workload, args.workload), "scale": args.scale, "l2": L2_NOTE},
            "run": {"gaps_total": int(all_gaps), "gaps_this_rank": len(mine), "filled": int(all_filled),
                       "fills_equal_to_removed_bases": exact, "solid_kmers": n_kmers,
                       "graph_build_s_untimed": round(t_graph, 3), "host_threads": host_threads if world == 1 else min(32, cores),
                       "parallelism": f"{world} x (one process per GPU, replicated k-mer table, ONE list of {world} x {n_base_specs} gaps split "
                                      "longest-processing-time-first; per-GPU work = the named workload)",
                       "strong_scaling_named_workload": None if strong_ms is None else {
                           "gaps": n_base_specs, "ms_per_step": strong_ms, "gaps_per_s": n_base_specs / (strong_ms / 1e3),
                           "note": "the named workload's own gap list split over the N GPUs: the pass is bound by the longest gap's "
                                   "chain of dependent gathers, which sharding does not shorten (DESIGN.md)"},
                       "table_broadcast_s_untimed": round(t_bcast, 4), "table_bytes": table_bytes,
                       "records_checked_against_oracle": n_checked},
            "dp_cells_per_s": all_cells / (ms_max / 1e3), "dp_cells_per_step": all_cells,
            "hash_probes_per_s": (all_lookups + all_nodeops) / (ms_max / 1e3),
            "hash_probes_per_step": {"kmer_table_gathers": all_lookups, "node_table_ops": all_nodeops},
            "e2e": {"value": len(specs) / (ms_e2e / 1e3), "unit": "gaps/s", "h2d_bytes_per_step": h2d_b,
                    "d2h_bytes_per_step": d2h_b, "ms_per_step": ms_e2e, "breakdown_ms_last_step": e2e_detail,
                    "call": "g2s_fill_records (C ABI) on host strings: scan, H2D, device pass, D2H, tracebacks, splice, log",
                    "engine": multi_note or "g2s_create on GPU 0", "output_identical_every_step": e2e_same},
            "gpu_launches": int(all_launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic.get("dram_bytes_per_launch"),
                         "traffic_source": traffic.get("source"), "sectors_per_request": traffic.get("sectors_per_request"),
                         "algorithmic_bytes_per_launch": dom_cells * BYTES_PER_CELL, "peak_source": peak_src,
                         "algorithmic_bytes_per_cell": BYTES_PER_CELL, "kernel_ms_per_launch": dom_ms,
                         "kernel_share_of_step": dom_ms / (sum(tier_ms.values()) / args.steps) if sum(tier_ms.values()) else None,
                         "tier_ms_per_step": {k: v / args.steps for k, v in tier_ms.items()},
                         "tier_cells": {"chain": t_dev.cells_chain, "thin": t_dev.cells_thin, "wide": t_dev.cells_wide},
                         "note": "rank 0's kernel; thin gaps are chains of dependent gathers: latency bound, see DESIGN.md"},
            "timed_region_wall_s": t_region,
            "phase_cycles_slowest_gap": dict(zip(PHASES, phase_slowest)),
            "phase_cycles_sum": dict(zip(PHASES, phase_sum)),
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    if e2e_eng is not None:
        e2e_eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
