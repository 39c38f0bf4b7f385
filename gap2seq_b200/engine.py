"""ctypes binding of libg2s_gpu.so - the C ABI declared in include/g2s.h.

This is the Python-side mirror of the reference's only call boundary for the hot path,
``Gap2Seq::fill_gap`` (reference src/Gap2Seq.hpp:74-76; callers src/Gap2Seq.cpp:252 and :380), plus the gap
driver around it (``Gap2Seq::execute``'s scanner, src/Gap2Seq.cpp:285-434).  All compute happens in the CUDA
library; there is no Python or CPU fallback: if the library is missing, or no CUDA device is present, the
calls raise ``G2SError``.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libg2s_gpu.so")

ST_ORDER_DEPENDENT, ST_MEM_LIMIT, ST_GENERAL_SUBGRAPH, ST_UNABLE_TO_BACKTRACE, ST_BIG_TIER = 1, 2, 4, 8, 16
ST_DEVICE_TRACED = 32
GAP_SKIP_CONFIDENT, GAP_ALL_PATHS = 1, 2
ABI_VERSION = 4   # G2S_ABI_VERSION of include/g2s.h


class G2SError(RuntimeError):
    pass


class G2SConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("k", C.c_int32), ("max_fuz", C.c_int32),
                ("max_mem_bytes", C.c_int64), ("host_threads", C.c_int32), ("reserved", C.c_int32)]


class G2SSubgraphStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("vertices", "edges", "nontrivial_components",
                                          "size_nontrivial_components", "vertices_final", "edges_final")]

    def astuple(self):
        return tuple(getattr(self, n) for n, _ in self._fields_)


class G2SGap(C.Structure):
    _fields_ = [("kmer_left", C.c_char_p), ("kmer_right", C.c_char_p), ("gap_len", C.c_int32), ("gap_err", C.c_int32),
                ("left_max_fuz", C.c_int32), ("right_max_fuz", C.c_int32), ("flags", C.c_int32), ("graph", C.c_int32),
                ("kmer_right_starts", C.c_char_p), ("present_bits", C.c_uint64)]


class G2SResult(C.Structure):
    _fields_ = [("count", C.c_int32), ("left_fuz", C.c_int32), ("right_fuz", C.c_int32), ("status", C.c_int32),
                ("fill_len", C.c_int32), ("substats_valid", C.c_int32), ("forward_count", C.c_int32),
                ("n_path_lengths", C.c_int32), ("path_lengths", C.c_int32 * 2), ("reached_fuz", C.c_int32),
                ("rand_calls", C.c_int32), ("substats", G2SSubgraphStats), ("cells_right", C.c_uint64),
                ("cells_left", C.c_uint64), ("nodes", C.c_uint64)]


class G2STiming(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("kernel_ms", C.c_double), ("d2h_ms", C.c_double), ("host_ms", C.c_double),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("launches", C.c_uint32),
                ("gaps_big_tier", C.c_uint32), ("max_tier_run", C.c_uint32), ("gaps_rerun_scratch", C.c_uint32),
                ("gaps_rerun_out_full", C.c_uint32), ("gaps_mem_limit", C.c_uint32),
                ("table_lookups", C.c_uint64), ("node_ops", C.c_uint64),
                ("kernel_ms_thin", C.c_double), ("kernel_ms_wide", C.c_double), ("launches_thin", C.c_uint32),
                ("launches_wide", C.c_uint32), ("thin_shape", C.c_uint32), ("wide_cluster", C.c_uint32),
                ("cells_thin", C.c_uint64), ("cells_wide", C.c_uint64), ("kernel_ms_chain", C.c_double),
                ("launches_chain", C.c_uint32), ("gaps_chain", C.c_uint32), ("cells_chain", C.c_uint64)]


class G2SRunOptions(C.Structure):
    _fields_ = [("k", C.c_int32), ("dist_error", C.c_int32), ("fuz", C.c_int32), ("skip_confident", C.c_int32),
                ("unique_paths", C.c_int32), ("all_paths", C.c_int32)]


# every symbol include/g2s.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = ["g2s_last_error", "g2s_abi_version", "g2s_create", "g2s_create_multi", "g2s_num_devices", "g2s_destroy", "g2s_import_kmers", "g2s_h5_read_solid", "g2s_import_h5",
               "g2s_build_from_reads", "g2s_build_from_read_files", "g2s_build_graphs_from_reads", "g2s_graph_num_kmers", "g2s_table_raw", "g2s_table_adopt", "g2s_num_kmers", "g2s_export_kmers", "g2s_contains", "g2s_fill_gap", "g2s_fill_gap_ex",
               "g2s_fill_batch", "g2s_upload_gaps", "g2s_run_resident", "g2s_finish_resident", "g2s_submit",
               "g2s_batch_peek", "g2s_batch_trace", "g2s_batch_free", "g2s_fill_records", "g2s_free",
               "g2s_get_timing", "g2s_get_phase_cycles"]

_lib = None


def load_library(path: str | None = None):
    """Load libg2s_gpu.so; fails loudly when it has not been built (run __graft_entry__.build())."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("G2S_LIB") or LIB_PATH   # G2S_LIB: an experiment build of the same ABI
    if not os.path.exists(p):
        raise G2SError(f"{p} is missing: build it with `make -C gap2seq_b200/csrc` (there is no CPU fallback)")
    L = C.CDLL(p)
    L.g2s_last_error.restype = C.c_char_p
    L.g2s_create.argtypes = [C.POINTER(G2SConfig), C.POINTER(C.c_void_p)]
    L.g2s_create_multi.argtypes = [C.POINTER(G2SConfig), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_void_p)]
    L.g2s_num_devices.argtypes = [C.c_void_p]
    L.g2s_destroy.argtypes = [C.c_void_p]
    L.g2s_import_kmers.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.g2s_build_from_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    L.g2s_build_from_read_files.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_size_t, C.c_int, C.POINTER(C.c_uint64)]
    L.g2s_build_graphs_from_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    L.g2s_graph_num_kmers.argtypes = [C.c_void_p, C.c_size_t]
    L.g2s_graph_num_kmers.restype = C.c_int64
    L.g2s_h5_read_solid.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_int32)]
    L.g2s_import_h5.argtypes = [C.c_void_p, C.c_char_p]
    L.g2s_free.argtypes = [C.c_void_p]
    L.g2s_table_raw.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                C.POINTER(C.c_uint64)]
    L.g2s_table_adopt.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_void_p)]
    L.g2s_num_kmers.restype = C.c_int64
    L.g2s_num_kmers.argtypes = [C.c_void_p]
    L.g2s_export_kmers.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.g2s_contains.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p]
    L.g2s_fill_gap.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_longlong, C.c_char_p, C.c_int, C.c_int,
                               C.POINTER(G2SSubgraphStats)]
    L.g2s_fill_batch.argtypes = [C.c_void_p, C.POINTER(G2SGap), C.c_size_t, C.POINTER(G2SResult), C.POINTER(C.c_char_p)]
    L.g2s_upload_gaps.argtypes = [C.c_void_p, C.POINTER(G2SGap), C.c_size_t]
    L.g2s_run_resident.argtypes = [C.c_void_p]
    L.g2s_finish_resident.argtypes = [C.c_void_p, C.POINTER(G2SResult), C.POINTER(C.c_char_p)]
    L.g2s_fill_records.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_size_t,
                                   C.POINTER(G2SRunOptions), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_size_t)]
    L.g2s_free.argtypes = [C.c_void_p]
    L.g2s_get_timing.argtypes = [C.c_void_p, C.POINTER(G2STiming)]
    L.g2s_get_phase_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    if path is None:
        _lib = L
    return L


@dataclass
class GapSpec:
    """Arguments of one fill_gap call (reference src/Gap2Seq.hpp:74-76)."""
    left: str
    right: str
    gap_len: int
    gap_err: int = 500
    left_max_fuz: int = 10
    right_max_fuz: int = 10
    skip_confident: bool = False
    all_paths: bool = True
    right_starts: str | None = None   # see g2s_gap.kmer_right_starts (right flank longer than k + right_max_fuz)
    present_bits: int | None = None   # see g2s_gap.present_bits: the caller's own contains() answers for the flank k-mers
    graph: int = 0                    # the gap's graph after build_graphs_from_reads (per-gap filtered-reads mode)

    def fill_capacity(self, k: int) -> int:
        return self.gap_len + k + self.gap_err + self.left_max_fuz + self.right_max_fuz + 3


def pack_gaps(gaps: list[GapSpec]):
    """list[GapSpec] -> (ctypes array of g2s_gap, keep-alive list)."""
    arr = (G2SGap * max(len(gaps), 1))()
    keep = []
    for i, g in enumerate(gaps):
        l, r = g.left.encode(), g.right.encode()
        keep += [l, r]
        arr[i].kmer_left, arr[i].kmer_right = l, r
        arr[i].gap_len, arr[i].gap_err = g.gap_len, g.gap_err
        arr[i].left_max_fuz, arr[i].right_max_fuz = g.left_max_fuz, g.right_max_fuz
        arr[i].flags = (GAP_SKIP_CONFIDENT if g.skip_confident else 0) | (GAP_ALL_PATHS if g.all_paths else 0)
        arr[i].graph = g.graph
        if g.present_bits is not None:
            arr[i].flags |= 4   # G2S_GAP_PRESENT_BITS
            arr[i].present_bits = g.present_bits
        if g.right_starts is not None:
            rs = g.right_starts.encode()
            keep.append(rs)
            arr[i].kmer_right_starts = rs
    return arr, keep


class Engine:
    """One device-resident de Bruijn graph + the gap-filling passes that run against it."""

    def __init__(self, k: int = 31, max_fuz: int = 10, device: int = 0, max_mem_bytes: int = 0, host_threads: int = 0,
                 lib=None, devices: list[int] | None = None):
        """`devices`: several GPUs of the box behind one engine (g2s_create_multi): every wave of gaps is sharded
        over them and gathered in input order; `device` is ignored then."""
        self.L = lib or load_library()
        self.k = k
        cfg = G2SConfig(ABI_VERSION, device, k, max_fuz, max_mem_bytes, host_threads, 0)
        h = C.c_void_p()
        if devices is not None:
            arr = (C.c_int32 * len(devices))(*devices)
            self._check(self.L.g2s_create_multi(C.byref(cfg), arr, len(devices), C.byref(h)))
        else:
            self._check(self.L.g2s_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self._resident = None

    def _check(self, rc: int) -> int:
        if rc < 0:
            raise G2SError(f"libg2s_gpu error {rc}: {self.L.g2s_last_error().decode()}")
        return rc

    def close(self):
        if getattr(self, "h", None):
            self.L.g2s_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- graph
    def import_kmers(self, canon: np.ndarray) -> None:
        canon = np.ascontiguousarray(canon, dtype=np.uint64)
        self._check(self.L.g2s_import_kmers(self.h, canon.ctypes.data, canon.size))

    def build_from_read_files(self, paths: list[str], solid: int = 2) -> int:
        """Graph from FASTA / FASTQ files, parsed on the device; returns the number of records found."""
        arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
        n = C.c_uint64(0)
        self._check(self.L.g2s_build_from_read_files(self.h, arr, len(paths), solid, C.byref(n)))
        return int(n.value)

    def build_graphs_from_reads(self, read_sets: list, solid: int = 2) -> None:
        """Per-gap filtered-reads mode: graph g is built from read_sets[g] (a [n, read_len] uint8 ASCII array, or a list
        of byte strings); all graphs in one device pass.  Gaps then name their graph (fill_batch(..., graphs=[...]))."""
        chunks, lens, first = [], [], [0]
        for rs in read_sets:
            if isinstance(rs, np.ndarray):
                rs = np.ascontiguousarray(rs, dtype=np.uint8)
                if rs.size:
                    chunks.append(rs.reshape(-1))
                    lens.extend([rs.shape[1]] * rs.shape[0])
                first.append(first[-1] + (rs.shape[0] if rs.size else 0))
            else:
                for r in rs:
                    b = r if isinstance(r, (bytes, bytearray)) else str(r).encode()
                    chunks.append(np.frombuffer(bytes(b), dtype=np.uint8))
                    lens.append(len(b))
                first.append(first[-1] + len(rs))
        bases = np.concatenate(chunks) if chunks else np.zeros(1, dtype=np.uint8)
        off = np.concatenate([[0], np.cumsum(np.asarray(lens, dtype=np.uint64))]).astype(np.uint64)
        first = np.asarray(first, dtype=np.uint64)
        self._check(self.L.g2s_build_graphs_from_reads(self.h, bases.ctypes.data, off.ctypes.data, len(lens), first.ctypes.data,
                                                       len(read_sets), solid))

    def graph_num_kmers(self, graph: int) -> int:
        return int(self.L.g2s_graph_num_kmers(self.h, graph))

    def import_h5(self, path: str) -> None:
        """Import the solid k-mers of a GATB graph file (<reads>.h5, group dsk/solid)."""
        self._check(self.L.g2s_import_h5(self.h, path.encode()))

    def build_from_reads(self, reads: np.ndarray, solid: int = 2) -> None:
        """reads: [n_reads, read_len] uint8 ASCII."""
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        n, length = reads.shape
        off = np.arange(n + 1, dtype=np.uint64) * np.uint64(length)
        self._check(self.L.g2s_build_from_reads(self.h, reads.ctypes.data, off.ctypes.data, n, solid))

    def table_raw(self):
        """(device pointer, capacity in slots, k-mers, bytes) of the built table - for replication to other GPUs."""
        p, cap, n, nb = C.c_void_p(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self.L.g2s_table_raw(self.h, C.byref(p), C.byref(cap), C.byref(n), C.byref(nb)))
        return p.value, cap.value, n.value, nb.value

    def table_adopt(self, capacity_slots: int, n_kmers: int) -> int:
        """Allocate an empty table of the given shape; returns its device pointer (fill it, e.g. by broadcast)."""
        p = C.c_void_p()
        self._check(self.L.g2s_table_adopt(self.h, capacity_slots, n_kmers, C.byref(p)))
        return p.value

    def num_kmers(self) -> int:
        return int(self.L.g2s_num_kmers(self.h))

    def export_kmers(self):
        n = self.num_kmers()
        keys = np.empty(n, dtype=np.uint64)
        adj = np.empty(n, dtype=np.uint8)
        self._check(self.L.g2s_export_kmers(self.h, keys.ctypes.data, adj.ctypes.data))
        return keys, adj

    def contains(self, kmers: list[str]) -> np.ndarray:
        buf = "".join(kmers).encode()
        out = np.zeros(len(kmers), dtype=np.uint8)
        self._check(self.L.g2s_contains(self.h, buf, len(kmers), out.ctypes.data))
        return out.astype(bool)

    # ---- fill
    def fill_gap(self, g: GapSpec):
        """The drop-in call: returns (count, left_fuz, right_fuz, fill string, substats tuple)."""
        lf, rf = C.c_int(0), C.c_int(0)
        buf = C.create_string_buffer(g.fill_capacity(self.k))
        ss = G2SSubgraphStats()
        cnt = self.L.g2s_fill_gap(self.h, g.left.encode(), g.right.encode(), g.gap_len, self.k, g.gap_err,
                                  g.left_max_fuz, g.right_max_fuz, C.byref(lf), C.byref(rf), 0, buf,
                                  int(g.skip_confident), int(g.all_paths), C.byref(ss))
        if cnt < -1:
            self._check(cnt)
        fill = C.string_at(C.addressof(buf) + (g.left_max_fuz - lf.value)).decode() if buf.raw.strip(b"\0") else ""
        return cnt, lf.value, rf.value, fill, ss.astuple()

    def _out_buffers(self, gaps: list[GapSpec]):
        bufs = [C.create_string_buffer(g.fill_capacity(self.k)) for g in gaps]
        ptrs = (C.c_char_p * max(len(gaps), 1))()
        for i, b in enumerate(bufs):
            ptrs[i] = C.cast(b, C.c_char_p)
        return bufs, ptrs

    def fill_batch(self, gaps: list[GapSpec]):
        """n gaps in one device pass; rand() is consumed in list order.  Returns (results, fills)."""
        arr, keep = pack_gaps(gaps)
        res = (G2SResult * max(len(gaps), 1))()
        bufs, ptrs = self._out_buffers(gaps)
        self._check(self.L.g2s_fill_batch(self.h, arr, len(gaps), res, ptrs))
        return [res[i] for i in range(len(gaps))], [b.value.decode() for b in bufs]

    def upload(self, gaps: list[GapSpec]) -> None:
        arr, keep = pack_gaps(gaps)
        self._check(self.L.g2s_upload_gaps(self.h, arr, len(gaps)))
        self._resident = gaps

    def run_resident(self) -> None:
        self._check(self.L.g2s_run_resident(self.h))

    def finish_resident(self):
        gaps = self._resident
        res = (G2SResult * max(len(gaps), 1))()
        bufs, ptrs = self._out_buffers(gaps)
        self._check(self.L.g2s_finish_resident(self.h, res, ptrs))
        return [res[i] for i in range(len(gaps))], [b.value.decode() for b in bufs]

    def fill_records(self, seqs: list[str], comments: list[str], *, dist_error=500, fuz=10, skip_confident=False,
                     unique=False, all_paths=True, fn=None):
        """Gap driver of Gap2Seq::execute(): returns (filled seqs, log text, gapcount, filledcount, per-gap results)."""
        n = len(seqs)
        sa = (C.c_char_p * max(n, 1))(*[s.encode() for s in seqs])
        ca = (C.c_char_p * max(n, 1))(*[c.encode() for c in comments])
        opt = G2SRunOptions(self.k, dist_error, fuz, int(skip_confident), int(unique), int(all_paths))
        filled = (C.c_void_p * max(n, 1))()
        log = C.c_void_p()
        gc, fc = C.c_int32(0), C.c_int32(0)
        gr, ngr = C.c_void_p(), C.c_size_t(0)
        f = fn or self.L.g2s_fill_records
        self._check(f(self.h, sa, ca, n, C.byref(opt), filled, C.byref(log), C.byref(gc), C.byref(fc), C.byref(gr),
                      C.byref(ngr)))
        out = [C.string_at(filled[i]).decode() for i in range(n)]
        logs = C.string_at(log.value).decode()
        results = []
        if gr.value:
            arr = C.cast(gr, C.POINTER(G2SResult))
            results = [G2SResult.from_buffer_copy(arr[i]) for i in range(ngr.value)]
        free = self.L.g2s_free
        for i in range(n):
            free(filled[i])
        free(log)
        if gr.value:
            free(gr)
        return out, logs, gc.value, fc.value, results

    def prepare_records(self, seqs: list[str], comments: list[str]):
        """Host buffers (NUL-terminated C strings) for fill_records_c: what a C/C++ caller of g2s_fill_records holds."""
        n = len(seqs)
        return (n, (C.c_char_p * max(n, 1))(*[s.encode() for s in seqs]), (C.c_char_p * max(n, 1))(*[c.encode() for c in comments]))

    def fill_records_c(self, prepared, *, dist_error=500, fuz=10, skip_confident=False, unique=False, all_paths=True,
                       want_results=False):
        """g2s_fill_records on prepared host buffers; returns the raw outputs (filled char**, log char*, counts) and a
        release function, so that a caller can time exactly the C-ABI call."""
        n, sa, ca = prepared
        opt = G2SRunOptions(self.k, dist_error, fuz, int(skip_confident), int(unique), int(all_paths))
        filled = (C.c_void_p * max(n, 1))()
        log = C.c_void_p()
        gc, fc = C.c_int32(0), C.c_int32(0)
        gr, ngr = C.c_void_p(), C.c_size_t(0)
        self._check(self.L.g2s_fill_records(self.h, sa, ca, n, C.byref(opt), filled, C.byref(log), C.byref(gc), C.byref(fc),
                                            C.byref(gr) if want_results else None, C.byref(ngr) if want_results else None))

        def release():
            free = self.L.g2s_free
            for i in range(n):
                free(filled[i])
            free(log)
            if gr.value:
                free(gr)
        return filled, log, gc.value, fc.value, release

    def phase_cycles(self):
        """SM cycles per phase (init, right, forward, backward, pack, total) of the last device pass: (sum, slowest gap)."""
        a, b = (C.c_double * 14)(), (C.c_double * 14)()
        self._check(self.L.g2s_get_phase_cycles(self.h, a, b))
        return list(a), list(b)

    def timing(self) -> G2STiming:
        t = G2STiming()
        self._check(self.L.g2s_get_timing(self.h, C.byref(t)))
        return t


def h5_solid_kmers(path: str):
    """(canonical k-mers as uint64 array, k-mer size) of a GATB graph file - host only, no device needed."""
    L = load_library()
    p, n, ks = C.c_void_p(), C.c_size_t(), C.c_int32()
    if L.g2s_h5_read_solid(path.encode(), C.byref(p), C.byref(n), C.byref(ks)) != 0:
        raise G2SError(L.g2s_last_error().decode())
    try:
        out = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(max(n.value, 1),))[:n.value].copy()
    finally:
        L.g2s_free(p)
    return out, ks.value
