"""ctypes binding of oracle/_build/libg2s_oracle.so (the CPU restatement of the reference hot path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` leg.  The product package gap2seq_b200 never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = os.path.join(ROOT, "oracle", "_build", "libg2s_oracle.so")


class OrcResult(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("count", "left_fuz", "right_fuz", "forward_count", "n_path_lengths", "path_len0", "path_len1",
                 "reached_fuz", "order_dependent", "unable_to_backtrace", "substats_valid", "rand_calls")] + \
               [("substats", C.c_uint64 * 6), ("cells_right", C.c_uint64), ("cells_left", C.c_uint64),
                ("nodes_right", C.c_uint64), ("nodes_left", C.c_uint64), ("fill_len", C.c_int32), ("pad", C.c_int32)]


def build_oracle() -> str:
    if not os.path.exists(_LIB) or any(
            os.path.getmtime(os.path.join(ROOT, "oracle", f)) > os.path.getmtime(_LIB)
            for f in ("g2s_oracle.cpp", "g2s_oracle.hpp", "g2s_oracle_capi.cpp")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        L.orc_graph_from_reads.restype = C.c_void_p
        L.orc_graph_from_reads.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_size_t]
        L.orc_graph_from_kmers.restype = C.c_void_p
        L.orc_graph_from_kmers.argtypes = [C.c_int, C.c_void_p, C.c_size_t]
        L.orc_graph_size.restype = C.c_size_t
        L.orc_graph_size.argtypes = [C.c_void_p]
        L.orc_graph_kmers.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_graph_free.argtypes = [C.c_void_p]
        L.orc_srand.argtypes = [C.c_uint]
        L.orc_fill_gap.restype = C.c_int
        L.orc_fill_gap.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.POINTER(OrcResult), C.c_char_p, C.c_size_t]
        L.orc_fill_record.restype = C.c_void_p
        L.orc_fill_record.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class OracleGraph:
    """Exact solid k-mer set with GATB neighbour semantics (oracle/g2s_oracle.hpp Graph)."""

    def __init__(self, k: int, handle):
        self.k, self.h = k, handle

    @classmethod
    def from_reads(cls, k: int, solid: int, reads: np.ndarray) -> "OracleGraph":
        n, L = reads.shape
        buf = np.ascontiguousarray(reads).tobytes()
        off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
        return cls(k, lib().orc_graph_from_reads(k, solid, buf, off.ctypes.data, n))

    @classmethod
    def from_kmers(cls, k: int, kmers: np.ndarray) -> "OracleGraph":
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        return cls(k, lib().orc_graph_from_kmers(k, kmers.ctypes.data, kmers.size))

    def size(self) -> int:
        return lib().orc_graph_size(self.h)

    def kmers(self, with_adjacency: bool = False):
        n = self.size()
        out = np.empty(n, dtype=np.uint64)
        adj = np.empty(n, dtype=np.uint8) if with_adjacency else None
        lib().orc_graph_kmers(self.h, out.ctypes.data, adj.ctypes.data if with_adjacency else None)
        return (out, adj) if with_adjacency else out

    def fill_gap(self, left: str, right: str, gap_len: int, gap_err: int, lmf: int, rmf: int, skip_confident=False,
                 all_paths=True):
        res = OrcResult()
        cap = gap_len + self.k + gap_err + lmf + rmf + 64
        buf = C.create_string_buffer(cap)
        lib().orc_fill_gap(self.h, left.encode(), right.encode(), gap_len, self.k, gap_err, lmf, rmf,
                           int(skip_confident), int(all_paths), C.byref(res), buf, cap)
        return res, buf.value.decode()

    def fill_record(self, seq: str, comment: str, *, d_err=500, fuz=10, skip_confident=False, unique=False,
                    all_paths=True, stats=False):
        log = C.c_void_p()
        gc, fc = C.c_int(0), C.c_int(0)
        p = lib().orc_fill_record(self.h, seq.encode(), comment.encode(), self.k, d_err, fuz, int(skip_confident),
                                  int(unique), int(all_paths), int(stats), C.byref(log), C.byref(gc), C.byref(fc))
        filled = C.string_at(p).decode()
        logs = C.string_at(log.value).decode()
        lib().orc_free(p)
        lib().orc_free(log)
        return filled, logs, gc.value, fc.value

    def __del__(self):
        try:
            lib().orc_graph_free(self.h)
        except Exception:
            pass


def srand(seed: int) -> None:
    lib().orc_srand(seed)
