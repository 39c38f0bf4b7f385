"""TEST INFRASTRUCTURE ONLY: ctypes binding of tests/emul/libg2s_emul.so, the host (-DG2S_HOST_EMUL, group size 1)
build of the per-gap device algorithm.  Lets the CPU suite check the kernel's sequential logic, the host finish
stage and the scaffold driver against the golden outputs without a GPU.  Never imported by the product."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from gap2seq_b200 import engine as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "tests", "emul", "libg2s_emul.so")
_SRC = [os.path.join(ROOT, "tests", "emul", "g2s_emul.cpp")] + [
    os.path.join(ROOT, "gap2seq_b200", "csrc", f) for f in ("g2s_fill.cuh", "g2s_chain.cuh", "g2s_common.cuh", "g2s_finish.h", "g2s_scaffold.h", "g2s_pool.h")]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in _SRC):
            subprocess.run(["g++", "-O2", "-std=c++17", "-DG2S_HOST_EMUL", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                            _SRC[0], "-o", _SO, "-lpthread"], check=True)
        L = C.CDLL(_SO)
        L.emul_graph_from_kmers.restype = C.c_void_p
        L.emul_graph_from_kmers.argtypes = [C.c_int, C.c_void_p, C.c_size_t]
        L.emul_graph_free.argtypes = [C.c_void_p]
        L.emul_adjacency.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.emul_fill_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(E.G2SGap), C.c_size_t, C.POINTER(E.G2SResult),
                                      C.POINTER(C.c_char_p), C.c_int]
        L.emul_fill_records.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_size_t,
                                        C.POINTER(E.G2SRunOptions), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_size_t)]
        L.emul_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class EmulEngine:
    def __init__(self, k: int, kmers: np.ndarray):
        self.k = k
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        self.h = lib().emul_graph_from_kmers(k, kmers.ctypes.data, kmers.size)

    def adjacency(self, kmers: np.ndarray) -> np.ndarray:
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        out = np.empty(kmers.size, dtype=np.uint8)
        lib().emul_adjacency(self.h, kmers.ctypes.data, kmers.size, out.ctypes.data)
        return out

    def fill_batch(self, gaps, cap0: int = 0):
        arr, keep = E.pack_gaps(gaps)
        res = (E.G2SResult * max(len(gaps), 1))()
        bufs = [C.create_string_buffer(g.fill_capacity(self.k)) for g in gaps]
        ptrs = (C.c_char_p * max(len(gaps), 1))()
        for i, b in enumerate(bufs):
            ptrs[i] = C.cast(b, C.c_char_p)
        rc = lib().emul_fill_batch(self.h, self.k, arr, len(gaps), res, ptrs, cap0)
        assert rc == 0, rc
        return [res[i] for i in range(len(gaps))], [b.value.decode() for b in bufs]

    def fill_records(self, seqs, comments, *, dist_error=500, fuz=10, skip_confident=False, unique=False, all_paths=True):
        n = len(seqs)
        sa = (C.c_char_p * max(n, 1))(*[s.encode() for s in seqs])
        ca = (C.c_char_p * max(n, 1))(*[c.encode() for c in comments])
        opt = E.G2SRunOptions(self.k, dist_error, fuz, int(skip_confident), int(unique), int(all_paths))
        filled = (C.c_void_p * max(n, 1))()
        log = C.c_void_p()
        gc, fc = C.c_int32(0), C.c_int32(0)
        gr, ngr = C.c_void_p(), C.c_size_t(0)
        rc = lib().emul_fill_records(self.h, sa, ca, n, C.byref(opt), filled, C.byref(log), C.byref(gc), C.byref(fc),
                                     C.byref(gr), C.byref(ngr))
        assert rc == 0, rc
        out = [C.string_at(filled[i]).decode() for i in range(n)]
        logs = C.string_at(log.value).decode()
        arr = C.cast(gr, C.POINTER(E.G2SResult))
        results = [E.G2SResult.from_buffer_copy(arr[i]) for i in range(ngr.value)]
        for i in range(n):
            lib().emul_free(filled[i])
        lib().emul_free(log)
        lib().emul_free(gr)
        return out, logs, gc.value, fc.value, results

    def __del__(self):
        try:
            lib().emul_graph_free(self.h)
        except Exception:
            pass
