"""CPU-side checks of the boundary: the C ABI library loads, exports every symbol include/g2s.h declares, and
fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

from gap2seq_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "g2s.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(g2s_[a-z_0-9]+)\s*\(", txt)))


def test_header_symbols_are_all_bound_and_exported():
    decl = _declared_symbols()
    assert sorted(engine.ABI_SYMBOLS) == decl
    lib = engine.load_library()
    for s in decl:
        assert hasattr(lib, s), s
    assert lib.g2s_abi_version() == engine.ABI_VERSION == 4


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(engine.G2SError):
        engine.Engine(k=31)


def test_struct_sizes_match_header():
    assert C.sizeof(engine.G2SGap) == 48
    assert C.sizeof(engine.G2SResult) == 48 + 48 + 24
    assert C.sizeof(engine.G2SConfig) == 32


def test_record_gaps_matches_reference_flank_rule():
    from gap2seq_b200.scaffold import record_gaps
    seq = "A" * 100 + "N" * 50 + "C" * 35
    (g,) = record_gaps(seq, k=31, fuz=10)
    assert (g.left_max_fuz, g.right_max_fuz, g.gap_len) == (10, 4, 50)
    assert g.left == "A" * 41 and g.right == "C" * 35
    assert record_gaps("A" * 30 + "N" * 5 + "C" * 60, k=31, fuz=10) == []
