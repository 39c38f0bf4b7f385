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


def test_struct_sizes_match_header(tmp_path):
    """sizeof of every structure that crosses the C ABI, as gcc sees include/g2s.h, against the ctypes mirror."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "g2s.h"\nint main(void) { printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(g2s_gap), '
                   'sizeof(g2s_result), sizeof(g2s_config), sizeof(g2s_timing), sizeof(g2s_run_options), sizeof(g2s_subgraph_stats)); return 0; }\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    mirror = [engine.G2SGap, engine.G2SResult, engine.G2SConfig, engine.G2STiming, engine.G2SRunOptions, engine.G2SSubgraphStats]
    assert sizes == [C.sizeof(m) for m in mirror]
    assert C.sizeof(engine.G2SGap) == 56 and C.sizeof(engine.G2SConfig) == 32


def test_record_gaps_matches_reference_flank_rule():
    from gap2seq_b200.scaffold import record_gaps
    seq = "A" * 100 + "N" * 50 + "C" * 35
    (g,) = record_gaps(seq, k=31, fuz=10)
    assert (g.left_max_fuz, g.right_max_fuz, g.gap_len) == (10, 4, 50)
    assert g.left == "A" * 41 and g.right == "C" * 35
    assert record_gaps("A" * 30 + "N" * 5 + "C" * 60, k=31, fuz=10) == []
