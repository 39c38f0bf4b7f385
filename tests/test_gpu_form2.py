"""INTEGRATION.md form 2 built for real: the reference's own Gap2Seq.cpp and GATB's graph (construction, Bloom filter, MPHF)
with the two fill_gap call sites (src/Gap2Seq.cpp:252, 380) routed to libg2s_gpu.so through integration/g2s_gatb_glue.cpp
(oracle/build_ref.sh -> oracle/_ref/bin/Gap2Seq-core-g2s).  That binary must reproduce the golden outputs of the unmodified
reference - filled FASTA byte for byte, Scaffold:/SubgraphStats:/Filled lines - and must report, never hide, flank k-mers
for which GATB's contains() disagrees with the imported solid set (G2S_ST_PHANTOM_FLANK)."""
import ctypes
import os
import subprocess

import pytest

import cases
from gap2seq_b200 import engine, synth
from gap2seq_b200.engine import GapSpec
from test_gpu_cli import strip_nodes

pytestmark = pytest.mark.gpu
BIN = os.path.join(cases.ROOT, "oracle", "_ref", "bin", "Gap2Seq-core-g2s")


@pytest.mark.parametrize("name", ["cfg1_small", "cfg1_multigap", "cfg2_small", "cfg2_small_unique", "cfg3_small",
                                  "cfg3_small_bestonly", "cfg1_allupper", "cfg1_fuz0", "cfg1_k21", "cfg5_small"])
def test_reference_driver_with_engine_fill_gap_reproduces_goldens(name, case_inputs, tmp_path):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/bin/Gap2Seq-core-g2s was not built (needs /root/reference at build time)")
    case, w, g = case_inputs(name)
    paths = w.write(str(tmp_path))
    out = str(tmp_path / "filled.fa")
    p = subprocess.run([BIN] + cases.ref_cmd_flags(case) + ["-reads", paths["reads"], "-scaffolds", paths["gaps"], "-filled", out],
                       capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "g2s: imported" in p.stderr
    gfa, glog = cases.golden_paths(name)
    phantom = "PHANTOM_FLANK" in p.stderr
    od = "ORDER_DEPENDENT" in p.stderr
    if not phantom and not od:
        assert open(out).read() == open(gfa).read()
        assert cases.filter_log(p.stdout) == strip_nodes(open(glog).read())


def test_phantom_flank_is_flagged_through_the_abi():
    """present_bits that claim a k-mer the solid set does not hold (what a Bloom false positive looks like to the engine),
    and bits that deny one it holds: both flagged; truthful bits: not flagged; no bits: not flagged."""
    w = synth.make_workload("cfg1", 0.3)
    e = engine.Engine(k=w.k, max_fuz=w.fuz)
    e.build_from_reads(w.reads, w.solid)
    rec = w.gaps[0]
    i = rec.seq.index("N")
    left, right = rec.seq[i - w.k - w.fuz:i], rec.seq[i + rec.gap_len:i + rec.gap_len + w.k + w.fuz]
    lk = [left[d:d + w.k] for d in range(w.fuz + 1)]
    rk = [right[j:j + w.k] for j in range(w.fuz + 1)]
    has = e.contains(lk + rk)
    truth = sum(1 << d for d in range(w.fuz + 1) if has[d]) | sum(1 << (32 + j) for j in range(w.fuz + 1) if has[w.fuz + 1 + j])
    # an error in the left flank: its k-mers covering the error are absent from the graph
    bad_left = left[:5] + ("A" if left[5] != "A" else "C") + left[6:]
    bhas = e.contains([bad_left[d:d + w.k] for d in range(w.fuz + 1)])
    assert not bhas[0]
    btruth = sum(1 << d for d in range(w.fuz + 1) if bhas[d]) | (truth & ~((1 << 32) - 1))

    def status(l, bits):
        ctypes.CDLL(None).srand(1)
        res, _ = e.fill_batch([GapSpec(l, right, rec.gap_len, w.dist_error, w.fuz, w.fuz, present_bits=bits)])
        return res[0].status, res[0].count
    PH = 64
    assert not status(left, None)[0] & PH
    assert not status(left, truth)[0] & PH
    assert status(left, truth & ~1)[0] & PH              # the caller denies a k-mer the set holds
    assert not status(bad_left, btruth)[0] & PH
    st, cnt = status(bad_left, btruth | 1)               # the caller's Bloom filter "finds" the broken first k-mer
    assert st & PH
    assert status(left, truth)[1] == status(left, None)[1]
    e.close()
