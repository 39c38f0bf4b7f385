"""GapCutter / GapMerger of this package (gap2seq_b200/csrc/g2s_gapcutter.cpp, g2s_gapmerger.cpp)
against the reference binaries built under oracle/_ref/bin (TEST INFRASTRUCTURE): same files byte for byte, for scaffold
sets that hit every branch of the cutter (gaps with room on both sides, a last gap, gaps sharing a flank with and without
splitting, runs of gaps around sequence too short for a flank with and without -mask, N at the start / end, lower-case n,
multi-line FASTA, records without gaps)."""
import os
import subprocess

import numpy as np
import pytest

import cases
import oracle_lib
from gap2seq_b200 import synth

LIB = os.path.join(cases.ROOT, "gap2seq_b200", "lib")
REF = os.path.join(cases.ROOT, "oracle", "_ref", "bin")
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "GapCutter")) or not os.path.exists(os.path.join(LIB, "GapCutter")),
                               reason="reference or package binaries not built (run __graft_entry__.build())")


def make_scaffolds(path, seed, n_scaffolds=6):
    rng = np.random.default_rng(seed)
    genome = synth.random_genome(40000, rng).tobytes().decode()
    out = []
    for s in range(n_scaffolds):
        a = int(rng.integers(0, len(genome) - 6000))
        seq = list(genome[a:a + int(rng.integers(300, 5000))])
        for _ in range(int(rng.integers(0, 7))):
            L = int(rng.choice([1, 5, 20, 31, 60, 200]))
            p = int(rng.integers(0, max(1, len(seq) - L)))
            seq[p:p + L] = ("n" if rng.random() < 0.3 else "N") * min(L, len(seq) - p)
            if rng.random() < 0.5:   # a second gap close by: shared flanks, short islands
                q = p + L + int(rng.choice([3, 20, 31, 40, 62, 70]))
                L2 = int(rng.choice([2, 15, 80]))
                if q + L2 < len(seq):
                    seq[q:q + L2] = "N" * L2
        if s == 1:
            seq[:12] = "N" * 12
        if s == 2:
            seq[-9:] = "n" * 9
        out.append((f"scaf{s} some description {s}", "".join(seq)))
    with open(path, "w") as f:
        for i, (name, seq) in enumerate(out):
            f.write(">" + name + "\n")
            if i % 2:   # multi-line record
                for j in range(0, len(seq), 70):
                    f.write(seq[j:j + 70] + "\n")
            else:
                f.write(seq + "\n")
    return out


def run(tool_dir, tool, args, cwd):
    p = subprocess.run([os.path.join(tool_dir, tool)] + args, cwd=cwd, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    return p.stdout


def read(path):
    return open(path).read() if os.path.exists(path) else None


@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("flags", [[], ["-mask"], ["-no-split"], ["-mask", "-no-split", "-k", "21", "-fuz", "3"], ["-fuz", "0"]])
def test_cutter_and_merger_equal_the_reference_tools(seed, flags, tmp_path):
    d = str(tmp_path)
    scaffolds = make_scaffolds(os.path.join(d, "scaf.fa"), seed)
    outs = {}
    for tag, tool_dir in (("ref", REF), ("b200", LIB)):
        log = run(tool_dir, "GapCutter", ["-scaffolds", "scaf.fa", "-contigs", f"c_{tag}.fa", "-gaps", f"g_{tag}.fa", "-bed",
                                          f"b_{tag}.bed"] + flags, d)
        outs[tag] = (read(os.path.join(d, f"c_{tag}.fa")), read(os.path.join(d, f"g_{tag}.fa")), read(os.path.join(d, f"b_{tag}.bed")),
                     log.replace(f"_{tag}.", "_X."))
    assert outs["ref"] == outs["b200"]
    if outs["ref"][1] is None:
        return   # nothing was cut
    merged = {}
    for tag, tool_dir in (("ref", REF), ("b200", LIB)):
        log = run(tool_dir, "GapMerger", ["-scaffolds", f"m_{tag}.fa", "-contigs", "c_ref.fa", "-gaps", "g_ref.fa"], d)
        merged[tag] = (read(os.path.join(d, f"m_{tag}.fa")), log.replace(f"m_{tag}.fa", "M"))
    assert merged["ref"] == merged["b200"]
    if "-mask" not in flags:   # cutting and merging unfilled gaps gives the scaffolds back
        want = "".join(f">{n}\n{s}\n" for n, s in scaffolds)
        assert merged["b200"][0] == want


def test_cut_fill_merge_restores_an_error_free_genome(case_inputs, tmp_path):
    """The whole flow of the reference (GapCutter -> gap filling -> GapMerger) with this package's tools and the oracle's
    gap driver in the middle: on error-free reads every cut gap is filled with the removed bases."""
    if not os.path.exists(os.path.join(LIB, "GapCutter")):
        pytest.skip("package binaries not built")
    case, w, g = case_inputs("cfg1_small")
    d = str(tmp_path)
    genome = w.genome.tobytes().decode()
    seq = list(genome[:9000])
    for p, L in [(800, 120), (2500, 60), (2500 + 60 + 100, 40), (5000, 300), (7000, 90)]:
        seq[p:p + L] = "N" * L
    with open(os.path.join(d, "scaf.fa"), "w") as f:
        f.write(">chr1 test\n" + "".join(seq) + "\n")
    run(LIB, "GapCutter", ["-scaffolds", "scaf.fa", "-contigs", "c.fa", "-gaps", "g.fa", "-bed", "b.bed"], d)
    names, seqs = [], []
    for ln in open(os.path.join(d, "g.fa")).read().splitlines():
        (names if ln.startswith(">") else seqs).append(ln[1:] if ln.startswith(">") else ln)
    oracle_lib.srand(1)
    with open(os.path.join(d, "filled.fa"), "w") as f:
        for n, s in zip(names, seqs):
            filled, _, gc, fc = g.fill_record(s, n, d_err=200, fuz=10)
            assert gc == fc == 1
            f.write(f">{n}\n{filled}\n")
    run(LIB, "GapMerger", ["-scaffolds", "m.fa", "-contigs", "c.fa", "-gaps", "filled.fa"], d)
    assert open(os.path.join(d, "m.fa")).read().upper() == ">CHR1 TEST\n" + genome[:9000] + "\n"
