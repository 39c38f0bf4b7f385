"""Named parity cases shared by tests/golden/make_golden.py (which runs the unmodified reference
binary in the build container) and the tests (which re-create the same seeded inputs and compare
the oracle / the CUDA engine with the committed golden outputs).

Every case is (workload builder, Gap2Seq-core flags).  Inputs are regenerated from numpy seeds; only the
reference OUTPUTS are committed under tests/golden/.
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass, field

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from gap2seq_b200 import synth  # noqa: E402


@dataclass
class Case:
    name: str
    workload: str                 # synth.make_workload name
    scale: float
    flags: dict = field(default_factory=dict)   # k, fuz, dist_error, best_only, unique, all_upper
    inverted_copies: bool = False
    multigap: int = 0             # >0: one scaffold record holding this many gaps
    seed: int | None = None
    note: str = ""

    def opts(self) -> dict:
        o = dict(k=31, solid=2, fuz=10, dist_error=500, best_only=False, unique=False, all_upper=False, randseed=1)
        o.update(self.flags)
        return o


CASES = [
    Case("cfg1_small", "cfg1", 0.2),
    Case("cfg1_full", "cfg1", 1.0, note="BASELINE.json configs[0]"),
    Case("cfg1_bestonly", "cfg1", 0.5, flags=dict(best_only=True)),
    Case("cfg1_allupper", "cfg1", 0.5, flags=dict(all_upper=True)),
    Case("cfg1_fuz0", "cfg1", 0.2, flags=dict(fuz=0), note="Q5: nothing filled in all-paths mode"),
    Case("cfg1_fuz0_bestonly", "cfg1", 0.2, flags=dict(fuz=0, best_only=True)),
    Case("cfg1_fuz1", "cfg1", 0.2, flags=dict(fuz=1)),
    Case("cfg1_fuz3_bestonly", "cfg1", 0.2, flags=dict(fuz=3, best_only=True)),
    Case("cfg1_disterr20", "cfg1", 0.2, flags=dict(dist_error=20), note="Q7: dist-error < k fills nothing"),
    Case("cfg1_disterr101", "cfg1", 0.5, flags=dict(dist_error=101), note="odd gap_err (Q8)"),
    Case("cfg1_k21", "cfg1", 0.2, flags=dict(k=21)),
    Case("cfg1_multigap", "cfg1", 0.3, multigap=4, note="Q1: multi-gap record loses its prefix"),
    Case("cfg1_jitter", "cfg1", 0.5, seed=11, flags=dict(dist_error=200), note="gap length estimates off by up to 20 %"),
    Case("cfg2_small", "cfg2", 0.05, note="1 % substitution errors: bubbles and tips"),
    Case("cfg2_small_bestonly", "cfg2", 0.05, flags=dict(best_only=True)),
    Case("cfg2_small_unique", "cfg2", 0.05, flags=dict(unique=True)),
    Case("cfg2_small_solid3", "cfg2", 0.03, flags=dict(solid=3), seed=22),
    Case("cfg3_small", "cfg3", 0.012, note="repeat family + tandem arrays: non-trivial SCCs, saturated counts"),
    Case("cfg3_small_bestonly", "cfg3", 0.012, flags=dict(best_only=True)),
    Case("cfg3_seed33", "cfg3", 0.015, seed=33),
    Case("cfg3_inverted", "cfg3", 0.012, inverted_copies=True, note="Q3: order-dependent gaps are fenced"),
    Case("cfg4_small", "cfg4", 0.0006),
    Case("cfg5_small", "cfg5", 0.003, note="insertion genotyping as gap records (Q12)"),
]
CASE_BY_NAME = {c.name: c for c in CASES}


def build(case: Case) -> synth.Workload:
    w = synth.make_workload(case.workload, case.scale, inverted_copies=case.inverted_copies, seed=case.seed)
    o = case.opts()
    w.k, w.solid, w.fuz, w.dist_error, w.best_only = o["k"], o["solid"], o["fuz"], o["dist_error"], o["best_only"]
    if case.name == "cfg1_jitter":
        rng = np.random.default_rng(111)
        w.gaps = synth.place_gaps(w.genome, rng.integers(50, 401, size=8), rng, flank=100, est_jitter=0.2)
    if case.multigap:
        # one record: genome prefix with `multigap` N-runs, some closer than k+fuz to each other
        rng = np.random.default_rng(7)
        g = w.genome[:12000].copy()
        pos = [1000, 1000 + 300 + 35, 4000, 8000][: case.multigap]
        lens = [300, 120, 57, 200][: case.multigap]
        seq = bytearray(g.tobytes())
        for p, L in zip(pos, lens):
            seq[p:p + L] = b"N" * L
        w.gaps = [synth.GapRecord(name="multi scaffold", seq=seq.decode(), truth="", pos=0, gap_len=0)]
    return w


def ref_cmd_flags(case: Case) -> list[str]:
    o = case.opts()
    f = ["-k", str(o["k"]), "-solid", str(o["solid"]), "-fuz", str(o["fuz"]), "-dist-error", str(o["dist_error"]),
         "-randseed", str(o["randseed"]), "-nb-cores", "1"]
    if o["best_only"]:
        f.append("-best-only")
    if o["unique"]:
        f.append("-unique")
    if o["all_upper"]:
        f.append("-all-upper")
    return f


def golden_paths(name: str) -> tuple[str, str]:
    d = os.path.join(ROOT, "tests", "golden")
    return os.path.join(d, name + ".filled.fa"), os.path.join(d, name + ".log")


def filter_log(text: str) -> str:
    """Keep the lines the reference prints for the hot path (SURVEY 8c operating notes)."""
    keep = []
    for line in text.splitlines():
        if line.startswith("Filled scaffolds file"):
            continue
        if line.startswith(("Scaffold:", "SubgraphStats:", "Filled ", "Nodes:")):
            if line.startswith("Nodes:"):
                line = line.split(" Total memory")[0]
            keep.append(line)
    return "\n".join(keep) + "\n"
