"""Host-side helpers around scaffold records (pure Python, no compute).

``record_gaps`` restates, for callers that want explicit ``GapSpec`` lists (bench.py's HBM-resident timing, the
per-gap tests), how ``Gap2Seq::execute`` derives the arguments of ``fill_gap`` from a record
(reference src/Gap2Seq.cpp:340-380): left_max_fuz = min(i + k - prevGapEnd, fuz), the left flank starts at
i - k - left_max_fuz, right_max_fuz = min(len - (i + k), fuz), and a gap is only attempted when both flanks
are complete.  Only records whose gaps do not interact are supported here (every gap's left flank must lie after
the previous gap's right flank); the general case lives in the C++ driver (csrc/g2s_scaffold.h).
"""
from __future__ import annotations

from .engine import GapSpec


def record_gaps(seq: str, k: int = 31, fuz: int = 10, dist_error: int = 500, skip_confident: bool = False,
                all_paths: bool = True) -> list[GapSpec]:
    out: list[GapSpec] = []
    n = len(seq)
    i = 0
    prev = 0
    while i < n:
        if seq[i] not in "Nn":
            i += 1
            continue
        lmf = min(i + k - prev, fuz)
        ks = i - k - lmf
        gs = i
        while i < n and seq[i] in "Nn":
            i += 1
        gap = i - gs
        rmf = min(n - (i + k), fuz)
        ok = rmf >= 0 and i + k + rmf <= n and not any(c in "Nn" for c in seq[i:i + k + rmf])
        if ks >= prev and ok:
            if out and ks < prev + fuz:
                raise ValueError("interacting gaps: use Engine.fill_records")
            out.append(GapSpec(seq[ks:ks + k + lmf], seq[i:i + k + rmf], gap, dist_error, lmf, rmf, skip_confident,
                               all_paths))
        prev = i
    return out
