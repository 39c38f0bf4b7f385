#!/usr/bin/env python
"""Compare a run of this engine with a run of the reference Gap2Seq-core on the same scaffold file, gap by gap.

    tools/compare_with_reference.py ours.log ours.fa reference.log reference.fa

The logs are the tools' stdout (only `Scaffold:` / `SubgraphStats:` / `Filled` lines are read), the FASTA files their
-filled outputs.  Gaps the REFERENCE gave up on because its DP tables outgrew -max-mem (`Memory limit exceeded`,
src/Gap2Seq.cpp:140-147; byte accounting src/Gap2Seq.cpp:454-476, result src/Gap2Seq.cpp:1525-1526) are LISTED SEPARATELY
and are not mismatches: the engine bounds its tables by device tiers instead of reproducing libstdc++'s allocation sizes,
so it fills (or gives up on) a different set.  Because such a gap may or may not have consumed rand() in the reference
(the limit can strike after a path was found and traced), records BEHIND the first such gap are compared too but a
difference there is reported in its own class (`after_reference_skip`).  Exit status 0 iff there is no true mismatch."""
from __future__ import annotations

import json
import re
import sys


def parse_log(text: str):
    """-> list of gaps in print order: dict(comment, scaffold_line, subgraph_line, mem_limit)"""
    gaps = []
    for ln in text.splitlines():
        if ln.startswith("Scaffold: "):
            m = re.match(r"Scaffold: (.*) GapStart: (\d+) GapEnd: (\d+) GapLength: (\d+) PathsFound: (-?\d+)", ln)
            gaps.append({"comment": m.group(1) if m else ln, "scaffold_line": ln, "subgraph_line": None,
                         "mem_limit": ln.rstrip().endswith("Memory limit exceeded")})
        elif ln.startswith("SubgraphStats:") and gaps:
            gaps[-1]["subgraph_line"] = ln
    return gaps


def parse_fasta(text: str):
    lines = text.splitlines()
    return [(lines[i][1:], lines[i + 1] if i + 1 < len(lines) else "") for i in range(0, len(lines), 2) if lines[i].startswith(">")]


def compare(our_log: str, our_fa: str, ref_log: str, ref_fa: str) -> dict:
    og, rg = parse_log(our_log), parse_log(ref_log)
    out = {"gaps": len(rg), "identical": 0, "reference_skipped_max_mem": [], "engine_mem_limit": [], "after_reference_skip": [],
           "mismatch": [], "records": 0, "records_identical": 0, "records_with_reference_skip": [], "records_after_reference_skip": [],
           "records_mismatch": []}
    if len(og) != len(rg):
        out["mismatch"].append({"gap": None, "why": f"gap counts differ: ours {len(og)}, reference {len(rg)}"})
    first_skip = None
    skipped_comments = set()
    for i, (a, b) in enumerate(zip(og, rg)):
        if b["mem_limit"]:
            out["reference_skipped_max_mem"].append({"gap": i, "comment": b["comment"], "ours": a["scaffold_line"]})
            skipped_comments.add(b["comment"])
            if first_skip is None:
                first_skip = i
            continue
        if a["mem_limit"]:
            out["engine_mem_limit"].append({"gap": i, "comment": a["comment"]})
        same = a["scaffold_line"] == b["scaffold_line"] and a["subgraph_line"] == b["subgraph_line"]
        if same:
            out["identical"] += 1
        elif first_skip is not None:
            out["after_reference_skip"].append({"gap": i, "comment": a["comment"]})
        else:
            out["mismatch"].append({"gap": i, "ours": a["scaffold_line"], "reference": b["scaffold_line"]})
    ofa, rfa = parse_fasta(our_fa), parse_fasta(ref_fa)
    out["records"] = len(rfa)
    if len(ofa) != len(rfa):
        out["records_mismatch"].append({"record": None, "why": "record counts differ"})
    seen_skip = False
    for (oc, os_), (rc, rs) in zip(ofa, rfa):
        if rc in skipped_comments:
            out["records_with_reference_skip"].append(rc)
            seen_skip = True
        elif (oc, os_) == (rc, rs):
            out["records_identical"] += 1
        elif seen_skip:
            out["records_after_reference_skip"].append(rc)
        else:
            out["records_mismatch"].append({"record": rc})
    return out


def main():
    if len(sys.argv) != 5:
        sys.exit(__doc__)
    texts = [open(p).read() for p in sys.argv[1:5]]
    res = compare(*texts)
    print(json.dumps(res, indent=1))
    sys.exit(1 if res["mismatch"] or res["records_mismatch"] else 0)


if __name__ == "__main__":
    main()
