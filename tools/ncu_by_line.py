#!/usr/bin/env python
"""Per-source-line view of an `ncu --set full --import-source on` capture of one of the fill kernels.

The kernels inline everything from g2s_fill.cuh, and ncu's CSV source page lists SASS only; this joins that page with
the line table of the cubin (nvdisasm -g) by instruction order and sums instructions executed and stall samples per
source line.   usage: tools/ncu_by_line.py <report.ncu-rep> <mangled kernel substring> [top N]"""
import csv
import re
import subprocess
import sys
import tempfile
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("G2S_LIB") or os.path.join(ROOT, "gap2seq_b200", "lib", "libg2s_gpu.so")], cwd=tmp,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, os.environ.get("G2S_CUBIN", "g2s_fill") + ".sm_100a.cubin")], capture_output=True, text=True).stdout
    cur, line, inside, res = None, None, False, []
    for ln in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            inside = kernel_sub in m.group(1)
            continue
        if ln.lstrip().startswith(".section"):
            inside = False
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            # keep the innermost location (nvdisasm prints "inlined at" chains after the first)
            if "inlined at" not in ln:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            res.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return res


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    sel = ["--kernel-name", "regex:" + os.environ["G2S_NCU_KERNEL"]] if os.environ.get("G2S_NCU_KERNEL") else []
    page = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
    rows = list(csv.reader(page.splitlines()))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    body = rows[2:]
    sass = sass_lines(ksub)
    if len(sass) != len(body):
        print(f"warning: {len(sass)} SASS instructions in the cubin, {len(body)} in the report", file=sys.stderr)
    agg = {}
    tot_i = tot_s = 0
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for (addr, text, loc), r in zip(sass, body):
        inst = int(r[col["Instructions Executed"]] or 0)
        samp = int(r[col["# Samples"]] or 0)
        a = agg.setdefault(loc, {"inst": 0, "samp": 0, "st": {}})
        a["inst"] += inst
        a["samp"] += samp
        for h in stall_cols:
            v = int(r[col[h]] or 0)
            if v:
                a["st"][h] = a["st"].get(h, 0) + v
        tot_i += inst
        tot_s += samp
    src = {}
    print(f"total warp instructions {tot_i}, stall samples {tot_s}")
    for key in ("samp", "inst"):
        print(f"--- top lines by {key}")
        for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][key])[:top]:
            if loc is None:
                continue
            f, ln = loc
            if f not in src:
                try:
                    src[f] = open(os.path.join(ROOT, "gap2seq_b200", "csrc", f)).read().splitlines()
                except OSError:
                    src[f] = []
            code = src[f][ln - 1].strip()[:90] if 0 < ln <= len(src[f]) else ""
            st = ", ".join(f"{k[6:]} {v}" for k, v in sorted(a["st"].items(), key=lambda kv: -kv[1])[:3])
            print(f"{f}:{ln:5d} inst {100 * a['inst'] / max(tot_i, 1):5.1f}%  samp {100 * a['samp'] / max(tot_s, 1):5.1f}%  [{st}]  {code}")


if __name__ == "__main__":
    main()
