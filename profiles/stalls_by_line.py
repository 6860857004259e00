#!/usr/bin/env python
"""Attribute the warp-stall samples of an `ncu --set full --import-source on` report to CUDA source lines.

    python profiles/stalls_by_line.py gpurun_out/X.ncu-rep <kernel regex> <mangled-name substring> [min share %]

ncu's CSV export of the source page is SASS-only, so the SASS rows (in address order) are zipped with
`nvdisasm -g` of the same function from graphflow_b200/libccn_b200.so (built with -lineinfo), whose `//## File ...
line N` markers give the source line of every instruction.  Needs ncu, cuobjdump and nvdisasm; no GPU.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "graphflow_b200", "libccn_b200.so")


def sass_rows(rep, kernel_regex):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel_regex],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    tables, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            tables.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    return tables[0]


def line_map(func_substr):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
    lines = []
    for fn in sorted(os.listdir(tmp)):
        if not fn.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, fn)], capture_output=True, text=True).stdout
        inside, cur_line = False, None
        for ln in txt.splitlines():
            if ln.startswith(".text."):
                inside = func_substr in ln
                cur_line = None
                if inside and lines:
                    return lines  # first match only
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.search(r"/\*[0-9a-f]{4,}\*/", ln):
                lines.append(cur_line)
        if lines:
            return lines
    return lines


def main():
    rep, kre, fsub = sys.argv[1:4]
    min_share = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    t = sass_rows(rep, kre)
    hdr = t["hdr"]
    i_samp = hdr.index("Warp Stall Sampling (All Samples)")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    lm = line_map(fsub)
    if len(lm) != len(t["rows"]):
        print("# warning: %d SASS rows in the report vs %d in nvdisasm; attribution may drift" % (len(t["rows"]), len(lm)))
    per = defaultdict(lambda: [0.0, defaultdict(float), 0])
    total = 0.0
    for k, r in enumerate(t["rows"]):
        s = float(r[i_samp] or 0)
        total += s
        key = lm[k] if k < len(lm) and lm[k] else ("?", 0)
        per[key][0] += s
        per[key][2] += 1
        for i, h in stall_cols:
            per[key][1][h] += float(r[i] or 0)
    src = {}
    print("kernel: %s\ntotal samples: %d\n" % (t["name"], total))
    print("| line | share | SASS instr | top stall reasons | source |")
    print("|---|---:|---:|---|---|")
    for key, (s, reasons, cnt) in sorted(per.items(), key=lambda kv: -kv[1][0]):
        if 100 * s / total < min_share:
            continue
        fn, ln = key
        if fn not in src:
            p = os.path.join(ROOT, "graphflow_b200", "csrc", fn)
            src[fn] = open(p).read().splitlines() if os.path.exists(p) else []
        text = src[fn][ln - 1].strip() if 0 < ln <= len(src[fn]) else ""
        top = sorted(reasons.items(), key=lambda kv: -kv[1])[:3]
        tops = ", ".join("%s %.0f%%" % (h.replace("stall_", ""), 100 * v / max(s, 1)) for h, v in top if v > 0)
        print("| %s:%d | %.1f%% | %d | %s | `%s` |" % (fn, ln, 100 * s / total, cnt, tops, text[:90].replace("|", "/")))


if __name__ == "__main__":
    main()
