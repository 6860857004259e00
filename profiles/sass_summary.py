#!/usr/bin/env python
"""SASS evidence per kernel of graphflow_b200/libccn_b200.so (no GPU needed): instruction count and the mnemonics that show
which hardware paths a kernel uses -- UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st: tensor memory), UTMALDG (TMA tensor
load), UBLKCP (cp.async.bulk: 1-D TMA copy), LDGSTS (cp.async), RED / ATOMG (global reductions / atomics), SHFL, BAR.

    python profiles/sass_summary.py > profiles/r02_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "graphflow_b200", "libccn_b200.so")
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "LDGSTS", "RED", "ATOMG", "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "FFMA", "SYNCS"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k in ("RED", "ATOMG", "UTCHMMA", "UTMALDG", "UBLKCP") and op.startswith(k)):
                    counts[cur][k] += 1
    dm = demangle(list(counts))
    print("# SASS summary of libccn_b200.so (sm_100a), per kernel\n")
    print("Columns: SASS instructions, then occurrences of the mnemonics that prove a hardware path (`UTCHMMA` = tcgen05.mma, `LDTM`/`STTM` = "
          "tcgen05.ld/st, `UTMALDG` = TMA tensor load, `UBLKCP` = cp.async.bulk, `LDGSTS` = cp.async, `RED`/`ATOMG` = global reductions).\n")
    print("| kernel | instr | " + " | ".join(KEYS) + " |")
    print("|---|---:|" + "---:|" * len(KEYS))
    rows = []
    for k, c in counts.items():
        name = dm.get(k, k)
        name = re.sub(r"ccn::\(anonymous namespace\)::", "", name)
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"\(.*\)$", "", name)
        rows.append((name, c))
    for name, c in sorted(rows):
        print("| `%s` | %d | %s |" % (name[:70], c["total"], " | ".join(str(c[k]) if c[k] else "" for k in KEYS)))


if __name__ == "__main__":
    sys.exit(main())
