#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call brought back into the small text summaries committed under profiles/.

    python profiles/summarize_ncu.py launches gpurun_out/X_launches.csv  > profiles/X_launches.md
    python profiles/summarize_ncu.py full     gpurun_out/X_prof.ncu-rep  > profiles/X_full.md

`launches` reads the CSV written by  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...
and prints per-kernel launch counts, total device time and the share of all time spent in this repo's kernels.
`full` reads an `ncu --set full` report (needs the ncu binary, no GPU) and prints the metrics the roofline uses.
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

FULL_METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers",
    "launch__grid_size",
    "launch__block_size",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("ccn::<unnamed>::", "").replace("ccn::", "")
    return name.strip()


def launches(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
            rows.append((short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], ns))
    agg = OrderedDict()
    for name, grid, block, ns in rows:
        a = agg.setdefault(name, [0, 0.0, set()])
        a[0] += 1
        a[1] += ns
        a[2].add(grid)
    total = sum(a[1] for a in agg.values())
    ours = sum(a[1] for n, a in agg.items() if n.startswith("k_"))
    print("| kernel | launches | total us | share of all | share of ours | grids |")
    print("|---|---:|---:|---:|---:|---|")
    for name, (cnt, ns, grids) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        own = "%.1f%%" % (100 * ns / ours) if name.startswith("k_") and ours else "-"
        g = ", ".join(sorted(grids)[:4]) + (" ..." if len(grids) > 4 else "")
        print("| `%s` | %d | %.1f | %.1f%% | %s | %s |" % (name[:70], cnt, ns / 1e3, 100 * ns / total, own, g))
    print("\n%d launches, %.1f us total device time, %.1f us (%.1f%%) in this repo's kernels (`k_*`).  Per-launch times "
          "are cold-cache and serialised by the profiler: compare shares, not absolutes." %
          (len(rows), total / 1e3, ours / 1e3, 100 * ours / max(total, 1)))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    cols = [(m, hdr.index(m)) for m in FULL_METRICS if m in hdr]
    k = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("### `%s`  grid %s block %s\n" % (short(r[k]), r[hdr.index("launch__grid_size")], r[hdr.index("launch__block_size")]))
        print("| metric | value | unit |\n|---|---:|---|")
        for m, i in cols:
            print("| %s | %s | %s |" % (m, r[i], units[i]))
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", ""))
            wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", ""))
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd *= scale.get(units[hdr.index("dram__bytes_read.sum")], 1)
            wr *= scale.get(units[hdr.index("dram__bytes_write.sum")], 1)
            t = float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))
            t *= {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(units[hdr.index("gpu__time_duration.sum")], 1e-9)
            print("| **traffic = dram read + write** | %.0f | byte |" % (rd + wr))
            print("| **traffic / duration (under ncu)** | %.1f | GB/s |" % ((rd + wr) / t / 1e9))
        except (ValueError, ZeroDivisionError):
            pass
        print()


if __name__ == "__main__":
    if len(sys.argv) != 3 or sys.argv[1] not in ("launches", "full"):
        sys.exit(__doc__)
    (launches if sys.argv[1] == "launches" else full)(sys.argv[2])
