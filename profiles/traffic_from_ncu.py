#!/usr/bin/env python
"""profiles/traffic.json and a per-kernel table from an `ncu --set full` capture of profiles/ncu_target.py (128 instances per
launch): DRAM bytes per instance of every kernel of the op-level step and of one fused-promotion level.

    python profiles/traffic_from_ncu.py gpurun_out/X.ncu-rep profiles/X_level_traffic.md
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INST = 128
N, C = 32, 64


def main(rep, md):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3}
    recs = []
    for r in rows[2:]:
        name = re.sub(r"\(ccn::.*", "", r[col["Kernel Name"]])
        name = re.sub(r"^.*?(k_[a-z0-9_]+)", r"\1", name).replace("(int)", "").replace("(bool)", "")
        rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * scale.get(units[col["dram__bytes_read.sum"]], 1)
        wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * scale.get(units[col["dram__bytes_write.sum"]], 1)
        t = float(r[col["gpu__time_duration.sum"]].replace(",", "")) * tscale.get(units[col["gpu__time_duration.sum"]], 1e-9)
        l2 = float(r[col["lts__t_bytes.sum"]].replace(",", "")) * scale.get(units[col["lts__t_bytes.sum"]], 1) if "lts__t_bytes.sum" in col else 0
        warps = r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]] if "sm__warps_active.avg.pct_of_peak_sustained_active" in col else ""
        tens = r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]] if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in col else ""
        recs.append({"kernel": name.strip(), "read": rd, "write": wr, "t": t, "l2": l2, "warps": warps, "tensor": tens,
                     "regs": r[col["launch__registers_per_thread"]], "grid": r[col["launch__grid_size"]], "block": r[col["launch__block_size"]]})
    alg = 4 * (N ** 3 * C + N * N + 18 * N * N * C)
    traffic = {"_comment": "DRAM traffic per contraction instance (N=32, C=64) from the ncu --set full capture of profiles/ncu_target.py "
                           "(128 instances per launch): dram__bytes_read.sum + dram__bytes_write.sum, divided by 128. bench.py multiplies by "
                           "the instances per launch.", "source": os.path.relpath(md, ROOT), "N": N, "C": C}
    lines = ["# DRAM traffic per instance, round 2 (ncu --set full of profiles/ncu_target.py; 128 instances per launch, N=32, C=64)\n",
             "| kernel | grid x block | regs | duration under ncu | DRAM read / inst | DRAM write / inst | warps active | tensor pipe |",
             "|---|---|---:|---:|---:|---:|---:|---:|"]
    level = 0.0
    for x in recs:
        lines.append("| `%s` | %s x %s | %s | %.1f us | %.2f MB | %.2f MB | %.1f %% | %.1f %% |" % (
            x["kernel"][:60], x["grid"], x["block"], x["regs"], x["t"] * 1e6, x["read"] / INST / 1e6, x["write"] / INST / 1e6,
            float(x["warps"] or 0), float(x["tensor"] or 0)))
        k = x["kernel"]
        per = (x["read"] + x["write"]) / INST
        if k.startswith("k_fwd_fused<64, 0, 0>"):
            traffic["fwd_fused"] = {"bytes_per_instance": round(per), "algorithmic_bytes_per_instance": alg}
        elif k.startswith("k_bwd_fused<64, 0, 0, 0>"):
            traffic["bwd_fused"] = {"bytes_per_instance": round(per), "algorithmic_bytes_per_instance": alg}
        else:
            level += per
    lines.append("\nOne fused-promotion LEVEL forward + backward (every kernel above except the two op-level ones): **%.2f MB of DRAM traffic per "
                 "instance**; the stacked T (8.39 MB) and gT (8.39 MB) never touch DRAM, Z + gZ + f + gf are 4 x 0.26 MB; the rest is the "
                 "X / gX round trips (5 x 4.72 MB).  `k_fwd_fused<64, 1, 0>` = forward with the promotion gathered in, "
                 "`k_bwd_fused<64, 0, 1, 0>` = backward scattering into gf." % (level / 1e6))
    traffic["level_bytes_per_instance"] = round(level)
    open(md, "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
