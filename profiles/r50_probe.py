#!/usr/bin/env python
"""Timing probe of StackTensor3D + RisiContraction_50 at BASELINE.json config 5 (N=48, C=128, batch 256) on one B200:
contractions/s, per-kernel device time and the fraction of the HBM roofline (algorithmic bytes
8 (N^3 C + 50 N^2 C + N^2) per forward+backward instance).  Run on the GPU box:
    python profiles/r50_probe.py [batch]  > gpurun_out/rXX_r50.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphflow_b200  # noqa: E402

N, C = 48, 128
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = graphflow_b200.Context(0)
g = torch.Generator(device="cuda").manual_seed(1)
T = torch.rand((B, N, N, N, C), device="cuda", generator=g) * 2 - 1
adj = (torch.rand((B, N, N), device="cuda", generator=g) < 0.08).float()
adj = ((adj + adj.transpose(1, 2) + torch.eye(N, device="cuda")) > 0).float()
out = torch.empty((B, N, N, 50 * C), device="cuda")
gout = torch.rand((B, N, N, 50 * C), device="cuda", generator=g) * 2 - 1
gT = torch.empty_like(T)
for _ in range(2):
    ctx.contract50_forward(T, adj, out=out)
    ctx.contract50_backward(gout, adj, gT=gT)
torch.cuda.synchronize()
ctx.set_kernel_timing(True)
steps = 3
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(steps):
    ctx.contract50_forward(T, adj, out=out)
    ctx.contract50_backward(gout, adj, gT=gT)
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / steps
kt = {k: {"ms_per_step": v[0] / steps, "launches_per_step": v[1] / steps} for k, v in ctx.kernel_timing().items()}
bytes_inst = 8 * (N ** 3 * C + 50 * N * N * C + N * N)
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
rate = B / (ms * 1e-3)
print(json.dumps({"workload": "RisiContraction_50 fwd+bwd N=48 C=128 batch %d" % B, "ms_per_step": ms, "contractions_per_s": rate,
                  "algorithmic_bytes_per_instance": bytes_inst, "achieved_gbs": rate * bytes_inst / 1e9, "hbm_peak_gbs": peak,
                  "roofline_frac": rate * bytes_inst / 1e9 / peak, "roofline_contractions_per_s": peak * 1e9 / bytes_inst,
                  "kernels": kt}))
