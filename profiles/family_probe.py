#!/usr/bin/env python
"""Timing probe of the other contraction-family members through `ccn_contract_family_*` on one B200 at the headline
shape (N=32, C=64): RisiContraction_4, RisiContraction_10, RisiContraction_18_dropout (9 of 18 slabs kept) and, for
comparison, the same 18 slabs through the fused 18-way kernels.  Roofline: HBM, algorithmic bytes per forward+backward
instance 8 (N^3 C + S N^2 C + N^2) with S slabs (SURVEY.md section 8d with 18 -> S).
    python profiles/family_probe.py [batch] > gpurun_out/rXX_family.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphflow_b200  # noqa: E402

N, C = 32, 64
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = graphflow_b200.Context(0)
g = torch.Generator(device="cuda").manual_seed(1)
T = torch.rand((B, N, N, N, C), device="cuda", generator=g) * 2 - 1
adj = (torch.rand((B, N, N), device="cuda", generator=g) < 0.06).float()
adj = ((adj + adj.transpose(1, 2) + torch.eye(N, device="cuda")) > 0).float()
gT = torch.empty_like(T)
peak_file = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
peak = json.load(open(peak_file))["hbm_gbs"] if os.path.exists(peak_file) else 6650.0
use9 = [k % 2 == 0 for k in range(18)]
cases = {
    "RisiContraction_4": (4, None, lambda o: ctx.contract_family_forward(4, T, out=o), lambda go: ctx.contract_family_backward(4, go, gT=gT)),
    "RisiContraction_10": (10, None, lambda o: ctx.contract_family_forward(10, T, adj, out=o),
                           lambda go: ctx.contract_family_backward(10, go, adj, gT=gT)),
    "RisiContraction_18_dropout(9 kept)": (18, use9, lambda o: ctx.contract_family_forward(18, T, adj, keep_mask=use9, out=o),
                                           lambda go: ctx.contract_family_backward(18, go, adj, keep_mask=use9, gT=gT)),
    "RisiContraction_18 (fused kernels, reference point)": (18, None, lambda o: ctx.contract18_forward(T, adj, out=o),
                                                            lambda go: ctx.contract18_backward(go, adj, gT=gT)),
}
res = {}
for name, (S, _, fwd, bwd) in cases.items():
    out = torch.empty((B, N, N, S * C), device="cuda")
    gout = torch.rand((B, N, N, S * C), device="cuda", generator=g) * 2 - 1
    for _ in range(2):
        fwd(out)
        bwd(gout)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 5
    e0.record()
    for _ in range(steps):
        fwd(out)
        bwd(gout)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ctx.set_kernel_timing(True)  # a separate pass: per-kernel events serialise the launches
    fwd(out)
    bwd(gout)
    kt = {k: round(v[0], 4) for k, v in ctx.kernel_timing().items()}
    ctx.set_kernel_timing(False)
    bytes_inst = 8 * (N ** 3 * C + S * N * N * C + N * N)
    rate = B / (ms * 1e-3)
    res[name] = {"slabs": S, "ms_per_step": ms, "contractions_per_s": rate, "algorithmic_bytes_per_instance": bytes_inst,
                 "achieved_gbs": rate * bytes_inst / 1e9, "roofline_frac": rate * bytes_inst / 1e9 / peak, "kernels_ms": kt}
    del out, gout
print(json.dumps({"workload": "contraction family fwd+bwd, N=%d C=%d batch %d" % (N, C, B), "hbm_peak_gbs": peak, "variants": res}))
