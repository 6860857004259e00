#!/usr/bin/env python
"""Debug / timing probe of the feature-mix forward (run on the GPU box): errors of the tensor-core and SIMT paths
against a float64 torch reference, and device time per call."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphflow_b200  # noqa: E402
from graphflow_b200 import _lib  # noqa: E402

ctx = graphflow_b200.Context(0)
torch.manual_seed(0)
SHAPES = [(256, 64, 64), (1024, 1152, 64), (2048 + 17, 1152, 64), (640, 72, 16), (128 * 150 + 5, 576, 32), (512 * 1024, 1152, 64)]
if len(sys.argv) > 1 and sys.argv[1] == "big":  # only the headline shape (what the ncu captures profile)
    SHAPES = SHAPES[-1:]
for (M, K, P) in SHAPES:
    X = torch.rand((M, K), device="cuda") * 2 - 1
    W = (torch.rand((K, P), device="cuda") * 2 - 1) * 0.2
    b = torch.rand((P,), device="cuda") - 0.5
    ref = (X.double() @ W.double())
    den = ref.abs().max().item()
    for name, path in (("tensor", _lib.MIX_TENSOR), ("simt", _lib.MIX_SIMT)):
        ctx.set_mix_path(path)
        Y, Z = ctx.mix_forward(X, W, b)
        torch.cuda.synchronize()
        err = (Y.double() - ref).abs().max().item() / den
        zref = torch.nn.functional.leaky_relu(ref + b.double(), 0.01)
        zerr = (Z.double() - zref).abs().max().item() / zref.abs().max().item()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        ev0.record()
        for _ in range(reps):
            ctx.mix_forward(X, W, b, want_Y=False)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / reps
        gbs = 4.0 * M * (K + P) / (ms * 1e-3) / 1e9
        tf = 2.0 * M * K * P / (ms * 1e-3) / 1e12
        print("M=%d K=%d P=%d %-6s errY=%.2e errZ=%.2e  %.3f ms  %.0f GB/s  %.1f TFLOP/s(useful fp32)" % (M, K, P, name, err, zerr, ms, gbs, tf), flush=True)
    # backward (grad-X on the tensor cores when the shape allows; grad-W / grad-bias on the SIMT kernels)
    gZ = torch.rand((M, P), device="cuda") * 2 - 1
    Yp = (X.double() @ W.double()).float()
    for name, path in (("tensor", _lib.MIX_AUTO), ("simt", _lib.MIX_SIMT)):
        ctx.set_mix_path(path)
        gX = torch.empty((M, K), device="cuda")
        ctx.set_kernel_timing(True)
        for i in range(4):
            if i == 1:
                ctx.set_kernel_timing(True)
            ctx.mix_backward(X, W, gZ, bias=b, Y=Yp, gX=gX)
        torch.cuda.synchronize()
        kt = ctx.kernel_timing()
        ctx.set_kernel_timing(False)
        gy = gZ.double() * torch.where(Yp.double() + b.double() > 0, 1.0, 0.01)
        gref = gy @ W.double().t()
        gerr = (gX.double() - gref).abs().max().item() / gref.abs().max().item()
        print("   backward %-6s errgX=%.2e  " % (name, gerr) + "  ".join("%s %.3f ms" % (k, v[0] / v[1]) for k, v in kt.items()), flush=True)
    ctx.set_mix_path(_lib.MIX_AUTO)
    if err != err:
        break
ctx.close()
