"""The kernels of the op-level headline and of one fused-promotion level, once each after a warm-up, at 128 instances
(4 graphs x 32 vertices, N = 32, C = 64 -> 64) -- the target of the round's `ncu --set full` capture:

    ncu --set full --clock-control none --import-source on -k regex:'k_(fwd|bwd)_fused|k_mix' -s <warm-up launches> -o gpurun_out/X python profiles/ncu_target.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import graphflow_b200  # noqa: E402
from bench import level_workload, molecular_adjacency  # noqa: E402

n, C, Co, G = 32, 64, 64, 4
dev = torch.device("cuda", 0)
ctx = graphflow_b200.Context(0)
w = level_workload(G, n, C, 3)
B = w["instances"]
d = lambda x: torch.from_numpy(x).to(dev)  # noqa: E731
rng = np.random.default_rng(0)
T = torch.rand((B, n, n, n, C), device=dev) * 2 - 1
gout = torch.rand((B, n, n, 18 * C), device=dev) * 2 - 1
adj = torch.from_numpy(np.stack([molecular_adjacency(n, rng) for _ in range(B)])).to(dev)
out = torch.empty((B, n, n, 18 * C), device=dev)
gT = torch.empty_like(T)
f = torch.rand(w["f_size"], device=dev) * 2 - 1
K = (torch.rand((18 * C, Co), device=dev) - 0.5) * 0.1
bias = torch.rand(Co, device=dev) - 0.5
gZ = torch.rand((B * n * n, Co), device=dev) - 0.5
f_off, m, pos, adjl = d(w["f_off"]), d(w["m"]), d(w["pos"]), d(w["adj"])
X = torch.zeros((B, n * n, 18 * C), device=dev)
gf = torch.zeros(w["f_size"], device=dev)


def once():
    ctx.contract18_forward(T, adj, out=out)
    ctx.contract18_backward(gout, adj, gT=gT)
    _, Y, _ = ctx.gather_level_forward(f, f_off, m, pos, adjl, K, bias, n, X=X)
    ctx.gather_level_backward(gZ, X, Y, K, bias, adjl, f_off, m, pos, gf, n)


for _ in range(int(os.environ.get("NCU_WARM", "2"))):
    once()
torch.cuda.synchronize()
once()
torch.cuda.synchronize()
print("ncu target done: launches per pass =", ctx.kernel_launches // (int(os.environ.get("NCU_WARM", "2")) + 1))
