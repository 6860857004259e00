#!/usr/bin/env python
"""Per-phase timeline of the fused kernels (globaltimer marks of thread 0 of every tile): prints the mean and the
percentiles of every phase.  Run on the GPU box:  python profiles/phase_trace.py [batch]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphflow_b200  # noqa: E402
from bench import molecular_adjacency  # noqa: E402


def make_inputs(B, n, C, seed, device):
    g = torch.Generator(device=device).manual_seed(seed)
    T = torch.rand((B, n, n, n, C), device=device, generator=g) * 2 - 1
    gout = torch.rand((B, n, n, 18 * C), device=device, generator=g) * 2 - 1
    rng = np.random.default_rng(seed)
    adj = torch.from_numpy(np.stack([molecular_adjacency(n, rng) for _ in range(B)]).astype(np.float32)).to(device)
    return T, adj, gout

FWD = ["adjacency+acquire", "stream", "partials+publish+passA", "wait siblings", "pass B"]
BWD = ["adjacency+acquire", "a-side sweep+publish", "b-side", "wait siblings", "phase 1c", "stream gT"]


def report(name, tr, labels):
    tr = tr.cpu().numpy().astype(np.float64)
    tr = tr[tr[:, 0] > 0]
    print("%s: %d tiles, kernel span %.1f us" % (name, len(tr), (tr[:, :len(labels) + 1].max() - tr[:, 0].min()) / 1e3))
    for k, lab in enumerate(labels):
        d = (tr[:, k + 1] - tr[:, k]) / 1e3
        print("  %-28s mean %7.2f us   p10 %7.2f   p50 %7.2f   p90 %7.2f" % (lab, d.mean(), *np.percentile(d, [10, 50, 90])))
    life = (tr[:, len(labels)] - tr[:, 0]) / 1e3
    print("  %-28s mean %7.2f us   p10 %7.2f   p50 %7.2f   p90 %7.2f" % ("tile lifetime", life.mean(), *np.percentile(life, [10, 50, 90])))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    n, C = 32, 64
    ctx = graphflow_b200.Context(0)
    T, adj, gout = make_inputs(B, n, C, 1, torch.device("cuda", 0))
    out = torch.empty((B, n, n, 18 * C), device="cuda")
    gT = torch.empty((B, n, n, n, C), device="cuda")
    for _ in range(3):
        ctx.contract18_forward(T, adj, out=out)
        ctx.contract18_backward(gout, adj, gT=gT)
    tiles = B * (n * C // 128)  # 256-thread forward tiles, 128-thread backward tiles: sized for the larger count
    tr = torch.zeros((tiles, 8), dtype=torch.int64, device="cuda")
    ctx.set_phase_trace(tr)
    ctx.contract18_forward(T, adj, out=out)
    torch.cuda.synchronize()
    report("forward", tr, FWD)
    tr.zero_()
    ctx.contract18_backward(gout, adj, gT=gT)
    torch.cuda.synchronize()
    report("backward", tr, BWD)
    # the fused-promotion variants (gather forward / scatter backward): 256-thread tiles, same marks
    from bench import level_workload
    G = max(1, B // n)
    w = level_workload(G, n, C, 3)
    Bl = w["instances"]
    dd = lambda x: torch.from_numpy(x).cuda()  # noqa: E731
    f = torch.rand(w["f_size"], device="cuda") * 2 - 1
    f_off, m, pos, adjl = dd(w["f_off"]), dd(w["m"]), dd(w["pos"]), dd(w["adj"])
    X = torch.zeros((Bl, n, n, 18 * C), device="cuda")
    gX = torch.rand((Bl, n, n, 18 * C), device="cuda") - 0.5
    gf = torch.zeros(w["f_size"], device="cuda")
    for _ in range(3):
        ctx.gather_contract18_forward(f, f_off, m, pos, adjl, n, C, out=X)
        ctx.gather_contract18_backward(gX, adjl, f_off, m, pos, gf)
    tr.zero_()
    ctx.gather_contract18_forward(f, f_off, m, pos, adjl, n, C, out=X)
    torch.cuda.synchronize()
    report("gather forward", tr, FWD)
    tr.zero_()
    ctx.gather_contract18_backward(gX, adjl, f_off, m, pos, gf)
    torch.cuda.synchronize()
    report("scatter backward", tr, BWD)
    ctx.set_phase_trace(None)


if __name__ == "__main__":
    main()
