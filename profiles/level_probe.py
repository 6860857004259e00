"""Device-resident timing of the fused-promotion level kernels (16 graphs x 32 vertices, full fields, C = 64 -> 64) for every
CCN_FUSED_VARIANT given on the command line.  argv[1] = output JSON."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import graphflow_b200  # noqa: E402
from bench import level_workload  # noqa: E402

n = C = Co = None
n, C, Co, G = 32, 64, 64, 16
dev = torch.device("cuda", 0)
w = level_workload(G, n, C, 3)
Bl = w["instances"]
d = lambda x: torch.from_numpy(x).to(dev)  # noqa: E731
f = torch.rand(w["f_size"], device=dev) * 2 - 1
K = (torch.rand((18 * C, Co), device=dev) - 0.5) * 0.1
bias = torch.rand(Co, device=dev) - 0.5
gZ = torch.rand((Bl * n * n, Co), device=dev) - 0.5
f_off, m, pos, adj = d(w["f_off"]), d(w["m"]), d(w["pos"]), d(w["adj"])
X = torch.zeros((Bl, n * n, 18 * C), device=dev)
gX = torch.rand((Bl, n, n, 18 * C), device=dev) - 0.5
gf = torch.zeros(w["f_size"], device=dev)


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {}
ref_gf = None
for var in [int(v) for v in sys.argv[2:]] or [0]:
    os.environ["CCN_FUSED_VARIANT"] = str(var)
    ctx = graphflow_b200.Context(0)
    tf = timeit(lambda: ctx.gather_contract18_forward(f, f_off, m, pos, adj, n, C, out=X.view(Bl, n, n, 18 * C)))
    tb = timeit(lambda: ctx.gather_contract18_backward(gX, adj, f_off, m, pos, gf))
    gf.zero_()
    ctx.gather_contract18_backward(gX, adj, f_off, m, pos, gf)
    torch.cuda.synchronize()
    if ref_gf is None:
        ref_gf = gf.clone()
    res["variant_%d" % var] = {"gather_fwd_ms_per_512": tf, "scatter_bwd_ms_per_512": tb,
                               "max_rel_diff_vs_first_variant": float((gf - ref_gf).abs().max() / ref_gf.abs().max())}
    ctx.close()
json.dump(res, open(sys.argv[1], "w"), indent=1)
print(json.dumps(res, indent=1))
