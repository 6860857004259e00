"""A/B of the fused-kernel variant switches (CCN_FUSED_VARIANT bit mask, contract18_fused.cu): forward / backward ms and roofline
fraction at N=32, C=64, 512 instances, molecular adjacency, for every variant given on the command line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import graphflow_b200  # noqa: E402
from tests.util import molecular_adjacency  # noqa: E402

n, C, B, PEAK = 32, 64, 512, 6538.6
BYTES = 4 * (n ** 3 * C + n * n + 18 * n * n * C)
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
g = torch.Generator(device=dev).manual_seed(1)
T = torch.rand((B, n, n, n, C), device=dev, generator=g) * 2 - 1
gout = torch.rand((B, n, n, 18 * C), device=dev, generator=g) * 2 - 1
uniq = [molecular_adjacency(n, rng) for _ in range(64)]
adj = torch.from_numpy(np.stack([uniq[i % 64] for i in range(B)])).to(dev)
out = torch.empty((B, n, n, 18 * C), device=dev)
gT = torch.empty((B, n, n, n, C), device=dev)


def timeit(fn, reps=30, warm=3):
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


res = {"lib": os.environ.get("CCN_B200_LIB", "default")}
for var in [int(v) for v in sys.argv[2:]] or [0]:
    os.environ["CCN_FUSED_VARIANT"] = str(var)
    ctx = graphflow_b200.Context(0)
    tf = timeit(lambda: ctx.contract18_forward(T, adj, out=out))
    tb = timeit(lambda: ctx.contract18_backward(gout, adj, gT=gT))
    res["variant_%d" % var] = {"fwd_ms": tf, "bwd_ms": tb, "fwd_frac": BYTES * B / tf / 1e6 / PEAK, "bwd_frac": BYTES * B / tb / 1e6 / PEAK,
                               "step_per_s": B / (tf + tb) * 1e3}
    if var == 0:  # dense positive adjacency (Coulomb-matrix mode)
        adj_d = torch.rand((B, n, n), device=dev, generator=g) + 0.1
        tf = timeit(lambda: ctx.contract18_forward(T, adj_d, out=out))
        tb = timeit(lambda: ctx.contract18_backward(gout, adj_d, gT=gT))
        res["dense_adjacency"] = {"fwd_ms": tf, "bwd_ms": tb, "fwd_frac": BYTES * B / tf / 1e6 / PEAK, "bwd_frac": BYTES * B / tb / 1e6 / PEAK}
    ctx.close()
json.dump(res, open(sys.argv[1], "w"), indent=1)
print(json.dumps(res, indent=1))
