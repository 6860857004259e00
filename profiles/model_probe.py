#!/usr/bin/env python
"""Model-level timing probe (BASELINE.json config 2: SMP_beta, 3 levels, C=32, batch of 128 synthetic molecular graphs with
24 vertices, fp32, one B200): forward+backward step time of the batched B200 path, contractions per step = B * V * L, and
the reference's own CPU time for ONE graph of the same batch (unmodified SMP_beta, one core) for scale.
    python profiles/model_probe.py [batch] [V] [L] [C] [--chunk G] [--no-ref]
--chunk G processes the batch G graphs at a time (gradients accumulate), which bounds the activation memory: the level
activations of 512 graphs x 32 vertices at L=4, C=64 do not fit 180 GB at once."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphflow_b200.model import CCNModelB200  # noqa: E402
from tests.util import molecular_adjacency  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
V = int(sys.argv[2]) if len(sys.argv) > 2 else 24
L = int(sys.argv[3]) if len(sys.argv) > 3 else 3
C = int(sys.argv[4]) if len(sys.argv) > 4 else 32
F, D = 5, 2
KIND = sys.argv[sys.argv.index("--kind") + 1] if "--kind" in sys.argv else "beta"  # beta | ver8 | omega
FIELD = int(sys.argv[sys.argv.index("--field") + 1]) if "--field" in sys.argv else V
rng = np.random.default_rng(0)
graphs = []
for _ in range(B):
    adj = (molecular_adjacency(V, rng, self_loops=False) > 0).astype(np.int32)
    graphs.append((adj, np.eye(F)[rng.integers(0, F, V)]))
model = CCNModelB200(KIND, L, C, F, n_depth=D, max_field=FIELD)
nparams = model.num_params()
params = rng.uniform(-1, 1, nparams) * 0.02
model.set_flat_params(params)
t0 = time.perf_counter()
CH = int(sys.argv[sys.argv.index("--chunk") + 1]) if "--chunk" in sys.argv else B
tbs = [model.tables(graphs[i:i + CH]) for i in range(0, B, CH)]
tb = tbs[0]
t_tables = time.perf_counter() - t0


class _Agg:  # what the reporting below reads from the tables, summed over the chunks
    contractions = sum(t.contractions for t in tbs)
    levels = tbs[0].levels
    padded_rows = [sum(t.padded_rows[l] for t in tbs) for l in range(L)]
    real_rows = [sum(t.real_rows[l] for t in tbs) for l in range(L)]


GRAPH = "--graph" in sys.argv  # capture the whole step in ONE CUDA graph (no host launch gaps between the ~100 launches)
targets_dev = [torch.full((len(t.graphs),), float(V), device="cuda") for t in tbs]


def run_step():
    gf0, total = None, None
    for t, tg in zip(tbs, targets_dev):
        gf, loss, grads = model.forward_backward(t, tg)
        gf0 = gf if gf0 is None else gf0
        total = grads if total is None else total + grads
    return gf0, loss, total


for _ in range(2):
    run_step()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 5
if GRAPH:
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            run_step()
    torch.cuda.current_stream().wait_stream(side)
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        gf, loss, grads = run_step()
    cg.replay()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(steps):
        cg.replay()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    kt = {}
else:
    model.ctx.set_kernel_timing(True)
    ev0.record()
    for _ in range(steps):
        gf, loss, grads = run_step()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    kt = {k: v[0] / steps for k, v in model.ctx.kernel_timing().items()}
tb = _Agg
res = {"workload": "%s fwd+bwd, L=%d C=%d%s, %d graphs x %d vertices (%d graphs per chunk)" % (
           {"beta": "SMP_beta", "ver8": "SMP_2D_ver8", "omega": "SMP_omega_physics"}[KIND], L, C,
           " (widths %s, field <= %d)" % (model.widths, FIELD) if KIND == "omega" else "", B, V, CH), "ms_per_step": ms,
       "contractions_per_step": tb.contractions, "contractions_per_s": tb.contractions / (ms * 1e-3),
       "graphs_per_s": B / (ms * 1e-3), "bucket_n_max_per_level": [[b["n_max"] for b in lv] for lv in tb.levels],
       "padded_rows_per_level": tb.padded_rows, "real_rows_per_level": tb.real_rows,
       "host_table_build_s_once": t_tables, "kernels_ms_per_step": kt, "ours_kernel_ms_per_step": sum(kt.values()),
       "cuda_graph": GRAPH}
try:
    from oracle import pyoracle
    if pyoracle.model_available() and "--no-ref" not in sys.argv:
        t0 = time.perf_counter()
        if KIND == "omega":
            ref = pyoracle.ref_smp_omega_physics(graphs[0][0], graphs[0][1], FIELD, L, C, params, float(V))
        elif KIND == "ver8":
            ref = pyoracle.ref_smp_2d_ver8(graphs[0][0], graphs[0][1], L, C, D, params, float(V))
        else:
            ref = pyoracle.ref_smp_beta(graphs[0][0], graphs[0][1], L, C, D, params, float(V))
        dt = time.perf_counter() - t0
        err = float(np.abs(gf[0].cpu().numpy() - ref["feature"]).max() / np.abs(ref["feature"]).max())
        res.update({"reference_cpu_s_per_graph_1core": dt, "reference_graphs_per_s_1core": 1.0 / dt,
                    "feature_rel_err_vs_reference_graph0": err})
except Exception as e:  # noqa: BLE001
    res["reference_error"] = str(e)
print(json.dumps(res))
