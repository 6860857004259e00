"""Round-2 exploratory measurements on one B200 (not the bench): A/B of the fused-kernel variants, dense adjacency, the fused
gather / scatter level kernels, the host-buffer paths (pinned / pageable / registered) and the plain host-copy ceiling.
Writes one JSON object to the path given as argv[1]."""
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import graphflow_b200  # noqa: E402
from tests.util import level_tables, molecular_adjacency  # noqa: E402

res = {}
dev = torch.device("cuda", 0)
n, C, B = 32, 64, 512
PEAK = 6538.6
BYTES = 4 * (n ** 3 * C + n * n + 18 * n * n * C)


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


def inputs(dense):
    rng = np.random.default_rng(0)
    g = torch.Generator(device=dev).manual_seed(1)
    T = torch.rand((B, n, n, n, C), device=dev, generator=g) * 2 - 1
    gout = torch.rand((B, n, n, 18 * C), device=dev, generator=g) * 2 - 1
    if dense:
        adj = torch.rand((B, n, n), device=dev, generator=g) + 0.1
    else:
        uniq = [molecular_adjacency(n, rng) for _ in range(64)]
        adj = torch.from_numpy(np.stack([uniq[i % 64] for i in range(B)])).to(dev)
    return T, adj, gout


T, adj, gout = inputs(False)
out = torch.empty((B, n, n, 18 * C), device=dev)
gT = torch.empty((B, n, n, n, C), device=dev)
for var in (0, 1):
    os.environ["CCN_FUSED_VARIANT"] = str(var)
    ctx = graphflow_b200.Context(0)
    tf = timeit(lambda: ctx.contract18_forward(T, adj, out=out))
    tb = timeit(lambda: ctx.contract18_backward(gout, adj, gT=gT))
    res["variant_%d" % var] = {"fwd_ms": tf, "bwd_ms": tb, "fwd_frac": BYTES * B / tf / 1e6 / PEAK, "bwd_frac": BYTES * B / tb / 1e6 / PEAK,
                               "step_per_s": B / (tf + tb) * 1e3}
    ctx.close()
del os.environ["CCN_FUSED_VARIANT"]

ctx = graphflow_b200.Context(0)
Td, adjd, goutd = T, torch.rand((B, n, n), device=dev) + 0.1, gout
tf = timeit(lambda: ctx.contract18_forward(Td, adjd, out=out))
tb = timeit(lambda: ctx.contract18_backward(goutd, adjd, gT=gT))
res["dense_adjacency"] = {"fwd_ms": tf, "bwd_ms": tb, "fwd_frac": BYTES * B / tf / 1e6 / PEAK, "bwd_frac": BYTES * B / tb / 1e6 / PEAK,
                          "nnz": n * n}
del T, gT
torch.cuda.empty_cache()

# ---- fused level: 16 graphs x 32 vertices, every field full (dense T) -------------------------------------------------
rng = np.random.default_rng(3)
G, V, Co = 16, 32, 64
f_off, m, pos, adjs, fb, ib = [], [], [], [], [0], [0]
base = 0
for _ in range(G):
    prev = [list(rng.permutation(V)) for _ in range(V)]
    cur = [list(rng.permutation(V)) for _ in range(V)]
    fo, mm, pp, nn, fsz = level_tables(prev, cur, C, n, base)
    base += fsz
    f_off.append(fo), m.append(mm), pos.append(pp)
    A = molecular_adjacency(V, rng)
    for v in range(V):
        idx = np.asarray(cur[v])
        adjs.append(A[np.ix_(idx, idx)].ravel())
    fb.append(base)
    ib.append(ib[-1] + V)
f_off, m, pos, adjn = np.concatenate(f_off), np.concatenate(m), np.concatenate(pos), np.stack(adjs).astype(np.float32)
Bl = G * V
f = torch.rand(base, device=dev) * 2 - 1
K = (torch.rand((18 * C, Co), device=dev) - 0.5) * 0.1
bias = torch.rand(Co, device=dev) - 0.5
gZ = torch.rand((Bl * n * n, Co), device=dev) - 0.5
d = lambda x, dt: torch.from_numpy(np.ascontiguousarray(x, dt)).to(dev)  # noqa: E731
f_off_d, m_d, pos_d, adj_d = d(f_off, np.int64), d(m, np.int32), d(pos, np.int32), d(adjn.reshape(Bl, n, n), np.float32)
X = torch.zeros((Bl, n * n, 18 * C), device=dev)
gX = torch.empty_like(X)
gf = torch.zeros(base, device=dev)
gK, gb = torch.zeros_like(K), torch.zeros_like(bias)
state = {}


def lvl_fwd():
    state["XYZ"] = ctx.gather_level_forward(f, f_off_d, m_d, pos_d, adj_d, K, bias, n, X=X)


def lvl_bwd():
    _, Y, _ = state["XYZ"]
    ctx.gather_level_backward(gZ, X, Y, K, bias, adj_d, f_off_d, m_d, pos_d, gf, n, gK=gK, gbias=gb, gX=gX)


ctx.set_kernel_timing(True)
t_lf = timeit(lvl_fwd)
t_lb = timeit(lvl_bwd)
kt = ctx.kernel_timing()
ctx.set_kernel_timing(False)
res["fused_level_b512"] = {"fwd_ms": t_lf, "bwd_ms": t_lb, "level_ms": t_lf + t_lb, "instances_per_s": Bl / (t_lf + t_lb) * 1e3,
                           "kernels_ms_per_launch": {k: v[0] / v[1] for k, v in kt.items()}}
# the unfused chain for comparison: promote -> contract -> mix / mix bwd -> contract bwd -> promote bwd
Ts = torch.empty((Bl, n, n, n, C), device=dev)


def chain_fwd():
    ctx.promote_forward(f, f_off_d, m_d, pos_d, n, C, T=Ts)
    ctx.contract18_forward(Ts, adj_d, out=X.view(Bl, n, n, 18 * C))


def chain_bwd():
    ctx.contract18_backward(gX.view(Bl, n, n, 18 * C), adj_d, gT=Ts)
    ctx.promote_backward(Ts, f_off_d, m_d, pos_d, gf)


res["unfused_gather_contract_b512"] = {"fwd_ms": timeit(chain_fwd), "bwd_ms": timeit(chain_bwd)}
res["fused_gather_contract_b512"] = {
    "fwd_ms": timeit(lambda: ctx.gather_contract18_forward(f, f_off_d, m_d, pos_d, adj_d, n, C, out=X.view(Bl, n, n, 18 * C))),
    "bwd_ms": timeit(lambda: ctx.gather_contract18_backward(gX.view(Bl, n, n, 18 * C), adj_d, f_off_d, m_d, pos_d, gf))}
del Ts
torch.cuda.empty_cache()

# ---- host paths ----------------------------------------------------------------------------------------------------------
def wall(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


pin = lambda t: t.pin_memory()  # noqa: E731
hf, hgZ = pin(f.cpu()), pin(gZ.cpu())
hZ, hgf = pin(torch.empty((Bl * n * n, Co))), pin(torch.empty(base))
hK, hb, hgK, hgb = K.cpu(), bias.cpu(), torch.empty_like(K, device="cpu"), torch.empty(Co)
t64 = lambda x: torch.from_numpy(np.ascontiguousarray(x, np.int64))  # noqa: E731
t32 = lambda x: torch.from_numpy(np.ascontiguousarray(x, np.int32))  # noqa: E731
args = (hf, t64(np.asarray(fb)), t64(np.asarray(ib)), t64(f_off), t32(m), t32(pos), torch.from_numpy(adjn), hK, hb, hgZ, hZ, hgf, hgK, hgb, n)
for chunk in (128, 256):
    os.environ["CCN_LEVEL_CHUNK"] = str(chunk)
    tl = wall(lambda: ctx.gather_level_forward_backward_host(*args))
    res["level_host_pinned_chunk%d" % chunk] = {"s_per_call": tl, "instances_per_s": Bl / tl,
                                                 "h2d_bytes": 4 * (base + Bl * n * n * Co + Bl * n * n) + 12 * Bl * n + 4 * Bl * n * n,
                                                 "d2h_bytes": 4 * (base + Bl * n * n * Co)}
del os.environ["CCN_LEVEL_CHUNK"]

Be = 128
T, adj, gout = inputs(False)
hT, hA, hG = T[:Be].cpu(), adj[:Be].cpu(), gout[:Be].cpu()
hO, hGT = torch.empty((Be, n, n, 18 * C)), torch.empty((Be, n, n, n, C))
del T, gout
torch.cuda.empty_cache()
tp = wall(lambda: ctx.contract18_forward_backward_host(hT, hA, hG, hO, hGT), reps=2)
res["op_host_pageable"] = {"instances_per_s": Be / tp}
for t in (hT, hA, hG, hO, hGT):
    ctx.host_register(t)
tr = wall(lambda: ctx.contract18_forward_backward_host(hT, hA, hG, hO, hGT), reps=2)
res["op_host_registered"] = {"instances_per_s": Be / tr}
for t in (hT, hA, hG, hO, hGT):
    ctx.host_unregister(t)
pT, pA, pG, pO, pGT = (pin(t) for t in (hT, hA, hG, hO, hGT))
tpn = wall(lambda: ctx.contract18_forward_backward_host(pT, pA, pG, pO, pGT), reps=2)
res["op_host_pinned"] = {"instances_per_s": Be / tpn}

# ---- plain pinned-copy ceiling: H2D and D2H of the same byte volume concurrently on two streams ---------------------------
nbytes = 1 << 30
hsrc, hdst = pin(torch.empty(nbytes, dtype=torch.uint8)), pin(torch.empty(nbytes, dtype=torch.uint8))
dsrc, ddst = torch.empty(nbytes, dtype=torch.uint8, device=dev), torch.empty(nbytes, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    with torch.cuda.stream(s1):
        ddst.copy_(hsrc, non_blocking=True)
    with torch.cuda.stream(s2):
        hdst.copy_(dsrc, non_blocking=True)


tc = wall(both, reps=5)
with torch.cuda.stream(s1):
    th = wall(lambda: ddst.copy_(hsrc, non_blocking=True), reps=5)
res["host_copy_ceiling"] = {"duplex_gbs_each_way": nbytes / tc / 1e9, "h2d_only_gbs": nbytes / th / 1e9}

# ---- the reference's own CUDA kernels, rebuilt for sm_100a -----------------------------------------------------------------
exe = os.path.join(ROOT, "oracle", "_ref", "ref_gpu_bench")
if os.path.exists(exe):
    try:
        o = subprocess.run([exe, "32", "64", "3"], capture_output=True, text=True, timeout=600).stdout.strip().splitlines()[-1]
        res["ref_gpu_kernels"] = json.loads(o)
    except Exception as e:  # noqa: BLE001
        res["ref_gpu_kernels"] = {"error": repr(e)}
json.dump(res, open(sys.argv[1], "w"), indent=1)
print(json.dumps(res, indent=1))
