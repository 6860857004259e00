#!/usr/bin/env python
"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python profiles/sanitize_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphflow_b200  # noqa: E402
from graphflow_b200 import _lib  # noqa: E402

ctx = graphflow_b200.Context(0)
g = torch.Generator(device="cuda").manual_seed(0)


def rnd(*shape):
    return torch.rand(shape, device="cuda", generator=g) * 2 - 1


# fused contraction kernels (uniform and ragged), generic kernels
for (B, n, C, ragged) in ((3, 32, 64, False), (5, 19, 32, True), (2, 8, 4, False), (3, 7, 5, True)):
    T, adj, gout = rnd(B, n, n, n, C), (rnd(B, n, n) > 0.6).float(), rnd(B, n, n, 18 * C)
    nd = torch.tensor([max(1, n - 3 * i) for i in range(B)], dtype=torch.int32, device="cuda") if ragged else None
    out = ctx.contract18_forward(T, adj, n=nd)
    gT = ctx.contract18_backward(gout, adj, n=nd)
    gT = ctx.contract18_backward(gout, adj, gT=gT, n=nd, beta=1.0)
# RisiContraction_50: small (one channel per thread), the vectorised tiled kernels with sparse lists (N=20, C=128, sparse
# adjacency) and with dense tiles, ragged
T, adj, gout = rnd(2, 9, 9, 9, 8), rnd(2, 9, 9), rnd(2, 9, 9, 50 * 8)
ctx.contract50_forward(T, adj)
ctx.contract50_backward(gout, adj)
for dense in (False, True):
    n, C = 20, 128
    T, gout = rnd(2, n, n, n, C), rnd(2, n, n, 50 * C)
    adj = rnd(2, n, n) * (1.0 if dense else (rnd(2, n, n) > 0.7).float())
    nd = torch.tensor([n, n - 7], dtype=torch.int32, device="cuda")
    ctx.contract50_forward(T, adj, n=nd)
    ctx.contract50_backward(gout, adj, n=nd)
# the rest of the family: RisiContraction_4 / _10, slab dropout (train mask, test-mode scale)
n, C = 12, 32
T, adj = rnd(2, n, n, n, C), rnd(2, n, n) * (rnd(2, n, n) > 0.5).float()
use = [k % 3 != 0 for k in range(18)]
ctx.contract_family_forward(4, T)
ctx.contract_family_backward(4, rnd(2, n, n, 4 * C))
ctx.contract_family_forward(10, T, adj)
ctx.contract_family_backward(10, rnd(2, n, n, 10 * C), adj)
ctx.contract_family_forward(18, T, adj, keep_mask=use)
ctx.contract_family_forward(18, T, adj, out_scale=0.5)
ctx.contract_family_backward(18, rnd(2, n, n, 18 * C), adj, keep_mask=use)
# optimizers
p_, g_, m_, v_ = rnd(1000), rnd(1000), torch.zeros(1000, device="cuda"), torch.zeros(1000, device="cuda")
ctx.adam_step(p_, g_, m_, v_, 1e-3, 4, 0, True)
ctx.adam_step(p_, g_, m_, v_, 1e-3, 1, 1, False)
ctx.momentum_step(p_, g_, m_, 1e-2, 0.9, 4)
# feature mix: tensor-core forward / grad-X / grad-W and SIMT
for (M, K, P) in ((700, 1152, 64), (4100, 72, 32), (130, 36, 48), (33, 18, 3)):
    X, W, b, gZ = rnd(M, K), rnd(K, P) * 0.1, rnd(P), rnd(M, P)
    Y, Z = ctx.mix_forward(X, W, b)
    ctx.mix_backward(X, W, gZ, bias=b, Y=Y)
ctx.set_mix_path(_lib.MIX_SIMT)
Y, Z = ctx.mix_forward(rnd(300, 72), rnd(72, 16), rnd(16))
ctx.set_mix_path(_lib.MIX_AUTO)
# aux ops
ctx.tensor_mul_forward(rnd(2, 6, 5, 4), rnd(2, 5, 7, 4))
ctx.custom_matmul_tensor_forward(rnd(16, 72), rnd(5, 5, 72))
f = rnd(3 * 3 * 4)
pos = torch.tensor([0, 2, -1, 1] * 4, dtype=torch.int32, device="cuda")
ctx.promote_forward(f, torch.zeros(4, dtype=torch.int64, device="cuda"), torch.full((4,), 3, dtype=torch.int32, device="cuda"), pos, 4, 4)
# host pipeline
n, C = 16, 32
hT, hA, hG = rnd(3, n, n, n, C).cpu().pin_memory(), (rnd(3, n, n) > 0).float().cpu().pin_memory(), rnd(3, n, n, 18 * C).cpu().pin_memory()
hO, hGT = torch.empty((3, n, n, 18 * C)).pin_memory(), torch.empty((3, n, n, n, C)).pin_memory()
ctx.contract18_forward_backward_host(hT, hA, hG, hO, hGT)
# round 2: fused promotion (gather forward / scatter backward) with ragged fields and absent members, dense adjacency (the
# out-of-line dense product), an empty instance inside a ragged batch, the level entry points, the read-out, the level stack
import numpy as np  # noqa: E402

from tests.util import level_tables, molecular_adjacency  # noqa: E402

rng = np.random.default_rng(0)
for (V, C, nm) in ((12, 32, 12), (9, 8, 9), (20, 64, 16)):
    prev = [list(rng.permutation(V)[:rng.integers(1, V + 1)]) for _ in range(V)]
    cur = [list(rng.permutation(V)[:rng.integers(1, min(V, nm) + 1)]) for _ in range(V)]
    fo, mm, pp, nn, fsz = level_tables(prev, cur, C, nm)
    A = molecular_adjacency(V, rng)
    adjl = np.zeros((V, nm * nm), np.float32)
    for v in range(V):
        idx = np.asarray(cur[v])
        adjl[v, :len(idx) ** 2] = A[np.ix_(idx, idx)].ravel()
    d = lambda x, dt: torch.from_numpy(np.ascontiguousarray(x, dt)).cuda()  # noqa: E731
    f = rnd(fsz)
    K, b = rnd(18 * C, C) * 0.1, rnd(C)
    X, Y, Z = ctx.gather_level_forward(f, d(fo, np.int64), d(mm, np.int32), d(pp, np.int32), d(adjl, np.float32), K, b, nm,
                                       n=d(nn, np.int32))
    gZ = rnd(V * nm * nm, C) * (torch.arange(nm * nm, device="cuda")[None, :] < d(nn.astype(np.int64) ** 2, np.int64)[:, None]).reshape(-1, 1)
    gf = torch.zeros(fsz, device="cuda")
    ctx.gather_level_backward(gZ, X, Y, K, b, d(adjl, np.float32), d(fo, np.int64), d(mm, np.int32), d(pp, np.int32), gf, nm, n=d(nn, np.int32))
B, n, C = 6, 32, 64
T, gout = rnd(B, n, n, n, C), rnd(B, n, n, 18 * C)
adj_dense = torch.rand((B, n, n), device="cuda", generator=g) + 0.1
nd = torch.tensor([32, 0, 17, 32, 5, 0], dtype=torch.int32, device="cuda")
ctx.contract18_forward(T, adj_dense, n=nd)
ctx.contract18_backward(gout, adj_dense, n=nd)
# read-out head + loss
Zl, W, tgt = rnd(7, 25, 16), rnd(16), rnd(2)
ptr = torch.tensor([0, 3, 7], dtype=torch.int64, device="cuda")
ig = torch.tensor([0, 0, 0, 1, 1, 1, 1], dtype=torch.int32, device="cuda")
nr = torch.tensor([5, 3, 1, 4, 5, 2, 5], dtype=torch.int32, device="cuda")
shr, gfe, pred, loss = torch.empty((7, 16), device="cuda"), torch.empty((2, 16), device="cuda"), torch.empty(2, device="cuda"), torch.empty(2, device="cuda")
ctx._rc(ctx.lib.ccn_readout_forward(ctx.h, Zl.data_ptr(), 25 * 16, nr.data_ptr(), 5, 16, 7, ptr.data_ptr(), 2, W.data_ptr(), tgt.data_ptr(), 0.01,
                                    shr.data_ptr(), gfe.data_ptr(), pred.data_ptr(), loss.data_ptr(), None))
gZl, gW = torch.empty_like(Zl), torch.zeros(16, device="cuda")
ctx._rc(ctx.lib.ccn_readout_backward(ctx.h, shr.data_ptr(), gfe.data_ptr(), pred.data_ptr(), tgt.data_ptr(), W.data_ptr(), ig.data_ptr(),
                                     nr.data_ptr(), 5, 16, 7, 2, 0.01, gZl.data_ptr(), 25 * 16, gW.data_ptr(), None))
torch.cuda.synchronize()
print("sanitize probe done, fused error flag =", ctx.fused_error_flag())
ctx.close()
