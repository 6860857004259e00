"""GPU parity (through the C-ABI) of the promotion gather / scatter-add, TensorMul and CustomMatMulTensor against the
committed fixtures and the plain-C oracle.  The promotion is a copy: bit-exact forward."""
import os

import numpy as np
import pytest
import torch

from oracle import pyoracle
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def ctx():
    import graphflow_b200

    c = graphflow_b200.Context(0)
    yield c
    c.close()


def dev(x, dt=np.float32):
    return torch.from_numpy(np.ascontiguousarray(x, dt)).cuda()


def rel(x, ref):
    ref = np.asarray(ref, np.float64)
    den = np.abs(ref).max()
    return np.abs(np.asarray(x, np.float64) - ref).max() / (den if den > 0 else 1.0)


def test_golden(ctx):
    g = np.load(os.path.join(GOLDEN, "aux_ops.npz"))
    out = ctx.tensor_mul_forward(dev(g["tm_A"][None]), dev(g["tm_B"][None]))
    assert rel(out[0].cpu().numpy(), g["tm_out"]) < TOL
    gA, gB = dev(g["tm_gA0"][None]), dev(g["tm_gB0"][None])
    ctx.tensor_mul_backward(dev(g["tm_A"][None]), dev(g["tm_B"][None]), dev(g["tm_g"][None]), gA=gA, gB=gB, beta=1.0)
    assert rel(gA[0].cpu().numpy(), g["tm_gA"]) < TOL and rel(gB[0].cpu().numpy(), g["tm_gB"]) < TOL
    Y = ctx.custom_matmul_tensor_forward(dev(g["cm_Kt"]), dev(g["cm_X"]))
    assert rel(Y.cpu().numpy(), g["cm_Y"]) < TOL
    gKt, gX = dev(g["cm_gKt0"]), dev(g["cm_gX0"])
    ctx.custom_matmul_tensor_backward(dev(g["cm_Kt"]), dev(g["cm_X"]), dev(g["cm_gY"]), gKt=gKt, gX=gX, beta_x=1.0)
    assert rel(gKt.cpu().numpy(), g["cm_gKt"]) < TOL and rel(gX.cpu().numpy(), g["cm_gX"]) < TOL
    # promotion: one instance with a single slab padded to n_max = 6
    f, pos = g["pr_f"], g["pr_pos"]
    n, m, C = 6, 4, 3
    f_off = torch.zeros(n, dtype=torch.int64, device="cuda")
    mm = torch.full((n,), m, dtype=torch.int32, device="cuda")
    P = dev(np.tile(pos, (n, 1)), np.int32).reshape(-1)
    T = ctx.promote_forward(dev(f).reshape(-1), f_off, mm, P, n, C)
    for a in range(n):
        assert np.array_equal(T[0, a].cpu().numpy(), g["pr_Q"].astype(np.float32))
    gT = torch.zeros((1, n, n, n, C), device="cuda")
    gT[0, 2] = dev(g["pr_gQ"])
    gf = dev(g["pr_gf0"]).reshape(-1).clone()
    ctx.promote_backward(gT, f_off, mm, P, gf)
    assert rel(gf.cpu().numpy().reshape(m, m, C), g["pr_gf"]) < TOL


def test_custom_matmul_tensor_tensor_core_shape(ctx):
    """SMP_2D_ver8's mix shape (K stored [C, 18C], SMP_2D_ver8.h:130,526-527): runs on the tcgen05 kernel."""
    rng = np.random.default_rng(4)
    N, C = 16, 32
    Kt = rng.uniform(-0.2, 0.2, (C, 18 * C))
    X = rng.uniform(-1, 1, (N, N, 18 * C))
    gY = rng.uniform(-1, 1, (N, N, C))
    c = pyoracle.COracle("f64")
    before = ctx.kernel_timing if False else None
    Y = ctx.custom_matmul_tensor_forward(dev(Kt), dev(X))
    assert rel(Y.cpu().numpy(), c.custom_matmul_tensor_forward(Kt, X)) < TOL
    gKt, gX = ctx.custom_matmul_tensor_backward(dev(Kt), dev(X), dev(gY))
    rK, rX = c.custom_matmul_tensor_backward(Kt, X, gY)
    assert rel(gKt.cpu().numpy(), rK) < TOL and rel(gX.cpu().numpy(), rX) < TOL


def test_tensor_mul_batch(ctx):
    rng = np.random.default_rng(5)
    Bt, N, C = 3, 24, 32
    A, B, g = (rng.uniform(-1, 1, (Bt, N, N, C)) for _ in range(3))
    c = pyoracle.COracle("f64")
    out = ctx.tensor_mul_forward(dev(A), dev(B)).cpu().numpy()
    gA, gB = ctx.tensor_mul_backward(dev(A), dev(B), dev(g))
    for i in range(Bt):
        assert rel(out[i], c.tensor_mul_forward(A[i], B[i])) < TOL
        rA, rB = c.tensor_mul_backward(A[i], B[i], g[i])
        assert rel(gA[i].cpu().numpy(), rA) < TOL and rel(gB[i].cpu().numpy(), rB) < TOL


def test_promote_batch_ragged_and_shared_sources(ctx):
    """A level in miniature: every slab of every instance picks one of a pool of level l-1 tensors of different sizes
    (shared between instances, so the backward really collides) through a random partial selection."""
    rng = np.random.default_rng(6)
    C, n_max, B = 8, 7, 5
    sizes = [1, 3, 5, 7, 4]
    pool, offs, off = [], [], 0
    for m in sizes:
        pool.append(rng.uniform(-1, 1, (m, m, C)))
        offs.append(off)
        off += m * m * C
    f = np.concatenate([p.ravel() for p in pool])
    ns = rng.integers(1, n_max + 1, B).astype(np.int32)
    f_off = np.zeros((B, n_max), np.int64)
    mm = np.ones((B, n_max), np.int32)
    pos = np.full((B, n_max, n_max), -1, np.int32)
    src = np.zeros((B, n_max), np.int64)
    for i in range(B):
        for a in range(ns[i]):
            w = rng.integers(len(sizes))
            src[i, a], f_off[i, a], mm[i, a] = w, offs[w], sizes[w]
            k = rng.integers(0, min(ns[i], sizes[w]) + 1)
            rows = rng.choice(ns[i], k, replace=False)
            pos[i, a, rows] = rng.choice(sizes[w], k, replace=False)
    c = pyoracle.COracle("f32")
    n_dev = torch.from_numpy(ns).cuda()
    T = ctx.promote_forward(dev(f), dev(f_off, np.int64).reshape(-1), dev(mm, np.int32).reshape(-1), dev(pos, np.int32).reshape(-1),
                            n_max, C, n=n_dev)
    Th = T.cpu().numpy().reshape(B, -1)
    gT = rng.uniform(-1, 1, (B, n_max ** 3 * C)).astype(np.float32)
    gf_ref = [np.zeros_like(p) for p in pool]
    for i in range(B):
        n = ns[i]
        for a in range(n):
            want = c.promote_forward(pool[src[i, a]], pos[i, a, :n])
            got = Th[i, a * n * n * C:(a + 1) * n * n * C].reshape(n, n, C)
            assert np.array_equal(got, want)
            gq = gT[i, a * n * n * C:(a + 1) * n * n * C].reshape(n, n, C)
            gf_ref[src[i, a]] = pyoracle.COracle("f64").promote_backward(gq, pos[i, a, :n], sizes[src[i, a]], gf_ref[src[i, a]])
    gf = torch.zeros(f.size, device="cuda")
    ctx.promote_backward(dev(gT).reshape(B, n_max, n_max, n_max, C), dev(f_off, np.int64).reshape(-1), dev(mm, np.int32).reshape(-1),
                         dev(pos, np.int32).reshape(-1), gf, n=n_dev)
    assert rel(gf.cpu().numpy(), np.concatenate([p.ravel() for p in gf_ref])) < 1e-5
