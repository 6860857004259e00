"""GPU parity tests for StackTensor3D + RisiContraction_18 forward/backward, through the C-ABI.

Mirrors the reference's own harness (tests/test_RisiContraction_18_gpu.cu:67-234: same-input forward compare, then
random gradient and backward compare), against (a) the committed golden vectors made from the unmodified
reference, (b) the C oracle on seeded inputs, (c) the fp64 closed form at full size, and (d) size-independent
properties (adjoint identity, vertex-permutation equivariance, linearity).

Tolerance: north_star asks 1e-4 relative in fp32; the metric is per-slab max-abs error normalised by the slab's
max-abs.  Integer-valued KATs must match exactly.
"""
import os

import numpy as np
import pytest
import torch

from oracle import pyoracle
from tests.conftest import GOLDEN
from tests.util import molecular_adjacency, per_slab_errors, random_instance

pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.fixture(scope="module")
def ctx():
    import graphflow_b200

    c = graphflow_b200.Context(0)
    yield c
    c.close()


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda()


def run_forward(ctx, T, adj, **kw):
    return ctx.contract18_forward(dev(T[None]), dev(adj[None]), **kw)[0].cpu().numpy()


def run_backward(ctx, gout, adj, gT0=None, **kw):
    if gT0 is None:
        return ctx.contract18_backward(dev(gout[None]), dev(adj[None]), **kw)[0].cpu().numpy()
    g = dev(gT0[None])
    ctx.contract18_backward(dev(gout[None]), dev(adj[None]), gT=g, beta=1.0, **kw)
    return g[0].cpu().numpy()


def assert_slabs(x, ref, C, tol=TOL, what=""):
    errs = per_slab_errors(x, ref, C)
    assert errs.max() <= tol, "%s per-slab errors (case 1..18): %s" % (what, np.array2string(errs, precision=2))


def assert_grad(x, ref, tol=TOL, what=""):
    den = np.abs(ref).max()
    err = np.abs(np.asarray(x, np.float64) - ref).max() / (den if den > 0 else 1.0)
    assert err <= tol, "%s gradient error %.3e" % (what, err)


# ---- golden vectors from the unmodified reference ------------------------------------------------------------------
def test_kat_c1_exact(ctx):
    g = np.load(os.path.join(GOLDEN, "kat_c1_n8_c4.npz"))
    out = run_forward(ctx, g["T"], g["adj"])
    assert np.array_equal(out, g["out"]), per_slab_errors(out, g["out"], 4)
    gT = run_backward(ctx, g["gout"], g["adj"])
    assert np.array_equal(gT, g["gT"])


def test_golden_real_and_accumulate(ctx):
    g = np.load(os.path.join(GOLDEN, "real_n6_c8.npz"))
    assert_slabs(run_forward(ctx, g["T"], g["adj"]), g["out"], 8, what="real_n6_c8 fwd")
    assert_grad(run_backward(ctx, g["gout"], g["adj"], g["gT0"]), g["gT"], what="real_n6_c8 bwd(+=)")


def test_golden_signed_adjacency_both_modes(ctx):
    import graphflow_b200 as gf

    g = np.load(os.path.join(GOLDEN, "signed_n5_c3.npz"))
    assert_slabs(run_forward(ctx, g["T"], g["adj"]), g["out"], 3, what="signed fwd (positive part)")
    assert_slabs(run_forward(ctx, g["T"], g["adj"], adj_mode=gf.ADJ_RAW), g["out_raw"], 3, what="signed fwd (raw)")
    assert_grad(run_backward(ctx, g["gout"], g["adj"]), g["gT"], what="signed bwd")


# ---- C oracle on seeded inputs --------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,C,signed", [(8, 4, False), (10, 5, True), (12, 32, False), (7, 64, True), (16, 128, False),
                                        (33, 8, False), (3, 1, False), (1, 64, False), (9, 16, True), (11, 8, False),
                                        (31, 16, False)])
def test_vs_c_oracle(ctx, n, C, signed):
    rng = np.random.default_rng(1000 * n + C)
    T, adj, gout = random_instance(n, C, rng, signed)
    orc = pyoracle.COracle("f64")
    assert_slabs(run_forward(ctx, T, adj), orc.contract18_forward(T, adj), C, what="fwd n=%d C=%d" % (n, C))
    assert_grad(run_backward(ctx, gout, adj), orc.contract18_backward(gout, adj), what="bwd n=%d C=%d" % (n, C))


# ---- full size, fp64 closed form ------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,C", [(32, 64), (24, 32), (32, 128), (32, 32), (32, 16), (29, 8)])
def test_full_size_vs_closed_form(ctx, n, C):
    rng = np.random.default_rng(n * 7 + C)
    B = 3
    Ts, adjs, gouts = zip(*[random_instance(n, C, rng) for _ in range(B)])
    T, adj, gout = dev(np.stack(Ts)), dev(np.stack(adjs)), dev(np.stack(gouts))
    out = ctx.contract18_forward(T, adj).cpu().numpy()
    gT = ctx.contract18_backward(gout, adj).cpu().numpy()
    for i in range(B):
        assert_slabs(out[i], pyoracle.einsum18_forward(Ts[i], adjs[i]), C, what="fwd inst %d" % i)
        assert_grad(gT[i], pyoracle.einsum18_backward(gouts[i], adjs[i]), what="bwd inst %d" % i)


@pytest.mark.parametrize("other", ["generic"])
def test_fused_and_other_paths_agree(ctx, other):
    from graphflow_b200 import _lib

    rng = np.random.default_rng(5)
    n, C, B = 32, 64, 3
    Ts, adjs, gouts = zip(*[random_instance(n, C, rng, signed_adj=(i == 1)) for i in range(B)])
    T, adj, gout = dev(np.stack(Ts)), dev(np.stack(adjs)), dev(np.stack(gouts))
    out_f = ctx.contract18_forward(T, adj).cpu().numpy()
    gT_f = ctx.contract18_backward(gout, adj).cpu().numpy()
    assert ctx.fused_error_flag() == 0
    ctx.set_kernel_path(_lib.PATH_GENERIC)
    try:
        out_g = ctx.contract18_forward(T, adj).cpu().numpy()
        gT_g = ctx.contract18_backward(gout, adj).cpu().numpy()
    finally:
        ctx.set_kernel_path(_lib.PATH_AUTO)
    for i in range(B):
        assert_slabs(out_f[i], out_g[i], C, tol=2e-5, what="fused vs %s fwd" % other)
        assert_grad(gT_f[i], gT_g[i].astype(np.float64), tol=2e-5, what="fused vs %s bwd" % other)


def test_fused_slot_recycling(ctx):
    """More instances than scratch slots (slots ~ 2 * resident CTAs / tiles): every slot is reused several times, with
    ragged sizes so sibling counts differ between generations.  Checked against the fp64 closed form."""
    rng = np.random.default_rng(77)
    n_max, C, B = 32, 64, 400
    ns = rng.integers(1, n_max + 1, B).astype(np.int32)
    ns[:4] = [32, 1, 4, 5]
    base = [random_instance(int(k), C, rng) for k in (32, 17, 4, 1, 29)]
    T = torch.zeros((B, n_max ** 3 * C), device="cuda")
    adj = torch.zeros((B, n_max * n_max), device="cuda")
    gout = torch.zeros((B, n_max * n_max * 18 * C), device="cuda")
    picks = []
    for i in range(B):
        k = int(ns[i])
        Ti, Ai, Gi = random_instance(k, C, rng) if i < 24 else (None, None, None)
        if Ti is None:  # reuse a small pool (cropped) to keep the host side fast
            src = base[i % len(base)]
            k = min(k, src[0].shape[0])
            ns[i] = k
            Ti, Ai, Gi = src[0][:k, :k, :k], src[1][:k, :k], src[2][:k, :k]
        picks.append((np.ascontiguousarray(Ti), np.ascontiguousarray(Ai), np.ascontiguousarray(Gi)))
        T[i, :k ** 3 * C] = dev(picks[-1][0].ravel())
        adj[i, :k * k] = dev(picks[-1][1].ravel())
        gout[i, :k * k * 18 * C] = dev(picks[-1][2].ravel())
    nd = torch.from_numpy(ns).cuda()
    out = ctx.contract18_forward(T, adj, n=nd, n_max=n_max, C=C, batch=B)
    gT = ctx.contract18_backward(gout, adj, n=nd, n_max=n_max, C=C, batch=B)
    assert ctx.fused_error_flag() == 0
    out, gT = out.reshape(B, -1).cpu().numpy(), gT.reshape(B, -1).cpu().numpy()
    for i in list(range(24)) + list(range(B - 24, B)):
        k = int(ns[i])
        Ti, Ai, Gi = picks[i]
        assert_slabs(out[i, :k * k * 18 * C], pyoracle.einsum18_forward(Ti, Ai).ravel(), C, what="fwd inst %d n=%d" % (i, k))
        assert_grad(gT[i, :k ** 3 * C], pyoracle.einsum18_backward(Gi, Ai).ravel(), what="bwd inst %d n=%d" % (i, k))


@pytest.mark.parametrize("C", [64, 32, 128, 8, 16, 4])
def test_ragged_batch(ctx, C):
    """Instances of different n in one call (dense inside fixed-stride slots), incl. n = 1 and n = n_max."""
    rng = np.random.default_rng(11 + C)
    n_max = 32
    ns = [32, 17, 1, 24, 5, 31, 2]
    B = len(ns)
    sT, sA, sO = n_max ** 3 * C, n_max * n_max, n_max * n_max * 18 * C
    T = np.zeros((B, sT), np.float32)
    adj = np.zeros((B, sA), np.float32)
    gout = np.zeros((B, sO), np.float32)
    inst = []
    for i, n in enumerate(ns):
        t, a, g = random_instance(n, C, rng, signed_adj=(i % 3 == 2))
        inst.append((t, a, g))
        T[i, :t.size], adj[i, :a.size], gout[i, :g.size] = t.ravel(), a.ravel(), g.ravel()
    n_dev = torch.tensor(ns, dtype=torch.int32).cuda()
    out = torch.full((B, sO), float("nan"), device="cuda")
    ctx.contract18_forward(dev(T), dev(adj), out=out, n=n_dev, n_max=n_max, C=C, batch=B, strides=(sT, sA, sO))
    gT = torch.full((B, sT), float("nan"), device="cuda")
    ctx.contract18_backward(dev(gout), dev(adj), gT=gT, n=n_dev, n_max=n_max, C=C, batch=B, strides=(sO, sA, sT))
    out, gT = out.cpu().numpy(), gT.cpu().numpy()
    for i, n in enumerate(ns):
        t, a, g = inst[i]
        assert_slabs(out[i, :n * n * 18 * C], pyoracle.einsum18_forward(t, a), C, what="ragged fwd n=%d" % n)
        assert_grad(gT[i, :t.size].reshape(t.shape), pyoracle.einsum18_backward(g, a), what="ragged bwd n=%d" % n)
        assert np.isnan(out[i, n * n * 18 * C:]).all() and np.isnan(gT[i, t.size:]).all(), "wrote outside instance"


def test_slab_pointer_input_fuses_the_stack(ctx):
    """RisiContraction_18::add_tensor API: N separate [N,N,C] vertex tensors, no materialised stack."""
    rng = np.random.default_rng(21)
    n, C, B = 32, 64, 2
    Ts, adjs, gouts = zip(*[random_instance(n, C, rng) for _ in range(B)])
    slabs = [[dev(Ts[i][a]) for a in range(n)] for i in range(B)]
    table = torch.tensor([[s.data_ptr() for s in row] for row in slabs], dtype=torch.int64).cuda().reshape(-1)
    adj = dev(np.stack(adjs))
    out = ctx.contract18_forward(None, adj, slabs=table, n_max=n, C=C, batch=B).cpu().numpy()
    gs = [[torch.full((n, n, C), 0.5, device="cuda") for _ in range(n)] for _ in range(B)]
    gtable = torch.tensor([[s.data_ptr() for s in row] for row in gs], dtype=torch.int64).cuda().reshape(-1)
    ctx.contract18_backward(dev(np.stack(gouts)), adj, gslabs=gtable, n_max=n, C=C, batch=B, beta=1.0)
    torch.cuda.synchronize()
    for i in range(B):
        assert_slabs(out[i], pyoracle.einsum18_forward(Ts[i], adjs[i]), C, what="slab fwd")
        got = np.stack([g.cpu().numpy() for g in gs[i]])
        assert_grad(got, pyoracle.einsum18_backward(gouts[i], adjs[i]) + 0.5, what="slab bwd (+=)")


def test_raw_mode_fast_path(ctx):
    import graphflow_b200 as gf

    rng = np.random.default_rng(31)
    T, adj, gout = random_instance(32, 64, rng, signed_adj=True)
    assert_slabs(run_forward(ctx, T, adj, adj_mode=gf.ADJ_RAW), pyoracle.einsum18_forward(T, adj, False), 64, what="raw fwd")
    assert_grad(run_backward(ctx, gout, adj, adj_mode=gf.ADJ_RAW), pyoracle.einsum18_backward(gout, adj, False), what="raw bwd")
    assert_slabs(run_forward(ctx, T, adj), pyoracle.einsum18_forward(T, adj, True), 64, what="pos fwd")


def test_dense_adjacency_fast_path(ctx):
    rng = np.random.default_rng(41)
    T, _, gout = random_instance(32, 64, rng)
    adj = rng.uniform(0.1, 1.0, (32, 32)).astype(np.float32)
    assert_slabs(run_forward(ctx, T, adj), pyoracle.einsum18_forward(T, adj), 64, what="dense fwd")
    assert_grad(run_backward(ctx, gout, adj), pyoracle.einsum18_backward(gout, adj), what="dense bwd")


# ---- size-independent properties at BASELINE size --------------------------------------------------------------------
def test_adjoint_identity_full_batch(ctx):
    """<contract(T), G> == <T, contract^T(G)> per instance, N=32 C=64, batch 64 (crosses workspace chunks)."""
    torch.manual_seed(3)
    B, n, C = 64, 32, 64
    rng = np.random.default_rng(3)
    adj = dev(np.stack([molecular_adjacency(n, rng) for _ in range(B)]))
    T = torch.rand((B, n, n, n, C), device="cuda") * 2 - 1
    G = torch.rand((B, n, n, 18 * C), device="cuda") * 2 - 1
    ctx.set_workspace_limit(24 << 20)  # force several chunks
    try:
        out = ctx.contract18_forward(T, adj)
        gT = ctx.contract18_backward(G, adj)
    finally:
        ctx.set_workspace_limit(96 << 20)
    lhs = (out.double() * G.double()).flatten(1).sum(1)
    rhs = (T.double() * gT.double()).flatten(1).sum(1)
    rel = ((lhs - rhs).abs() / lhs.abs().clamp_min(1.0)).max().item()
    assert rel < 1e-5, rel


def test_permutation_equivariance(ctx):
    """Relabelling the receptive field permutes the output: out'[x,y] = out[pi(x), pi(y)]
    (the property behind tests/test_graph_permutation_invariant.cpp)."""
    rng = np.random.default_rng(17)
    n, C = 32, 64
    T, adj, _ = random_instance(n, C, rng)
    pi = rng.permutation(n)
    Tp = T[np.ix_(pi, pi, pi)]
    adjp = adj[np.ix_(pi, pi)]
    out = run_forward(ctx, T, adj)
    outp = run_forward(ctx, Tp, adjp)
    assert_slabs(outp, out[np.ix_(pi, pi)], C, tol=2e-5, what="permutation")


def test_linearity(ctx):
    rng = np.random.default_rng(19)
    n, C = 32, 64
    T1, adj, _ = random_instance(n, C, rng)
    T2 = rng.uniform(-1, 1, T1.shape).astype(np.float32)
    lhs = run_forward(ctx, 2.0 * T1 - 3.0 * T2, adj)
    rhs = 2.0 * run_forward(ctx, T1, adj).astype(np.float64) - 3.0 * run_forward(ctx, T2, adj)
    assert_slabs(lhs, rhs, C, tol=2e-5, what="linearity")


# ---- host-buffer entry points ------------------------------------------------------------------------------------------
def test_host_buffer_pipeline(ctx):
    rng = np.random.default_rng(23)
    n, C, B = 32, 64, 40  # several chunks of the staging ring
    adjs = np.stack([molecular_adjacency(n, rng) for _ in range(B)])
    T = torch.rand((B, n, n, n, C)).mul_(2).sub_(1).pin_memory()
    G = torch.rand((B, n, n, 18 * C)).mul_(2).sub_(1).pin_memory()
    adj = torch.from_numpy(adjs).pin_memory()
    out = torch.empty((B, n, n, 18 * C)).pin_memory()
    gT = torch.empty((B, n, n, n, C)).pin_memory()
    ctx.contract18_forward_backward_host(T, adj, G, out, gT)
    out_d = ctx.contract18_forward(T.cuda(), adj.cuda()).cpu()
    gT_d = ctx.contract18_backward(G.cuda(), adj.cuda()).cpu()
    assert torch.equal(out, out_d) and torch.equal(gT, gT_d)
    for i in (0, B - 1):
        assert_slabs(out[i].numpy(), pyoracle.einsum18_forward(T[i].numpy(), adjs[i]), C, what="host fwd")
    # separate calls, pageable memory, beta = 1
    out2 = ctx.contract18_forward_host(T[:3].clone(), adj[:3].clone())
    assert torch.equal(out2, out[:3])
    g0 = torch.ones((3, n, n, n, C))
    g2 = ctx.contract18_backward_host(G[:3].clone(), adj[:3].clone(), gT=g0, beta=1.0)
    assert torch.allclose(g2, gT[:3] + 1.0, rtol=0, atol=1e-3)


def test_errors_are_reported_not_fatal(ctx):
    import graphflow_b200 as gf

    T = torch.zeros((1, 4, 4, 4, 2), device="cuda")
    adj = torch.zeros((1, 4, 4), device="cuda")
    with pytest.raises(gf.CCNError):
        ctx.contract18_forward(T, adj, adj_mode=7)
    with pytest.raises(gf.CCNError):
        ctx.contract18_forward(None, adj, n_max=4, C=2, batch=1)  # neither T nor slabs
    with pytest.raises(TypeError):
        ctx.contract18_forward(T.double(), adj)
    # the context is still usable
    assert ctx.contract18_forward(T, adj).abs().sum().item() == 0.0
