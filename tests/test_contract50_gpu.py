"""GPU parity of StackTensor3D + RisiContraction_50 (BASELINE.json config 5) through the C-ABI: golden fixture from the
compiled reference, the einsum oracle at small and full (N=48, C=128) size, ragged batches, += semantics, and the
size-independent properties (linearity; the 18-way op is a slab subset of the 50-way op)."""
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


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda()


def test_r50_golden(ctx):
    g = np.load(os.path.join(GOLDEN, "r50_n5_c2.npz"))
    out = ctx.contract50_forward(dev(g["T"][None]), dev(g["adj"][None]))
    assert pyoracle.slab_rel_err(out[0].cpu().numpy(), g["out"], 50) < TOL
    gT = dev(g["gT0"][None])
    ctx.contract50_backward(dev(g["gout"][None]), dev(g["adj"][None]), gT=gT, beta=1.0)
    assert pyoracle.slab_rel_err(gT[0].cpu().numpy(), g["gT"], 1) < TOL


@pytest.mark.parametrize("N,C,B,density", [(7, 5, 3, 1.0), (12, 32, 2, 1.0), (48, 128, 1, 1.0),
                                             (48, 128, 2, 0.08), (24, 128, 2, 0.15), (20, 64, 2, 0.1), (48, 128, 1, 0.4),
                                             (13, 36, 2, 0.3), (31, 24, 3, 0.2), (9, 8, 2, 1.0), (50, 8, 1, 0.1)])
def test_r50_vs_einsum(ctx, N, C, B, density):
    """density < 1: non-symmetric weighted sparse adjacency -- the packed-list form of the vector kernels (taken when every row and
    column has at most 16 non-zeros); 0.4 at N=48 and density 1 walk the dense rows.  C = 36 and 24 leave a partial 32-channel
    chunk; N = 50 is past the vector kernels' limit (general tiled path); C = 5 takes the scalar kernels."""
    rng = np.random.default_rng(N * 100 + C)
    T = rng.uniform(-1, 1, (B, N, N, N, C)).astype(np.float32)
    adj = rng.uniform(-1, 1, (B, N, N)).astype(np.float32)
    if density < 1.0:
        adj *= rng.random((B, N, N)) < density
    gout = rng.uniform(-1, 1, (B, N, N, 50 * C)).astype(np.float32)
    out = ctx.contract50_forward(dev(T), dev(adj)).cpu().numpy()
    gT = ctx.contract50_backward(dev(gout), dev(adj)).cpu().numpy()
    for i in range(B):
        assert pyoracle.slab_rel_err(out[i], pyoracle.einsum50_forward(T[i], adj[i]), 50) < TOL
        assert pyoracle.slab_rel_err(gT[i], pyoracle.einsum50_backward(gout[i], adj[i]), 1) < TOL


def test_r50_ragged_batch(ctx):
    rng = np.random.default_rng(3)
    sizes, nm, C = [3, 9, 6, 1], 9, 4
    T = np.zeros((len(sizes), nm ** 3 * C), np.float32)
    adj = np.zeros((len(sizes), nm * nm), np.float32)
    gout = np.zeros((len(sizes), nm * nm * 50 * C), np.float32)
    inst = []
    for i, n in enumerate(sizes):
        t, a, g = rng.uniform(-1, 1, (n, n, n, C)), rng.uniform(-1, 1, (n, n)), rng.uniform(-1, 1, (n, n, 50 * C))
        T[i, :t.size], adj[i, :a.size], gout[i, :g.size] = t.ravel(), a.ravel(), g.ravel()
        inst.append((t, a, g))
    n_dev = torch.tensor(sizes, dtype=torch.int32, device="cuda")
    Td, ad, gd = dev(T.reshape(len(sizes), nm, nm, nm, C)), dev(adj.reshape(len(sizes), nm, nm)), dev(
        gout.reshape(len(sizes), nm, nm, 50 * C))
    out = ctx.contract50_forward(Td, ad, n=n_dev).cpu().numpy().reshape(len(sizes), -1)
    gT = ctx.contract50_backward(gd, ad, n=n_dev).cpu().numpy().reshape(len(sizes), -1)
    for i, (t, a, g) in enumerate(inst):
        n = sizes[i]
        ref = pyoracle.einsum50_forward(t, a)
        assert pyoracle.slab_rel_err(out[i, :ref.size].reshape(ref.shape), ref, 50) < TOL
        refb = pyoracle.einsum50_backward(g, a)
        assert pyoracle.slab_rel_err(gT[i, :refb.size].reshape(refb.shape), refb, 1) < TOL


def test_r50_contains_r18_and_is_linear(ctx):
    """Full config-5 shape, no CPU reference needed: slabs {1,3,5,...,50} of the 50-way op equal the 18-way op on a
    non-negative adjacency, and forward is linear in T."""
    N, C = 48, 128
    g = torch.Generator(device="cuda").manual_seed(5)
    T1 = torch.rand((1, N, N, N, C), device="cuda", generator=g) - 0.5
    T2 = torch.rand((1, N, N, N, C), device="cuda", generator=g) - 0.5
    adj = (torch.rand((1, N, N), device="cuda", generator=g) < 0.1).float()
    o1 = ctx.contract50_forward(T1, adj)
    o2 = ctx.contract50_forward(T2, adj)
    o12 = ctx.contract50_forward(T1 + 2 * T2, adj)
    scale = o12.abs().max().item()
    assert (o12 - (o1 + 2 * o2)).abs().max().item() / scale < 1e-5
    o18 = ctx.contract18_forward(T1, adj).reshape(N, N, 18, C)
    subset = [1, 3, 5, 6, 10, 11, 13, 17, 18, 23, 26, 27, 28, 38, 40, 43, 46, 50]
    sel = o1.reshape(N, N, 50, C)[:, :, [k - 1 for k in subset], :]
    for k in range(18):
        den = o18[:, :, k].abs().max().item()
        assert (sel[:, :, k] - o18[:, :, k]).abs().max().item() / den < 1e-5, k


def test_r50_reference_test_recipe_exact(ctx):
    """The reference R50 test's own inputs (tests/test_RisiContraction_50.cpp:49-80; golden from the compiled reference):
    integer-valued below 2^24, so the fp32 kernels must reproduce forward AND backward bit for bit."""
    g = np.load(os.path.join(GOLDEN, "kat_r50_n10_c5.npz"))
    out = ctx.contract50_forward(dev(g["T"][None]), dev(g["adj"][None]))[0].cpu().numpy()
    assert np.array_equal(out.astype(np.float64), g["out"])
    gT = ctx.contract50_backward(dev(g["gout"][None]), dev(g["adj"][None]))[0].cpu().numpy()
    assert np.array_equal(gT.astype(np.float64), g["gT"])
    # tests/test_RisiContraction_4_thread.cpp:49-66 uses the same tensors: RisiContraction_4 forward + backward, exactly
    o4 = ctx.contract_family_forward(4, dev(g["T"][None]))[0].cpu().numpy()
    t4 = ctx.contract_family_backward(4, dev(g["g4"][None]))[0].cpu().numpy()
    assert np.array_equal(o4.astype(np.float64), g["out4"]) and np.array_equal(t4.astype(np.float64), g["gT4"])
    # the family members that are sub-plans of the 50: RisiContraction_10 = its first ten slabs, exactly
    N, C = g["adj"].shape[0], g["T"].shape[3]
    o10 = ctx.contract_family_forward(10, dev(g["T"][None]), dev(g["adj"][None]))[0].cpu().numpy()
    assert np.array_equal(o10.reshape(N, N, 10, C).astype(np.float64), g["out"].reshape(N, N, 50, C)[:, :, :10])
