"""GPU parity of the other members of the contraction family through the C-ABI (`ccn_contract_family_*`):
RisiContraction_4, RisiContraction_10 and RisiContraction_18_dropout (train-mode slab masks, test-mode scaling) against
the golden vectors of the compiled reference and the einsum oracle, ragged batches, `+=` semantics, and the
cross-checks with the 18-way and 50-way kernels."""
import os

import numpy as np
import pytest
import torch

from oracle import pyoracle

pytestmark = pytest.mark.gpu
TOL = 1e-4
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "family_n5_c2.npz")


@pytest.fixture(scope="module")
def ctx():
    import graphflow_b200

    c = graphflow_b200.Context(0)
    yield c
    c.close()


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda()


def err(x, ref, slabs):
    return pyoracle.slab_rel_err(x.cpu().numpy() if torch.is_tensor(x) else x, ref, slabs)


def test_family_golden(ctx):
    g = np.load(GOLDEN)
    T, adj = dev(g["T"][None]), dev(g["adj"][None])
    assert err(ctx.contract_family_forward(4, T)[0], g["out4"], 4) < TOL
    assert err(ctx.contract_family_backward(4, dev(g["g4"][None]))[0], g["gT4"], 1) < TOL
    assert err(ctx.contract_family_forward(10, T, adj)[0], g["out10"], 10) < TOL
    assert err(ctx.contract_family_backward(10, dev(g["g10"][None]), adj)[0], g["gT10"], 1) < TOL
    for i in range(3):
        use = [bool(u) for u in g["drop%d_use" % i]]
        out = ctx.contract_family_forward(18, T, adj, keep_mask=use)[0].cpu().numpy()
        ref = g["drop%d_out" % i]
        N, C = ref.shape[0], ref.shape[2] // 18
        dropped = [k for k in range(18) if not use[k]]
        assert not out.reshape(N, N, 18, C)[:, :, dropped].any()          # exact zeros, as the reference leaves them
        keep = [k for k in range(18) if use[k]]
        assert pyoracle.slab_rel_err(out.reshape(N, N, 18, C)[:, :, keep].reshape(N, N, -1),
                                     ref.reshape(N, N, 18, C)[:, :, keep].reshape(N, N, -1), len(keep)) < TOL
        assert err(ctx.contract_family_backward(18, dev(g["g18"][None]), adj, keep_mask=use)[0], g["drop%d_gT" % i], 1) < TOL
    out = ctx.contract_family_forward(18, T, adj, out_scale=float(g["test_kept"]) / 18.0)[0]
    assert err(out, g["test_out"], 18) < TOL


@pytest.mark.parametrize("N,C,B,density", [(7, 5, 3, 1.0), (12, 32, 2, 1.0), (24, 128, 2, 0.15), (48, 128, 1, 0.08),
                                             (13, 36, 2, 0.3), (31, 24, 2, 0.2), (50, 8, 1, 0.1)])
def test_family_vs_einsum(ctx, N, C, B, density):
    rng = np.random.default_rng(N * 31 + C)
    T = rng.uniform(-1, 1, (B, N, N, N, C)).astype(np.float32)
    adj = rng.uniform(-1, 1, (B, N, N)).astype(np.float32)
    if density < 1.0:
        adj *= rng.random((B, N, N)) < density
    use = [bool(u) for u in rng.integers(0, 2, 18)]
    use[int(rng.integers(0, 18))] = True
    g = {k: rng.uniform(-1, 1, (B, N, N, k * C)).astype(np.float32) for k in (4, 10, 18)}
    o4 = ctx.contract_family_forward(4, dev(T)).cpu().numpy()
    o10 = ctx.contract_family_forward(10, dev(T), dev(adj)).cpu().numpy()
    o18 = ctx.contract_family_forward(18, dev(T), dev(adj), keep_mask=use).cpu().numpy()
    t4 = ctx.contract_family_backward(4, dev(g[4])).cpu().numpy()
    t10 = ctx.contract_family_backward(10, dev(g[10]), dev(adj)).cpu().numpy()
    t18 = ctx.contract_family_backward(18, dev(g[18]), dev(adj), keep_mask=use).cpu().numpy()
    for i in range(B):
        assert pyoracle.slab_rel_err(o4[i], pyoracle.einsum4_forward(T[i]), 4) < TOL
        assert pyoracle.slab_rel_err(o10[i], pyoracle.einsum10_forward(T[i], adj[i]), 10) < TOL
        ref = pyoracle.einsum18_dropout_forward(T[i], adj[i], use)
        assert np.abs(o18[i] - ref).max() / np.abs(ref).max() < TOL
        assert pyoracle.slab_rel_err(t4[i], pyoracle.einsum4_backward(g[4][i]), 1) < TOL
        assert pyoracle.slab_rel_err(t10[i], pyoracle.einsum10_backward(g[10][i], adj[i]), 1) < TOL
        assert pyoracle.slab_rel_err(t18[i], pyoracle.einsum18_dropout_backward(g[18][i], adj[i], use), 1) < TOL


def test_family_cross_checks(ctx):
    """All slabs kept: the 18-way variant equals ccn_contract18_*; variant 10 equals the first ten slabs of the 50."""
    N, C, B = 16, 32, 3
    gen = torch.Generator(device="cuda").manual_seed(9)
    T = torch.rand((B, N, N, N, C), device="cuda", generator=gen) * 2 - 1
    adj = torch.rand((B, N, N), device="cuda", generator=gen) * 2 - 1
    adj = adj * (torch.rand((B, N, N), device="cuda", generator=gen) < 0.3)
    o18 = ctx.contract18_forward(T, adj)
    f18 = ctx.contract_family_forward(18, T, adj)
    assert ((o18 - f18).abs().max() / o18.abs().max()).item() < 1e-5
    o50 = ctx.contract50_forward(T, adj).reshape(B, N, N, 50, C)
    f10 = ctx.contract_family_forward(10, T, adj).reshape(B, N, N, 10, C)
    assert ((o50[:, :, :, :10] - f10).abs().max() / f10.abs().max()).item() < 1e-5  # (atomic sums: not bitwise)
    g = torch.rand((B, N, N, 18 * C), device="cuda", generator=gen) * 2 - 1
    t18 = ctx.contract18_backward(g, adj)
    tf = ctx.contract_family_backward(18, g, adj)
    assert ((t18 - tf).abs().max() / t18.abs().max()).item() < 1e-5


def test_family_ragged_and_accumulate(ctx):
    rng = np.random.default_rng(5)
    sizes, nm, C = [3, 9, 6, 1], 9, 8
    B = len(sizes)
    T = np.zeros((B, nm ** 3 * C), np.float32)
    adj = np.zeros((B, nm * nm), np.float32)
    g = np.zeros((B, nm * nm * 10 * C), np.float32)
    ref_o, ref_t = [], []
    for i, n in enumerate(sizes):
        t, a = rng.uniform(-1, 1, (n, n, n, C)), rng.uniform(-1, 1, (n, n))
        gg = rng.uniform(-1, 1, (n, n, 10 * C))
        T[i, :t.size], adj[i, :a.size], g[i, :gg.size] = t.ravel(), a.ravel(), gg.ravel()
        ref_o.append(pyoracle.einsum10_forward(t, a))
        ref_t.append(pyoracle.einsum10_backward(gg, a))
    n_dev = torch.tensor(sizes, dtype=torch.int32, device="cuda")
    Td, ad, gd = dev(T.reshape(B, nm, nm, nm, C)), dev(adj.reshape(B, nm, nm)), dev(g.reshape(B, nm, nm, 10 * C))
    out = ctx.contract_family_forward(10, Td, ad, n=n_dev).cpu().numpy().reshape(B, -1)
    gT0 = torch.ones((B, nm, nm, nm, C), device="cuda")
    gT = ctx.contract_family_backward(10, gd, ad, gT=gT0.clone(), n=n_dev, beta=1.0).cpu().numpy().reshape(B, -1)
    for i, n in enumerate(sizes):
        assert pyoracle.slab_rel_err(out[i, :n * n * 10 * C].reshape(n, n, 10 * C), ref_o[i], 10) < TOL
        assert pyoracle.slab_rel_err(gT[i, :n ** 3 * C].reshape(n, n, n, C), ref_t[i] + 1.0, 1) < TOL


def test_family_rejects_bad_arguments(ctx):
    T = torch.zeros((1, 4, 4, 4, 2), device="cuda")
    adj = torch.zeros((1, 4, 4), device="cuda")
    with pytest.raises(Exception):
        ctx.contract_family_forward(7, T, adj)          # unknown variant
    with pytest.raises(Exception):
        ctx.contract_family_forward(10, T, None)        # variant 10 needs an adjacency
    with pytest.raises(Exception):
        ctx.contract_family_forward(4, T, out_scale=0.5)
