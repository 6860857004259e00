"""The fused level entry points against the ORACLE (not against the unfused calls): `ccn_gather_contract18_*` read / scatter the
promotion inside the fused contraction kernels (fusion step 2 of SURVEY 7.2: the stacked T / gT never exists),
`ccn_gather_level_*` chain them with the feature mix, `ccn_gather_level_forward_backward_host` is the same from host arrays.
Reference chain: SMP_beta.h:588-616 (MatTensorMul, TensorMatMul, StackTensor3D, RisiContraction_18, Reshape2D, MatMul, Reshape3D,
VectorAddTensor, LeakyReLU3D) through tests/util.oracle_gather_level; the small case also runs the compiled reference operators."""
import numpy as np
import pytest
import torch

from oracle import pyoracle
from tests.util import level_tables, molecular_adjacency, oracle_gather_level

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


def make_graphs(rng, graphs, V, C, n_max, full):
    """`graphs` graphs of V vertices.  full: every receptive field (both levels) is the whole vertex set in a random order
    (n = V everywhere, T completely dense).  else: random subsets (ragged sizes, absent members -> zero fill)."""
    f_off, m, pos, n, adj, fbounds, ibounds = [], [], [], [], [], [0], [0]
    base = 0
    for _ in range(graphs):
        if full:
            prev = [list(rng.permutation(V)) for _ in range(V)]
            cur = [list(rng.permutation(V)) for _ in range(V)]
        else:
            prev = [list(rng.permutation(V)[:rng.integers(1, V + 1)]) for _ in range(V)]
            cur = [list(rng.permutation(V)[:rng.integers(1, min(V, n_max) + 1)]) for _ in range(V)]
        fo, mm, pp, nn, fsz = level_tables(prev, cur, C, n_max, base)
        base += fsz
        f_off.append(fo), m.append(mm), pos.append(pp), n.append(nn)
        A = molecular_adjacency(V, rng)
        for v in range(V):
            a = np.zeros(n_max * n_max, np.float32)
            idx = np.asarray(cur[v])
            a[:len(idx) ** 2] = A[np.ix_(idx, idx)].ravel()        # reduced adjacency of phi_l(v), compact (SMP_beta.h:505-526)
            adj.append(a)
        fbounds.append(base)
        ibounds.append(ibounds[-1] + V)
    return (np.concatenate(f_off), np.concatenate(m), np.concatenate(pos), np.concatenate(n), np.stack(adj), base,
            np.asarray(fbounds, np.int64), np.asarray(ibounds, np.int64))


def check_level(ctx, rng, graphs, V, C, Co, n_max, full, host=False):
    f_off, m, pos, n, adj, fsz, fb, ib = make_graphs(rng, graphs, V, C, n_max, full)
    B = len(n)
    f = rng.uniform(-1, 1, fsz).astype(np.float32)
    K = rng.uniform(-0.1, 0.1, (18 * C, Co)).astype(np.float32)
    bias = rng.uniform(-0.5, 0.5, Co).astype(np.float32)
    gZ = rng.uniform(-1, 1, (B, n_max * n_max, Co)).astype(np.float32)
    rows = np.arange(n_max * n_max)[None, :] < (n.astype(np.int64) ** 2)[:, None]
    gZ *= rows[:, :, None]                                              # rows past n^2 of a compact instance carry no gradient
    Xs, Zs, gf_ref, gK_ref, gb_ref = oracle_gather_level(f, f_off, m, pos, n, adj, K, bias, gZ, n_max, C)
    if host:
        Z = torch.empty((B * n_max * n_max, Co)).pin_memory()
        gf = torch.empty(fsz).pin_memory()
        gK, gb = torch.empty((18 * C, Co)), torch.empty(Co)
        t = lambda x, dt=np.float32: torch.from_numpy(np.ascontiguousarray(x, dt))  # noqa: E731
        ctx.gather_level_forward_backward_host(t(f), t(fb, np.int64), t(ib, np.int64), t(f_off, np.int64), t(m, np.int32),
                                               t(pos, np.int32), t(adj), t(K), t(bias), t(gZ.reshape(-1, Co)), Z, gf, gK, gb, n_max)
        X = None
    else:
        nd = None if full else dev(n, np.int32)
        X, Y, Z = ctx.gather_level_forward(dev(f), dev(f_off, np.int64), dev(m, np.int32), dev(pos, np.int32), dev(adj), dev(K),
                                           dev(bias), n_max, n=nd)
        gf = torch.zeros(fsz, device="cuda")
        gf, gK, gb = ctx.gather_level_backward(dev(gZ.reshape(-1, Co)), X, Y, dev(K), dev(bias), dev(adj), dev(f_off, np.int64),
                                               dev(m, np.int32), dev(pos, np.int32), gf, n_max, n=nd)
        assert ctx.fused_error_flag() == 0
        X = X.cpu().numpy()
    Z = Z.cpu().numpy().reshape(B, n_max * n_max, Co)
    for i in range(B):
        ni = int(n[i])
        if X is not None:
            assert pyoracle.slab_rel_err(X[i, :ni * ni].reshape(ni, ni, 18 * C), Xs[i].reshape(ni, ni, 18 * C), 18) < TOL, i
        assert np.abs(Z[i, :ni * ni] - Zs[i]).max() <= TOL * np.abs(Zs[i]).max(), i
    for got, want, what in ((gf, gf_ref, "gf"), (gK, gK_ref, "gK"), (gb, gb_ref, "gbias")):
        got = got.cpu().numpy().astype(np.float64)
        assert np.abs(got - want).max() <= TOL * np.abs(want).max(), what


def test_fused_level_at_the_headline_shape(ctx):
    """N = 32, C = 64 -> 64, every field full (dense T): one graph of 32 vertices = 32 instances."""
    check_level(ctx, np.random.default_rng(1), 1, 32, 64, 64, 32, full=True)


@pytest.mark.parametrize("V,C,Co,n_max", [(12, 32, 32, 12), (20, 64, 32, 16), (9, 8, 16, 9), (32, 16, 8, 32)])
def test_fused_level_ragged_fields(ctx, V, C, Co, n_max):
    """Ragged receptive fields with absent members (zero fill), several graphs, narrow and wide channels."""
    check_level(ctx, np.random.default_rng(V * 7 + C), 3, V, C, Co, n_max, full=False)


def test_fused_level_from_host_arrays(ctx):
    """ccn_gather_level_forward_backward_host: 5 graphs of 32 vertices, chunked (CCN_LEVEL_CHUNK default 256 -> one chunk; the
    second call forces 2-graph chunks through the three-stream ring)."""
    import os

    check_level(ctx, np.random.default_rng(2), 2, 32, 64, 64, 32, full=True, host=True)
    os.environ["CCN_LEVEL_CHUNK"] = "40"
    try:
        check_level(ctx, np.random.default_rng(3), 5, 16, 32, 32, 16, full=True, host=True)
    finally:
        del os.environ["CCN_LEVEL_CHUNK"]


def test_fused_gather_matches_compiled_reference_operators(ctx):
    """Small case against the UNMODIFIED reference operators (MatTensorMul + TensorMatMul promotion, then the level chain of
    oracle/ref_shim.cpp): n = 6, C = 8 (a fused-kernel shape), one instance."""
    if not pyoracle.ref_available("f64"):
        pytest.skip("oracle/_ref not built")
    ref = pyoracle.RefOracle("f64")
    rng = np.random.default_rng(5)
    n, C, Co, mprev = 6, 8, 8, 7
    fs = [rng.integers(-3, 4, (mprev, mprev, C)).astype(np.float64) for _ in range(n)]
    ps = [rng.integers(-1, mprev, n).astype(np.int32) for _ in range(n)]
    adj = molecular_adjacency(n, rng).astype(np.float64)
    K = rng.integers(-2, 3, (18 * C, Co)).astype(np.float64)
    bias = rng.integers(-2, 3, Co).astype(np.float64)
    gZ = rng.integers(-2, 3, (n, n, Co)).astype(np.float64)
    T = np.stack([ref.promote(fs[a], ps[a]) for a in range(n)])
    contracted, Z, gT, gK, gb = ref.level_forward_backward(T, adj, K, bias, gZ)
    gfs = [ref.promote(fs[a], ps[a], gQ=gT[a])[1] for a in range(n)]
    f = np.concatenate([x.ravel() for x in fs])
    f_off = np.arange(n, dtype=np.int64) * mprev * mprev * C
    m = np.full(n, mprev, np.int32)
    pos = np.concatenate(ps)
    X, Y, Zd = ctx.gather_level_forward(dev(f), dev(f_off, np.int64), dev(m, np.int32), dev(pos, np.int32), dev(adj[None]), dev(K),
                                        dev(bias), n)
    gf = torch.zeros(f.size, device="cuda")
    gf, gKd, gbd = ctx.gather_level_backward(dev(gZ.reshape(-1, Co)), X, Y, dev(K), dev(bias), dev(adj[None]), dev(f_off, np.int64),
                                             dev(m, np.int32), dev(pos, np.int32), gf, n)
    # integer-valued inputs: the contraction is exact in fp32; the 3xTF32 mix is exact on small integers as well
    assert np.array_equal(X.cpu().numpy().reshape(n, n, 18 * C), contracted)
    assert np.abs(Zd.cpu().numpy().reshape(n, n, Co) - Z).max() <= 1e-5 * np.abs(Z).max()
    want_gf = np.concatenate([x.ravel() for x in gfs])
    for got, want in ((gf, want_gf), (gKd, gK), (gbd, gb)):
        assert np.abs(got.cpu().numpy() - want).max() <= 1e-5 * np.abs(want).max()


def test_stack_of_levels_from_host_arrays(ctx):
    """ccn_gather_levels_forward_backward_host: two levels with DIFFERENT receptive-field orders per level, device resident in
    between, 3 graphs of 16 vertices cut into 2-graph chunks, against the oracle chain run level by level."""
    import os

    rng = np.random.default_rng(12)
    G, V, C, Lv = 3, 16, 32, 2
    nn = V * V
    fields = [[[list(rng.permutation(V)) for _ in range(V)] for _ in range(G)] for _ in range(Lv + 1)]   # [level][graph][vertex]
    tabs = []
    for l in range(1, Lv + 1):
        fo, mm, pp, adj = [], [], [], []
        for g in range(G):
            a, b, c, n, fsz = level_tables(fields[l - 1][g], fields[l][g], C, V, g * V * nn * C)
            assert fsz == V * nn * C
            fo.append(a), mm.append(b), pp.append(c)
            A = molecular_adjacency(V, rng)
            adj += [A[np.ix_(fields[l][g][v], fields[l][g][v])].ravel() for v in range(V)]
        tabs.append((np.concatenate(fo), np.concatenate(mm), np.concatenate(pp), np.stack(adj).astype(np.float32)))
    B = G * V
    f = rng.uniform(-1, 1, B * nn * C).astype(np.float32)
    Ks = [rng.uniform(-0.1, 0.1, (18 * C, C)).astype(np.float32) for _ in range(Lv)]
    bs = [rng.uniform(-0.5, 0.5, C).astype(np.float32) for _ in range(Lv)]
    gZ = rng.uniform(-1, 1, (B, nn, C)).astype(np.float32)
    n_all = np.full(B, V, np.int32)
    # oracle: forward level by level, then backward in reverse (the gf of level 2 is the gZ of level 1)
    acts, saved = [f.astype(np.float64)], []
    for l in range(Lv):
        fo, mm, pp, adj = tabs[l]
        _, Zs, _, _, _ = oracle_gather_level(acts[-1], fo, mm, pp, n_all, adj, Ks[l], bs[l], np.zeros((B, nn, C)), V, C)
        acts.append(np.concatenate([z.ravel() for z in Zs]))
    g_cur, gK_ref, gb_ref = gZ.astype(np.float64), [None] * Lv, [None] * Lv
    for l in reversed(range(Lv)):
        fo, mm, pp, adj = tabs[l]
        _, _, gf_l, gK_ref[l], gb_ref[l] = oracle_gather_level(acts[l], fo, mm, pp, n_all, adj, Ks[l], bs[l], g_cur.reshape(B, nn, C), V, C)
        g_cur = gf_l
    t = lambda x, dt=np.float32: torch.from_numpy(np.ascontiguousarray(x, dt))  # noqa: E731
    Z, gf = torch.empty((B * nn, C)).pin_memory(), torch.empty(f.size).pin_memory()
    gK, gb = [torch.empty((18 * C, C)) for _ in range(Lv)], [torch.empty(C) for _ in range(Lv)]
    os.environ["CCN_LEVEL_CHUNK"] = str(2 * V)
    try:
        ctx.gather_levels_forward_backward_host(
            t(f), t(np.arange(G + 1) * V * nn * C, np.int64), t(np.arange(G + 1) * V, np.int64), [t(x[0], np.int64) for x in tabs],
            [t(x[1], np.int32) for x in tabs], [t(x[2], np.int32) for x in tabs], [t(x[3]) for x in tabs], [t(k) for k in Ks],
            [t(b) for b in bs], t(gZ.reshape(-1, C)), Z, gf, gK, gb, V)
    finally:
        del os.environ["CCN_LEVEL_CHUNK"]
    want_Z = acts[-1].reshape(B * nn, C)
    assert np.abs(Z.numpy() - want_Z).max() <= TOL * np.abs(want_Z).max()
    assert np.abs(gf.numpy() - g_cur).max() <= TOL * np.abs(g_cur).max()
    for l in range(Lv):
        assert np.abs(gK[l].numpy() - gK_ref[l]).max() <= TOL * np.abs(gK_ref[l]).max(), l
        assert np.abs(gb[l].numpy() - gb_ref[l]).max() <= TOL * np.abs(gb_ref[l]).max(), l


def test_stack_of_levels_with_device_readout_from_host_arrays(ctx):
    """ccn_gather_levels_readout_forward_backward_host: two levels + the read-out head + loss on the device, 3 graphs of 16
    vertices in 2-graph chunks, against the oracle chain (levels as above; read-out = ShrinkTensor -> LeakyReLU -> SumVectors ->
    InnerProduct -> SquaredLoss, SMP_beta.h:620-639, restated in numpy fp64)."""
    import os

    rng = np.random.default_rng(21)
    G, V, C, Lv = 3, 16, 32, 2
    nn = V * V
    fields = [[[list(rng.permutation(V)) for _ in range(V)] for _ in range(G)] for _ in range(Lv + 1)]
    tabs = []
    for l in range(1, Lv + 1):
        fo, mm, pp, adj = [], [], [], []
        for g in range(G):
            a, b, c, n, fsz = level_tables(fields[l - 1][g], fields[l][g], C, V, g * V * nn * C)
            fo.append(a), mm.append(b), pp.append(c)
            A = molecular_adjacency(V, rng)
            adj += [A[np.ix_(fields[l][g][v], fields[l][g][v])].ravel() for v in range(V)]
        tabs.append((np.concatenate(fo), np.concatenate(mm), np.concatenate(pp), np.stack(adj).astype(np.float32)))
    B = G * V
    f = rng.uniform(-1, 1, B * nn * C).astype(np.float32)
    Ks = [rng.uniform(-0.1, 0.1, (18 * C, C)).astype(np.float32) for _ in range(Lv)]
    bs = [rng.uniform(-0.5, 0.5, C).astype(np.float32) for _ in range(Lv)]
    W = rng.uniform(-0.05, 0.05, C).astype(np.float32)
    target = rng.uniform(-1, 1, G).astype(np.float32)
    n_all = np.full(B, V, np.int32)
    acts = [f.astype(np.float64)]
    for l in range(Lv):
        fo, mm, pp, adj = tabs[l]
        _, Zs, _, _, _ = oracle_gather_level(acts[-1], fo, mm, pp, n_all, adj, Ks[l], bs[l], np.zeros((B, nn, C)), V, C)
        acts.append(np.concatenate([z.ravel() for z in Zs]))
    ZL = acts[-1].reshape(G, V, nn, C)
    shr = ZL.sum(2)                                                   # ShrinkTensor
    vf = np.where(shr > 0, shr, 0.01 * shr)                           # LeakyReLU
    gfeat = vf.sum(1)                                                 # SumVectors
    pred = gfeat @ W.astype(np.float64)                               # InnerProduct
    loss = 0.5 * (pred - target) ** 2                                 # SquaredLoss
    d = pred - target
    gW_ref = (d[:, None] * gfeat).sum(0)
    gshr = d[:, None, None] * W[None, None, :] * np.where(shr > 0, 1.0, 0.01)
    g_cur = np.broadcast_to(gshr[:, :, None, :], ZL.shape).reshape(B, nn, C).copy()
    gK_ref, gb_ref = [None] * Lv, [None] * Lv
    for l in reversed(range(Lv)):
        fo, mm, pp, adj = tabs[l]
        _, _, gf_l, gK_ref[l], gb_ref[l] = oracle_gather_level(acts[l], fo, mm, pp, n_all, adj, Ks[l], bs[l], g_cur.reshape(B, nn, C), V, C)
        g_cur = gf_l
    t = lambda x, dt=np.float32: torch.from_numpy(np.ascontiguousarray(x, dt))  # noqa: E731
    gf = torch.empty(f.size).pin_memory()
    gK, gb = [torch.empty((18 * C, C)) for _ in range(Lv)], [torch.empty(C) for _ in range(Lv)]
    predict, lossd, gW = torch.empty(G), torch.empty(G), torch.empty(C)
    os.environ["CCN_LEVEL_CHUNK"] = str(2 * V)
    try:
        ctx.gather_levels_readout_forward_backward_host(
            t(f), t(np.arange(G + 1) * V * nn * C, np.int64), t(np.arange(G + 1) * V, np.int64), [t(x[0], np.int64) for x in tabs],
            [t(x[1], np.int32) for x in tabs], [t(x[2], np.int32) for x in tabs], [t(x[3]) for x in tabs], [t(k) for k in Ks],
            [t(b) for b in bs], t(W), t(target), predict, lossd, gf, gK, gb, gW, V)
    finally:
        del os.environ["CCN_LEVEL_CHUNK"]
    assert np.abs(predict.numpy() - pred).max() <= TOL * np.abs(pred).max()
    assert np.abs(lossd.numpy() - loss).max() <= TOL * max(np.abs(loss).max(), 1e-6)
    assert np.abs(gW.numpy() - gW_ref).max() <= TOL * np.abs(gW_ref).max()
    assert np.abs(gf.numpy() - g_cur).max() <= TOL * np.abs(g_cur).max()
    for l in range(Lv):
        assert np.abs(gK[l].numpy() - gK_ref[l]).max() <= TOL * np.abs(gK_ref[l]).max(), l
        assert np.abs(gb[l].numpy() - gb_ref[l]).max() <= TOL * np.abs(gb_ref[l]).max(), l
