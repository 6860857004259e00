"""Edge cases through the C-ABI: empty batches (a no-op, not an error), single-vertex receptive fields (n = 1), a ragged
batch that mixes n = 1 with the largest fused size, the first size beyond the fused kernels (n = 33: generic path), and
argument errors that must be reported, not crash."""
import numpy as np
import pytest
import torch

from oracle import pyoracle

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


def test_empty_batches_are_no_ops(ctx):
    N, C = 6, 8
    z = lambda *s: torch.empty(s, device="cuda")  # noqa: E731
    assert ctx.contract18_forward(z(0, N, N, N, C), z(0, N, N)).shape == (0, N, N, 18 * C)
    assert ctx.contract18_backward(z(0, N, N, 18 * C), z(0, N, N)).shape == (0, N, N, N, C)
    assert ctx.contract50_forward(z(0, N, N, N, C), z(0, N, N)).shape == (0, N, N, 50 * C)
    assert ctx.contract_family_forward(10, z(0, N, N, N, C), z(0, N, N)).shape == (0, N, N, 10 * C)
    assert ctx.contract_family_backward(18, z(0, N, N, 18 * C), z(0, N, N), keep_mask=[True] * 9 + [False] * 9).shape[0] == 0
    Y, Z = ctx.mix_forward(z(0, 18 * C), torch.rand((18 * C, C), device="cuda"), torch.rand((C,), device="cuda"))
    assert Y.shape == (0, C) and Z.shape == (0, C)
    launches = ctx.kernel_launches
    ctx.contract18_forward(z(0, N, N, N, C), z(0, N, N))
    assert ctx.kernel_launches == launches          # nothing was launched


@pytest.mark.parametrize("C", [4, 64])
def test_single_vertex_fields(ctx, C):
    """n = 1: T is one number per channel; every slab is that number times an adjacency scalar."""
    rng = np.random.default_rng(C)
    B = 5
    T = rng.uniform(-1, 1, (B, 1, 1, 1, C)).astype(np.float32)
    adj = rng.uniform(0.5, 2, (B, 1, 1)).astype(np.float32)
    gout = rng.uniform(-1, 1, (B, 1, 1, 18 * C)).astype(np.float32)
    out = ctx.contract18_forward(dev(T), dev(adj)).cpu().numpy()
    gT = ctx.contract18_backward(dev(gout), dev(adj)).cpu().numpy()
    for i in range(B):
        assert pyoracle.slab_rel_err(out[i], pyoracle.einsum18_forward(T[i], adj[i]), 18) < TOL
        assert pyoracle.slab_rel_err(gT[i], pyoracle.einsum18_backward(gout[i], adj[i]), 1) < TOL


@pytest.mark.parametrize("nm,C", [(32, 64), (33, 8)])
def test_ragged_batch_from_one_to_the_maximum(ctx, nm, C):
    """Sizes 1 .. n_max in one launch; n_max = 32 is the largest fused size, 33 the first that takes the generic path."""
    rng = np.random.default_rng(nm)
    sizes = [1, nm, 2, nm - 1, 17]
    B = len(sizes)
    T = np.zeros((B, nm ** 3 * C), np.float32)
    adj = np.zeros((B, nm * nm), np.float32)
    gout = np.zeros((B, nm * nm * 18 * C), np.float32)
    refs = []
    for i, n in enumerate(sizes):
        t = rng.uniform(-1, 1, (n, n, n, C))
        a = (rng.random((n, n)) < 0.2).astype(np.float64)
        a = np.maximum(a, a.T)
        np.fill_diagonal(a, 1.0)
        g = rng.uniform(-1, 1, (n, n, 18 * C))
        T[i, :t.size], adj[i, :a.size], gout[i, :g.size] = t.ravel(), a.ravel(), g.ravel()
        refs.append((pyoracle.einsum18_forward(t, a), pyoracle.einsum18_backward(g, a)))
    nd = torch.tensor(sizes, dtype=torch.int32, device="cuda")
    out = ctx.contract18_forward(dev(T.reshape(B, nm, nm, nm, C)), dev(adj.reshape(B, nm, nm)), n=nd).cpu().numpy().reshape(B, -1)
    gT = ctx.contract18_backward(dev(gout.reshape(B, nm, nm, 18 * C)), dev(adj.reshape(B, nm, nm)), n=nd).cpu().numpy().reshape(B, -1)
    for i, n in enumerate(sizes):
        assert pyoracle.slab_rel_err(out[i, :n * n * 18 * C].reshape(n, n, 18 * C), refs[i][0], 18) < TOL
        assert pyoracle.slab_rel_err(gT[i, :n ** 3 * C].reshape(n, n, n, C), refs[i][1], 1) < TOL


def test_argument_errors_are_reported(ctx):
    import graphflow_b200

    T = torch.zeros((1, 4, 4, 4, 2), device="cuda")
    adj = torch.zeros((1, 4, 4), device="cuda")
    with pytest.raises(graphflow_b200.CCNError):
        ctx.contract18_forward(T, adj, adj_mode=7)
    with pytest.raises((TypeError, ValueError)):
        ctx.contract18_forward(T.double(), adj)
    with pytest.raises((TypeError, ValueError)):
        ctx.contract18_forward(T.cpu(), adj)
    with pytest.raises(graphflow_b200.CCNError):                       # one instance must stay below 2^31 elements
        big = torch.empty((1,), device="cuda")
        ctx._rc(ctx.lib.ccn_contract18_forward(ctx.h, big.data_ptr(), None, adj.data_ptr(), big.data_ptr(), None, 2048, 512, 1,
                                               2048 ** 3 * 512, 2048 * 2048, 18 * 2048 * 2048 * 512, 0, None))


def test_empty_instance_inside_a_ragged_batch(ctx):
    """n_i = 0 in the middle of a batch with more instances than scratch slots: the empty instance's slot generation must
    still advance, so the instances that recycle its slot neither wait nor raise the (sticky) sibling-timeout flag."""
    rng = np.random.default_rng(3)
    nm, C, B = 32, 64, 160
    sizes = rng.integers(1, nm + 1, B).astype(np.int32)
    sizes[[0, 5, 41, 90]] = 0
    T = rng.uniform(-1, 1, (B, nm ** 3 * C)).astype(np.float32)
    adj = np.zeros((B, nm * nm), np.float32)
    gout = rng.uniform(-1, 1, (B, nm * nm * 18 * C)).astype(np.float32)
    for i, n in enumerate(sizes):
        a = (rng.random((n, n)) < 0.2).astype(np.float32)
        a = np.maximum(a, a.T)
        np.fill_diagonal(a, 1.0)
        adj[i, :n * n] = a.ravel()
    nd = torch.from_numpy(sizes).cuda()
    out = ctx.contract18_forward(dev(T.reshape(B, nm, nm, nm, C)), dev(adj.reshape(B, nm, nm)), n=nd)
    gT = ctx.contract18_backward(dev(gout.reshape(B, nm, nm, 18 * C)), dev(adj.reshape(B, nm, nm)), n=nd)
    assert ctx.fused_error_flag() == 0
    out, gT = out.cpu().numpy().reshape(B, -1), gT.cpu().numpy().reshape(B, -1)
    for i in (1, 6, 42, 91, B - 1):                                   # the instances right after the empty ones, and the last
        n = int(sizes[i])
        t, a, g = T[i, :n ** 3 * C].reshape(n, n, n, C), adj[i, :n * n].reshape(n, n), gout[i, :n * n * 18 * C].reshape(n, n, 18 * C)
        assert pyoracle.slab_rel_err(out[i, :n * n * 18 * C].reshape(n, n, 18 * C), pyoracle.einsum18_forward(t, a), 18) < TOL
        assert pyoracle.slab_rel_err(gT[i, :n ** 3 * C].reshape(n, n, n, C), pyoracle.einsum18_backward(g, a), 1) < TOL


def test_two_streams_on_one_context_are_ordered(ctx):
    """The context's scratch is shared, so calls on different streams must not overlap on the device: alternate two
    streams with no host synchronisation in between and compare every result with the single-stream one."""
    from tests.util import random_instance

    rng = np.random.default_rng(9)
    n, C, B = 32, 64, 48
    inst = [random_instance(n, C, rng) for _ in range(4)]
    T = dev(np.stack([inst[i % 4][0] for i in range(B)]))
    adj = dev(np.stack([inst[i % 4][1] for i in range(B)]))
    gout = dev(np.stack([inst[i % 4][2] for i in range(B)]))
    want_out = ctx.contract18_forward(T, adj).clone()
    want_gT = ctx.contract18_backward(gout, adj).clone()
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs, gTs = [], []
    for k in range(6):
        outs.append(ctx.contract18_forward(T, adj, stream=s1 if k % 2 == 0 else s2))
        gTs.append(ctx.contract18_backward(gout, adj, stream=s2 if k % 2 == 0 else s1))
    torch.cuda.synchronize()
    assert ctx.fused_error_flag() == 0
    for o, g in zip(outs, gTs):
        assert torch.equal(o, want_out) and torch.equal(g, want_gT)


def test_misaligned_gradient_takes_the_generic_kernels(ctx):
    """A gout base that is not 16-byte aligned must not reach the fused kernel's 16-byte cp.async reads."""
    from tests.util import random_instance

    rng = np.random.default_rng(11)
    n, C = 32, 64
    T, a, g = random_instance(n, C, rng)
    buf = torch.zeros(g.size + 4, device="cuda")
    shifted = buf[1:1 + g.size]                                        # 4-byte aligned only
    shifted.copy_(dev(g).reshape(-1))
    gT = ctx.contract18_backward(shifted, dev(a[None]), n_max=n, C=C, batch=1).cpu().numpy()
    assert pyoracle.slab_rel_err(gT[0], pyoracle.einsum18_backward(g, a), 1) < TOL
    tb = torch.zeros(T.size + 4, device="cuda")
    tsh = tb[1:1 + T.size]
    tsh.copy_(dev(T).reshape(-1))
    out = ctx.contract18_forward(tsh, dev(a[None]), n_max=n, C=C, batch=1).cpu().numpy()
    assert pyoracle.slab_rel_err(out[0], pyoracle.einsum18_forward(T, a), 18) < TOL


def test_frozen_context_refuses_to_grow():
    import graphflow_b200

    c = graphflow_b200.Context(0)
    try:
        small = torch.rand((2, 8, 8, 8, 8), device="cuda")
        adj = torch.eye(8, device="cuda").repeat(2, 1, 1)
        c.contract18_forward(small, adj)
        c.set_frozen(True)
        c.contract18_forward(small, adj)                               # same scratch: fine
        with pytest.raises(graphflow_b200.CCNError):
            c.contract18_forward(torch.rand((64, 32, 32, 32, 64), device="cuda"), torch.eye(32, device="cuda").repeat(64, 1, 1))
        c.set_frozen(False)
        c.contract18_forward(torch.rand((64, 32, 32, 32, 64), device="cuda"), torch.eye(32, device="cuda").repeat(64, 1, 1))
    finally:
        c.close()
