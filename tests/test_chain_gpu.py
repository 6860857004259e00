"""The chained C-ABI entry points (`ccn_gather_contract18_*`, `ccn_level_*`: one call per stage of a CCN level,
SMP_beta.h:588-616) give exactly what the single entry points give when called in sequence (those are checked against the
oracle in their own tests), and the level chain also matches the oracle's reference chain on the golden fixture."""
import os

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import graphflow_b200

    c = graphflow_b200.Context(0)
    yield c
    c.close()


def dev(x, dt=np.float32):
    return torch.from_numpy(np.ascontiguousarray(x, dt)).cuda()


def test_level_chain_matches_golden_reference_chain(ctx):
    """tests/golden/level_n6_c4.npz: contraction -> Reshape2D -> MatMul(K) -> +bias -> LeakyReLU of the compiled reference."""
    g = np.load(os.path.join(GOLDEN, "level_n6_c4.npz"))
    T, adj, K, b = dev(g["T"][None]), dev(g["adj"][None]), dev(g["K"]), dev(g["bias"])
    X, Y, Z = ctx.level_forward(T, adj, K, b)
    N, Co = g["adj"].shape[0], g["K"].shape[1]
    assert np.abs(X[0].cpu().numpy().reshape(N, N, -1) - g["contracted"]).max() <= 1e-4 * np.abs(g["contracted"]).max()
    assert np.abs(Z.cpu().numpy().reshape(N, N, Co) - g["Z"]).max() <= 1e-4 * np.abs(g["Z"]).max()
    gT, gK, gb = ctx.level_backward(dev(g["gZ"].reshape(N * N, Co)), X, Y, K, b, adj)
    for got, want in ((gT[0], g["gT"]), (gK, g["gK"]), (gb, g["gb"])):
        assert np.abs(got.cpu().numpy() - want).max() <= 1e-4 * np.abs(want).max()


@pytest.mark.parametrize("ragged", [False, True])
def test_level_chain_equals_the_single_calls(ctx, ragged):
    B, N, Ci, Co = 5, 12, 32, 32
    gen = torch.Generator(device="cuda").manual_seed(3)
    T = torch.rand((B, N, N, N, Ci), device="cuda", generator=gen) * 2 - 1
    adj = (torch.rand((B, N, N), device="cuda", generator=gen) < 0.3).float()
    K = (torch.rand((18 * Ci, Co), device="cuda", generator=gen) - 0.5) * 0.1
    b = torch.rand((Co,), device="cuda", generator=gen) - 0.5
    gZ = torch.rand((B * N * N, Co), device="cuda", generator=gen) - 0.5
    n = torch.tensor([12, 7, 12, 3, 9], dtype=torch.int32, device="cuda") if ragged else None
    if ragged:  # compact instances: rows past n^2 of an instance carry no data and must get a zero gradient
        rows = torch.arange(N * N, device="cuda")[None, :] < (n.long() ** 2)[:, None]
        gZ = gZ * rows.reshape(-1, 1)
    X, Y, Z = ctx.level_forward(T, adj, K, b, n=n)
    X1 = torch.zeros_like(X)
    ctx.contract18_forward(T, adj, out=X1.view(B, N, N, 18 * Ci), n=n)
    Y1, Z1 = ctx.mix_forward(X1.view(B * N * N, 18 * Ci), K, b)
    assert torch.equal(X, X1) and torch.equal(Y, Y1) and torch.equal(Z, Z1)
    gT, gK, gb = ctx.level_backward(gZ, X, Y, K, b, adj, n=n)
    gX1, gK1, gb1 = ctx.mix_backward(X1.view(B * N * N, 18 * Ci), K, gZ, bias=b, Y=Y1)
    gT1 = ctx.contract18_backward(gX1.view(B, N, N, 18 * Ci), adj, n=n)
    if ragged:
        sel = (torch.arange(N ** 3, device="cuda")[None, :] < (n.long() ** 3)[:, None]).reshape(B, N ** 3, 1)
        gT, gT1 = gT.reshape(B, N ** 3, Ci) * sel, gT1.reshape(B, N ** 3, Ci) * sel
    assert torch.equal(gT, gT1)
    assert ((gK - gK1).abs().max() <= 1e-5 * gK1.abs().max()).item()      # split-K atomics: not bitwise
    assert ((gb - gb1).abs().max() <= 1e-5 * gb1.abs().max()).item()


def test_gather_contract18_equals_promote_then_contract(ctx):
    rng = np.random.default_rng(8)
    B, n_max, C, m = 4, 8, 16, 6
    W = 10                                                   # level l-1 tensors, [m, m, C] each, packed back to back
    f = dev(rng.uniform(-1, 1, (W, m, m, C))).reshape(-1)
    f_off = dev(rng.integers(0, W, (B * n_max,)) * (m * m * C), np.int64)
    mm = torch.full((B * n_max,), m, dtype=torch.int32, device="cuda")
    pos = dev(rng.integers(-1, m, (B * n_max * n_max,)), np.int32)
    adj = (torch.rand((B, n_max, n_max), device="cuda") < 0.4).float()
    out = ctx.gather_contract18_forward(f, f_off, mm, pos, adj, n_max, C)
    T = ctx.promote_forward(f, f_off, mm, pos, n_max, C)
    assert torch.equal(out, ctx.contract18_forward(T, adj))
    gout = torch.rand((B, n_max, n_max, 18 * C), device="cuda") - 0.5
    gf = torch.zeros_like(f)
    ctx.gather_contract18_backward(gout, adj, f_off, mm, pos, gf)
    gf1 = torch.zeros_like(f)
    ctx.promote_backward(ctx.contract18_backward(gout, adj), f_off, mm, pos, gf1)
    assert ((gf - gf1).abs().max() <= 1e-5 * gf1.abs().max()).item()      # atomic scatter-add: not bitwise
    assert gf1.abs().max().item() > 0


def test_allreduce_grads_over_raw_nccl_communicators():
    """`ccn_allreduce_grads` on two GPUs (skipped on a single-GPU box): profiles/nccl_abi_probe.py under torchrun."""
    import json
    import subprocess
    import sys

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29541", os.path.join(root, "profiles", "nccl_abi_probe.py")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    assert res["allreduce_sum_ok"] and res["null_comm_rejected"]
