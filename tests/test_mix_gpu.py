"""GPU parity tests for the feature mix (Reshape2D + MatMul [+ VectorAddTensor + LeakyReLU3D]) through the C-ABI.

Mirrors tests/test_MatMul_gpu.cu (forward compare; backward with pre-loaded non-zero input gradients, :103-116).
Tolerance 1e-4 of the output's max-abs (north_star), checked against the fp64 oracle; the integer-valued golden
vector must match exactly."""
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


def rel(x, ref):
    ref = np.asarray(ref, np.float64)
    den = np.abs(ref).max()
    return np.abs(np.asarray(x, np.float64) - ref).max() / (den if den > 0 else 1.0)


def test_matmul_golden_exact(ctx):
    g = np.load(os.path.join(GOLDEN, "matmul_20x36x5.npz"))
    Y, _ = ctx.mix_forward(dev(g["X"]), dev(g["W"]))
    assert np.array_equal(Y.cpu().numpy(), g["Y"].astype(np.float32))
    gX, gW = dev(g["gX0"]), dev(g["gW0"])
    ctx.mix_backward(dev(g["X"]), dev(g["W"]), dev(g["gY"]), gX=gX, gW=gW, beta_x=1.0)
    assert np.array_equal(gX.cpu().numpy(), g["gX"].astype(np.float32))
    assert np.array_equal(gW.cpu().numpy(), g["gW"].astype(np.float32))


@pytest.mark.parametrize("M,K,P", [(1024, 1152, 64), (576, 576, 32), (300, 72, 4), (1, 18, 1), (2048 + 17, 1152, 64)])
def test_mix_forward_backward_vs_oracle(ctx, M, K, P):
    rng = np.random.default_rng(M + K + P)
    X = rng.uniform(-1, 1, (M, K))
    W = rng.uniform(-0.2, 0.2, (K, P))
    bias = rng.uniform(-0.5, 0.5, (P,))
    gZ = rng.uniform(-1, 1, (M, P))
    orc = pyoracle.COracle("f64")
    Y_ref = orc.matmul_forward(X, W)
    Z_ref = orc.bias_lrelu_forward(Y_ref, bias)
    gY_ref, gb_ref = orc.bias_lrelu_backward(Y_ref, bias, gZ)
    gX_ref, gW_ref = orc.matmul_backward(X, W, gY_ref)

    Y, Z = ctx.mix_forward(dev(X), dev(W), dev(bias))
    assert rel(Y.cpu().numpy(), Y_ref) < TOL and rel(Z.cpu().numpy(), Z_ref) < TOL
    # use the oracle's pre-activation so the lrelu mask is identical (entries within fp32 rounding of 0 could flip)
    gX, gW, gb = ctx.mix_backward(dev(X), dev(W), dev(gZ), bias=dev(bias), Y=dev(Y_ref))
    assert rel(gX.cpu().numpy(), gX_ref) < TOL
    assert rel(gW.cpu().numpy(), gW_ref) < TOL
    assert rel(gb.cpu().numpy(), gb_ref) < TOL


def test_level_chain_golden(ctx):
    """contraction -> mix -> +bias -> LeakyReLU and back, against the reference chain (SMP_beta.h:596-616)."""
    g = np.load(os.path.join(GOLDEN, "level_n6_c4.npz"))
    N, C, Cout = 6, 4, 4
    contracted = ctx.contract18_forward(dev(g["T"][None]), dev(g["adj"][None]))
    assert rel(contracted[0].cpu().numpy(), g["contracted"]) < TOL
    X = contracted.reshape(N * N, 18 * C)
    Y, Z = ctx.mix_forward(X, dev(g["K"]), dev(g["bias"]))
    assert rel(Z.cpu().numpy().reshape(N, N, Cout), g["Z"]) < TOL
    gX, gK, gb = ctx.mix_backward(X, dev(g["K"]), dev(g["gZ"].reshape(N * N, Cout)), bias=dev(g["bias"]), Y=Y)
    gT = ctx.contract18_backward(gX.reshape(1, N, N, 18 * C), dev(g["adj"][None]))
    assert rel(gb.cpu().numpy(), g["gb"]) < TOL
    assert rel(gK.cpu().numpy(), g["gK"]) < TOL
    assert rel(gT[0].cpu().numpy(), g["gT"]) < TOL
