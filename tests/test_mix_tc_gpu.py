"""GPU parity of the tcgen05 (3xTF32 split-precision) feature-mix forward against the fp64 oracle and against the
exact-fp32 SIMT kernel, through the C-ABI.  Integer-valued inputs must match exactly (every hi part is exact, every
lo part is zero, fp32 accumulation of integers below 2^24 is exact), like tests/test_MatMul_gpu.cu's recipe."""
import numpy as np
import pytest
import torch

from graphflow_b200 import _lib
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


def rel(x, ref):
    ref = np.asarray(ref, np.float64)
    den = np.abs(ref).max()
    return np.abs(np.asarray(x, np.float64) - ref).max() / (den if den > 0 else 1.0)


SHAPES = [(1024, 1152, 64),        # one instance of the headline shape (N=32, C=64)
          (2048 + 17, 1152, 64),   # ragged last tile
          (640, 72, 16),           # K tail (72 = 2 * 32 + 8), smallest P
          (128 * 150 + 5, 576, 32),  # more tiles than SMs: several tiles per CTA, both accumulators
          (4096, 2304, 128),       # C = 128
          (100, 36, 48),           # M < one tile, K barely above one chunk
          (3000, 288, 8),          # narrow outputs (SMP_omega_physics levels): N padded to 16 inside the MMA
          (777, 144, 4),
          (513, 576, 20)]          # P % 4 == 0 but not a multiple of 16


@pytest.mark.parametrize("M,K,P", SHAPES)
def test_mix_tc_forward_vs_oracle(ctx, M, K, P):
    rng = np.random.default_rng(M + K + P)
    X = rng.uniform(-1, 1, (M, K))
    W = rng.uniform(-0.2, 0.2, (K, P))
    bias = rng.uniform(-0.5, 0.5, (P,))
    orc = pyoracle.COracle("f64")
    Y_ref = orc.matmul_forward(X, W)
    Z_ref = orc.bias_lrelu_forward(Y_ref, bias)
    ctx.set_mix_path(_lib.MIX_TENSOR)
    try:
        before = ctx.kernel_launches
        Y, Z = ctx.mix_forward(dev(X), dev(W), dev(bias))
        torch.cuda.synchronize()
        assert ctx.kernel_launches - before == 2  # weight preparation + the tensor-core GEMM
        Y2, _ = ctx.mix_forward(dev(X), dev(W))   # no epilogue
    finally:
        ctx.set_mix_path(_lib.MIX_AUTO)
    ctx.set_mix_path(_lib.MIX_SIMT)
    Ys, _ = ctx.mix_forward(dev(X), dev(W))
    ctx.set_mix_path(_lib.MIX_AUTO)
    e = rel(Y.cpu().numpy(), Y_ref)
    assert e < TOL and rel(Z.cpu().numpy(), Z_ref) < TOL and rel(Y2.cpu().numpy(), Y_ref) < TOL
    # The split itself is fp32-accurate (dropped lo*lo term: 2^-22); what remains is the tensor core's fp32
    # accumulator, which truncates (round toward zero) once per MMA: K/8 * 3 sequential accumulations give ~1e-5 of
    # max|Y| at K = 1152 (measured 9.6e-6) against 1.2e-6 for the exact-fp32 SIMT kernel.  Bar: 1e-4; guard at 3e-5.
    assert e < 3e-5, (e, rel(Ys.cpu().numpy(), Y_ref))


def test_mix_tc_integer_inputs_exact(ctx):
    rng = np.random.default_rng(7)
    M, K, P = 1600, 720, 48  # tests/test_MatMul_gpu.cu:22-26 shape with P rounded to a multiple of 16
    X = rng.integers(0, 100, (M, K)).astype(np.float64)
    W = rng.integers(0, 100, (K, P)).astype(np.float64)
    ctx.set_mix_path(_lib.MIX_TENSOR)
    try:
        Y, _ = ctx.mix_forward(dev(X), dev(W))
    finally:
        ctx.set_mix_path(_lib.MIX_AUTO)
    assert np.array_equal(Y.cpu().numpy().astype(np.float64), X @ W)


def test_mix_tc_unsupported_shape_is_an_error(ctx):
    import graphflow_b200

    ctx.set_mix_path(_lib.MIX_TENSOR)
    try:
        with pytest.raises(graphflow_b200.CCNError):
            ctx.mix_forward(torch.zeros((8, 18), device="cuda"), torch.zeros((18, 1), device="cuda"))
    finally:
        ctx.set_mix_path(_lib.MIX_AUTO)


@pytest.mark.parametrize("M,K,P", [(1024, 1152, 64), (2048 + 17, 1152, 64), (640, 72, 32), (128 * 310 + 5, 576, 32), (100, 36, 64),
                                   (8192 + 77, 1152, 64), (5000, 72, 32), (4500, 576, 16), (6000, 288, 8), (4100, 144, 4),
                                   (900, 288, 12)])
def test_mix_tc_grad_x_vs_oracle(ctx, M, K, P):
    """gX = beta gX + (gZ * lrelu'(Y + b)) W^T on the tensor cores (resident split gY tile in TMEM, W^T streamed); for
    M >= 4096 also gW += X^T gY and gbias += colsum(gY) on the tensor cores (split over the rows, atomics)."""
    rng = np.random.default_rng(M + K + P + 1)
    X = rng.uniform(-1, 1, (M, K))
    W = rng.uniform(-0.2, 0.2, (K, P))
    bias = rng.uniform(-0.5, 0.5, (P,))
    gZ = rng.uniform(-1, 1, (M, P))
    gX0 = rng.uniform(-1, 1, (M, K))
    orc = pyoracle.COracle("f64")
    Y_ref = orc.matmul_forward(X, W)
    gY_ref, gb_ref = orc.bias_lrelu_backward(Y_ref, bias, gZ)
    gX_ref, gW_ref = orc.matmul_backward(X, W, gY_ref, gX_init=gX0)
    gX2_ref = None
    ctx.set_mix_path(_lib.MIX_TENSOR)
    try:
        gX = dev(gX0)
        _, gW, gb = ctx.mix_backward(dev(X), dev(W), dev(gZ), bias=dev(bias), Y=dev(Y_ref), gX=gX, beta_x=1.0)
        # no activation, fresh gX
        gX2, gW2, _ = ctx.mix_backward(dev(X), dev(W), dev(gZ))
    finally:
        ctx.set_mix_path(_lib.MIX_AUTO)
    assert rel(gX.cpu().numpy(), gX_ref) < 3e-5
    assert rel(gW.cpu().numpy(), gW_ref) < TOL and rel(gb.cpu().numpy(), gb_ref) < TOL
    gX2_ref, _ = orc.matmul_backward(X, W, gZ)
    assert rel(gX2.cpu().numpy(), gX2_ref) < 3e-5
    _, gW2_ref = orc.matmul_backward(X, W, gZ)
    assert rel(gW2.cpu().numpy(), gW2_ref) < TOL
