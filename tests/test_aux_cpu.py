"""CPU: the plain-C restatements of TensorMul, CustomMatMulTensor and the promotion X f X^T against the committed
fixtures (generated from the compiled reference by tests/golden/make_golden.py) and against the reference itself."""
import os

import numpy as np
import pytest

from oracle import pyoracle
from tests.conftest import GOLDEN


def test_restatements_match_golden():
    g = np.load(os.path.join(GOLDEN, "aux_ops.npz"))
    c = pyoracle.COracle("f64")
    assert np.abs(c.tensor_mul_forward(g["tm_A"], g["tm_B"]) - g["tm_out"]).max() < 1e-13
    gA, gB = c.tensor_mul_backward(g["tm_A"], g["tm_B"], g["tm_g"], g["tm_gA0"], g["tm_gB0"])
    assert np.abs(gA - g["tm_gA"]).max() < 1e-13 and np.abs(gB - g["tm_gB"]).max() < 1e-13
    assert np.abs(c.custom_matmul_tensor_forward(g["cm_Kt"], g["cm_X"]) - g["cm_Y"]).max() < 1e-13
    gKt, gX = c.custom_matmul_tensor_backward(g["cm_Kt"], g["cm_X"], g["cm_gY"], g["cm_gKt0"], g["cm_gX0"])
    assert np.abs(gKt - g["cm_gKt"]).max() < 1e-13 and np.abs(gX - g["cm_gX"]).max() < 1e-13
    assert np.array_equal(c.promote_forward(g["pr_f"], g["pr_pos"]), g["pr_Q"])
    assert np.abs(c.promote_backward(g["pr_gQ"], g["pr_pos"], 4, g["pr_gf0"]) - g["pr_gf"]).max() < 1e-13


@pytest.mark.skipif(not pyoracle.ref_available("f64"), reason="oracle/_ref not built")
def test_restatements_match_compiled_reference():
    rng = np.random.default_rng(9)
    r, c = pyoracle.RefOracle("f64"), pyoracle.COracle("f64")
    A, B, g = rng.uniform(-1, 1, (6, 6, 4)), rng.uniform(-1, 1, (6, 6, 4)), rng.uniform(-1, 1, (6, 6, 4))
    out, gA, gB = r.tensor_mul(A, B, g)
    assert np.abs(out - c.tensor_mul_forward(A, B)).max() < 1e-13
    cA, cB = c.tensor_mul_backward(A, B, g)
    assert np.abs(gA - cA).max() < 1e-13 and np.abs(gB - cB).max() < 1e-13
    # CustomMatMulTensor is MatMul with the weights transposed
    Kt, X = rng.uniform(-1, 1, (3, 10)), rng.uniform(-1, 1, (4, 4, 10))
    Y = r.custom_matmul_tensor(Kt, X)
    assert np.abs(Y.reshape(16, 3) - c.matmul_forward(X.reshape(16, 10), Kt.T.copy())).max() < 1e-13
    # promotion: a partial injective selection, as init_permutation_matrix builds it
    m, n, C = 7, 9, 3
    f = rng.uniform(-1, 1, (m, m, C))
    pos = np.full(n, -1, np.int32)
    rows = rng.choice(n, 5, replace=False)
    pos[rows] = rng.choice(m, 5, replace=False)
    gQ = rng.uniform(-1, 1, (n, n, C))
    Q, gf = r.promote(f, pos, gQ)
    assert np.array_equal(Q, c.promote_forward(f, pos))
    assert np.abs(gf - c.promote_backward(gQ, pos, m)).max() < 1e-13
