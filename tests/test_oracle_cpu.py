"""CPU tests that pin the oracle (oracle/ccn_oracle.c, oracle/pyoracle.py) to the reference.

1. against the committed golden vectors generated from the unmodified reference (tests/golden/make_golden.py);
2. against oracle/_ref itself when that library is present (it is wherever /root/reference was at build time).
"""
import os

import numpy as np
import pytest

from oracle import pyoracle
from tests.conftest import GOLDEN


def load(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="module")
def c64():
    return pyoracle.COracle("f64")


@pytest.fixture(scope="module")
def c32():
    return pyoracle.COracle("f32")


def test_pattern_table_cites_reference_lines(c64):
    # One forward update per case in GraphFlow/RisiContraction_18.h; strictly increasing line numbers.
    lines = c64.pattern_ref_lines()
    assert len(lines) == 18 and lines == sorted(lines) and lines[0] == 102 and lines[-1] == 318


def test_kat_c1_exact(c64, c32):
    g = load("kat_c1_n8_c4.npz")
    for orc in (c64, c32):
        out = orc.contract18_forward(g["T"], g["adj"])
        assert np.array_equal(out.astype(np.float64), g["out"].astype(np.float64))
        gT = orc.contract18_backward(g["gout"], g["adj"])
        assert np.array_equal(gT.astype(np.float64), g["gT"].astype(np.float64))
    assert np.array_equal(pyoracle.einsum18_forward(g["T"], g["adj"]), g["out"].astype(np.float64))
    assert np.array_equal(pyoracle.einsum18_backward(g["gout"], g["adj"]), g["gT"].astype(np.float64))


def test_real_and_accumulate(c64):
    g = load("real_n6_c8.npz")
    out = c64.contract18_forward(g["T"], g["adj"])
    assert pyoracle.slab_rel_err(out, g["out"]) < 1e-13
    gT = c64.contract18_backward(g["gout"], g["adj"], g["gT0"])
    assert np.abs(gT - g["gT"]).max() < 1e-12
    assert pyoracle.slab_rel_err(pyoracle.einsum18_forward(g["T"], g["adj"]), g["out"]) < 1e-13
    assert np.abs(pyoracle.einsum18_backward(g["gout"], g["adj"]) + g["gT0"] - g["gT"]).max() < 1e-12


def test_signed_adjacency_positive_part_and_raw(c64):
    g = load("signed_n5_c3.npz")
    assert pyoracle.slab_rel_err(c64.contract18_forward(g["T"], g["adj"], True), g["out"]) < 1e-13
    assert pyoracle.slab_rel_err(c64.contract18_forward(g["T"], g["adj"], False), g["out_raw"]) < 1e-13
    assert np.abs(c64.contract18_backward(g["gout"], g["adj"]) - g["gT"]).max() < 1e-12
    assert pyoracle.slab_rel_err(pyoracle.einsum18_forward(g["T"], g["adj"], True), g["out"]) < 1e-13
    assert pyoracle.slab_rel_err(pyoracle.einsum18_forward(g["T"], g["adj"], False), g["out_raw"]) < 1e-13
    assert np.abs(pyoracle.einsum18_backward(g["gout"], g["adj"]) - g["gT"]).max() < 1e-12


def test_fp32_restatement_close_to_fp64_reference(c32):
    g = load("real_n6_c8.npz")
    assert pyoracle.slab_rel_err(c32.contract18_forward(g["T"], g["adj"]), g["out"]) < 1e-5


def test_matmul_golden(c64):
    g = load("matmul_20x36x5.npz")
    assert np.array_equal(c64.matmul_forward(g["X"], g["W"]), g["Y"])
    gX, gW = c64.matmul_backward(g["X"], g["W"], g["gY"], g["gX0"], g["gW0"])
    assert np.array_equal(gX, g["gX"]) and np.array_equal(gW, g["gW"])


def test_level_chain_golden(c64):
    g = load("level_n6_c4.npz")
    N, C, Cout = 6, 4, 4
    contracted = c64.contract18_forward(g["T"], g["adj"])
    assert pyoracle.slab_rel_err(contracted, g["contracted"]) < 1e-13
    Y = c64.matmul_forward(contracted.reshape(N * N, 18 * C), g["K"])
    Z = c64.bias_lrelu_forward(Y, g["bias"])
    assert np.abs(Z.reshape(N, N, Cout) - g["Z"]).max() < 1e-12
    gY, gb = c64.bias_lrelu_backward(Y, g["bias"], g["gZ"].reshape(N * N, Cout))
    gX, gK = c64.matmul_backward(contracted.reshape(N * N, 18 * C), g["K"], gY)
    gT = c64.contract18_backward(gX.reshape(N, N, 18 * C), g["adj"])
    assert np.abs(gb - g["gb"]).max() < 1e-12
    assert np.abs(gK - g["gK"]).max() < 1e-11
    assert np.abs(gT - g["gT"]).max() < 1e-10


def test_adjoint_identity(c64):
    # <contract(T), G> == <T, contract^T(G)>: the size-independent property used at full size on the GPU.
    rng = np.random.default_rng(7)
    N, C = 7, 3
    T = rng.uniform(-1, 1, (N, N, N, C))
    adj = rng.uniform(-0.5, 1, (N, N))
    G = rng.uniform(-1, 1, (N, N, 18 * C))
    lhs = float((c64.contract18_forward(T, adj) * G).sum())
    rhs = float((T * c64.contract18_backward(G, adj)).sum())
    assert abs(lhs - rhs) < 1e-9 * max(1.0, abs(lhs))


@pytest.mark.skipif(not pyoracle.ref_available("f64"), reason="oracle/_ref not built here")
@pytest.mark.parametrize("N,C", [(4, 2), (9, 5), (10, 5)])
def test_restatement_vs_compiled_reference(c64, N, C):
    ref = pyoracle.RefOracle("f64")
    rng = np.random.default_rng(N * 100 + C)
    T = rng.uniform(-1, 1, (N, N, N, C))
    adj = rng.uniform(-1, 1, (N, N))
    G = rng.uniform(-1, 1, (N, N, 18 * C))
    assert pyoracle.slab_rel_err(c64.contract18_forward(T, adj), ref.contract18_forward(T, adj)) < 1e-13
    assert pyoracle.slab_rel_err(c64.contract18_forward(T, adj, False), ref.contract18_forward(T, adj, "definition")) < 1e-13
    assert pyoracle.slab_rel_err(c64.contract18_forward(T, adj, False), ref.contract18_forward(T, adj, "thread")) < 1e-13
    assert np.abs(c64.contract18_backward(G, adj) - ref.contract18_backward(G, adj)).max() < 1e-11
    assert pyoracle.slab_rel_err(pyoracle.einsum18_forward(T, adj), ref.contract18_forward(T, adj)) < 1e-13


@pytest.mark.skipif(not pyoracle.ref_available("f32"), reason="oracle/_ref not built here")
def test_fp32_reference_tree_matches(c32):
    ref = pyoracle.RefOracle("f32")
    rng = np.random.default_rng(5)
    N, C = 8, 4
    T = rng.uniform(-1, 1, (N, N, N, C)).astype(np.float32)
    up = np.triu((rng.uniform(size=(N, N)) < 0.3), 1).astype(np.float32)
    adj = up + up.T + np.eye(N, dtype=np.float32)
    assert pyoracle.slab_rel_err(c32.contract18_forward(T, adj), ref.contract18_forward(T, adj)) < 2e-6
