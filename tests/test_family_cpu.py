"""RisiContraction_4 / RisiContraction_10 / RisiContraction_18_dropout on the CPU: the einsum statements of
oracle/pyoracle.py against the golden vectors generated from the compiled reference (tests/golden/make_golden.py
`family`) and, when oracle/_ref is built, against the reference itself on fresh inputs."""
import os

import numpy as np
import pytest

from oracle import pyoracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "family_n5_c2.npz")


def test_einsum_statements_match_golden():
    g = np.load(GOLDEN)
    T, adj = g["T"], g["adj"]
    assert np.abs(pyoracle.einsum4_forward(T) - g["out4"]).max() < 1e-12
    assert np.abs(pyoracle.einsum4_backward(g["g4"]) - g["gT4"]).max() < 1e-12
    assert np.abs(pyoracle.einsum10_forward(T, adj) - g["out10"]).max() < 1e-12
    assert np.abs(pyoracle.einsum10_backward(g["g10"], adj) - g["gT10"]).max() < 1e-12
    for i in range(3):
        use = [bool(u) for u in g["drop%d_use" % i]]
        assert np.abs(pyoracle.einsum18_dropout_forward(T, adj, use) - g["drop%d_out" % i]).max() < 1e-12
        assert np.abs(pyoracle.einsum18_dropout_backward(g["g18"], adj, use) - g["drop%d_gT" % i]).max() < 1e-12
    # test mode: every slab, scaled by nKept / 18 (RisiContraction_18_dropout.h:467-472)
    scale = float(g["test_kept"]) / 18.0
    assert np.abs(scale * pyoracle.einsum18_forward(T, adj, True) - g["test_out"]).max() < 1e-12


def test_the_10_are_the_first_ten_of_the_50_and_the_4_need_no_adjacency():
    assert pyoracle.EINSUM10 == pyoracle.EINSUM50[:10]
    rng = np.random.default_rng(3)
    T = rng.uniform(-1, 1, (4, 4, 4, 2))
    out4 = pyoracle.einsum4_forward(T).reshape(4, 4, 4, 2)
    # slabs 0 and 1 are the adjacency-free factors of cases 1 and 5 of the 50 (abcf,de->abf / bcf with sum(A) = 1)
    A = np.full((4, 4), 1.0 / 16.0)
    out50 = pyoracle.einsum50_forward(T, A).reshape(4, 4, 50, 2)
    assert np.abs(out4[:, :, 0] - out50[:, :, 0]).max() < 1e-12
    assert np.abs(out4[:, :, 1] - out50[:, :, 4]).max() < 1e-12


@pytest.mark.skipif(not pyoracle.ref_available("f64"), reason="oracle/_ref not built")
def test_einsum_statements_match_compiled_reference():
    rng = np.random.default_rng(17)
    N, C = 4, 3
    T, adj = rng.uniform(-1, 1, (N, N, N, C)), rng.uniform(-1, 1, (N, N))
    g4, g10, g18 = (rng.uniform(-1, 1, (N, N, k * C)) for k in (4, 10, 18))
    r = pyoracle.RefOracle("f64")
    out, gT = r.contract4(T, g4)
    assert np.abs(out - pyoracle.einsum4_forward(T)).max() < 1e-12 and np.abs(gT - pyoracle.einsum4_backward(g4)).max() < 1e-12
    out, gT = r.contract10(T, adj, g10)
    assert np.abs(out - pyoracle.einsum10_forward(T, adj)).max() < 1e-12
    assert np.abs(gT - pyoracle.einsum10_backward(g10, adj)).max() < 1e-12
    masks = set()
    for seed in range(6):
        out, gT, use = r.contract18_dropout(T, adj, g18, 9, seed)
        masks.add(tuple(use))
        assert sum(use) == 9
        assert np.abs(out - pyoracle.einsum18_dropout_forward(T, adj, use)).max() < 1e-12
        assert np.abs(gT - pyoracle.einsum18_dropout_backward(g18, adj, use)).max() < 1e-12
    assert len(masks) > 1  # rand() really drives the selection


def test_r4_reference_test_recipe_kat_is_exact():
    """tests/test_RisiContraction_4_thread.cpp:49-66 (the tensors of the R50 KAT): integer-valued, so exact."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "kat_r50_n10_c5.npz"))
    assert np.array_equal(pyoracle.einsum4_forward(g["T"]), g["out4"])
    assert np.array_equal(pyoracle.einsum4_backward(g["g4"]), g["gT4"])
