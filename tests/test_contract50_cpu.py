"""RisiContraction_50 on the CPU: the plain-C restatement and the einsum statement against the compiled reference
(GraphFlow/RisiContraction_50.h, N^6 loops), and the factorised evaluation plan the CUDA kernels use
(graphflow_b200/csrc/gen/gen_r50_table.py -> r50_table.inc) against all of them."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import pyoracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "r50_n5_c2.npz")


def _generator():
    path = os.path.join(ROOT, "graphflow_b200", "csrc", "gen", "gen_r50_table.py")
    spec = importlib.util.spec_from_file_location("gen_r50_table", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _inputs(N, C, seed):
    rng = np.random.default_rng(seed)
    return (rng.uniform(-1, 1, (N, N, N, C)), rng.uniform(-1, 1, (N, N)), rng.uniform(-1, 1, (N, N, 50 * C)),
            rng.uniform(-1, 1, (N, N, N, C)))


def test_restatement_and_einsum_match_golden():
    g = np.load(GOLDEN)
    c = pyoracle.COracle("f64")
    assert np.abs(c.contract50_forward(g["T"], g["adj"]) - g["out"]).max() < 1e-12
    assert np.abs(pyoracle.einsum50_forward(g["T"], g["adj"]) - g["out"]).max() < 1e-12
    assert np.abs(c.contract50_backward(g["gout"], g["adj"], g["gT0"]) - g["gT"]).max() < 1e-12
    assert np.abs(pyoracle.einsum50_backward(g["gout"], g["adj"]) + g["gT0"] - g["gT"]).max() < 1e-12


@pytest.mark.skipif(not pyoracle.ref_available("f64"), reason="oracle/_ref not built")
def test_restatement_matches_compiled_reference():
    T, A, gout, g0 = _inputs(4, 3, 11)
    r, c = pyoracle.RefOracle("f64"), pyoracle.COracle("f64")
    assert np.abs(c.contract50_forward(T, A) - r.contract50_forward(T, A)).max() < 1e-12
    assert np.abs(c.contract50_backward(gout, A, g0) - r.contract50_backward(gout, A, g0)).max() < 1e-12
    r32, c32 = pyoracle.RefOracle("f32"), pyoracle.COracle("f32")
    assert pyoracle.slab_rel_err(c32.contract50_forward(T, A), r32.contract50_forward(T, A), 50) < 1e-5


def test_the_18_are_a_subset_of_the_50():
    """RisiContraction_18.h:102-318 tags every case with its number in the 50 ("(k/50)"); SURVEY.md section 8a row a9."""
    subset = [1, 3, 5, 6, 10, 11, 13, 17, 18, 23, 26, 27, 28, 38, 40, 43, 46, 50]
    assert [pyoracle.EINSUM50[k - 1] for k in subset] == pyoracle.EINSUM18


def test_factorised_plan_matches_einsum_and_table_is_current():
    gen = _generator()
    assert gen.EINSUM50 == pyoracle.EINSUM50
    T, A, _, _ = _inputs(6, 3, 5)
    assert np.abs(gen.emulate(gen.EINSUM50, T, A) - pyoracle.einsum50_forward(T, A)).max() < 1e-12
    rows = [ln for ln in open(os.path.join(ROOT, "graphflow_b200", "csrc", "r50_table.inc")) if ln.strip().startswith("{")]
    assert len(rows) == 50
    for k, ln in enumerate(rows):
        want = "{%d, %2d, %d, %d}" % gen.plan(gen.EINSUM50[k])
        assert ln.strip().startswith(want), (k, ln, want)


def test_reference_test_recipe_kat_is_exact():
    """tests/test_RisiContraction_50.cpp:49-80 (N=10, 5 channels, rand()%100 tensors, 0/1 adjacency): integer-valued and
    below 2^24, so the restatement reproduces the compiled reference exactly, in double and in float; the test's own
    observation (`Check the redundancy`, :101-118) -- symmetric inputs make many of the 50 slabs coincide -- holds too."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "kat_r50_n10_c5.npz"))
    for prec in ("f64", "f32"):
        c = pyoracle.COracle(prec)
        assert np.array_equal(c.contract50_forward(g["T"], g["adj"]).astype(np.float64), g["out"])
        assert np.array_equal(c.contract50_backward(g["gout"], g["adj"]).astype(np.float64), g["gT"])
    assert np.array_equal(pyoracle.einsum50_forward(g["T"], g["adj"]), g["out"])
    N, C = g["adj"].shape[0], g["T"].shape[3]
    slabs = g["out"].reshape(N, N, 50, C)
    distinct = {slabs[:, :, k].tobytes() for k in range(50)}
    assert len(distinct) < 50
