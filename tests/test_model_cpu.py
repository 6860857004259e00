"""Host graph preprocessing (graphflow_b200/graph.py) against the reference model SMP_beta: the receptive fields must be
the model's own, and driving the CPU oracle operators (promotion gather, 18-way contraction, MatMul, bias + leaky-ReLU)
with the derived index tables must reproduce SMP_beta::Feature -- i.e. the tables (pos / m / reduced adjacency) are right
before any GPU is involved."""
import os

import numpy as np
import pytest

from graphflow_b200.graph import GraphTables
from oracle import pyoracle
from tests.conftest import GOLDEN
from tests.util import molecular_adjacency

ALPHA = 0.01


def lrelu(x):
    return np.where(x > 0, x, ALPHA * x)


def oracle_feature(gt, params, L, C, F, D):
    """SMP_beta forward (SMP_beta.h:554-639) with the oracle's operators and graph.py's tables, fp64."""
    c = pyoracle.COracle("f64")
    off = 0

    def take(shape):
        nonlocal off
        k = int(np.prod(shape))
        out = params[off:off + k].reshape(shape)
        off += k
        return out

    H = take((C, F * (D + 1)))
    f_prev = [lrelu(H @ gt.features[v]).reshape(1, 1, C) for v in range(gt.V)]
    for l in range(L):
        K, b = take((18 * C, C)), take((C,))
        cur = []
        for v in range(gt.V):
            it = gt.levels[l][v]
            n = it["n"]
            T = np.stack([c.promote_forward(f_prev[it["src"][a]], it["pos"][a]) for a in range(n)])
            X = c.contract18_forward(T, it["adj"].astype(np.float64)).reshape(n * n, 18 * C)
            cur.append(c.bias_lrelu_forward(c.matmul_forward(X, K), b).reshape(n, n, C))
        f_prev = cur
    return sum(lrelu(f.sum((0, 1))) for f in f_prev)


def test_tables_reproduce_golden_model_features():
    g = np.load(os.path.join(GOLDEN, "smp_beta_model.npz"))
    L, C, F, D = int(g["L"]), int(g["C"]), int(g["F"]), int(g["D"])
    for gi in range(3):
        gt = GraphTables(g["adj%d" % gi], g["feat%d" % gi], L, D)
        assert [len(f) for f in gt.phi[L]] == list(g["phi%d" % gi])
        feat = oracle_feature(gt, g["params"], L, C, F, D)
        assert np.abs(feat - g["feature%d" % gi]).max() < 1e-10 * max(1.0, np.abs(g["feature%d" % gi]).max())


@pytest.mark.skipif(not pyoracle.model_available(), reason="oracle/_ref model shim not built")
def test_receptive_fields_match_reference_model():
    rng = np.random.default_rng(2)
    L, C, F, D = 3, 2, 4, 2
    params = rng.uniform(-0.05, 0.05, pyoracle.smp_beta_num_params(L, C, F, D))
    for V in (5, 12, 17):
        adj = (molecular_adjacency(V, rng, self_loops=False) > 0).astype(np.int32)
        feat = np.eye(F)[rng.integers(0, F, V)]
        ref = pyoracle.ref_smp_beta(adj, feat, L, C, D, params, 1.0)
        gt = GraphTables(adj, feat, L, D)
        assert gt.phi == ref["phi"]
        feat_o = oracle_feature(gt, params, L, C, F, D)
        assert np.abs(feat_o - ref["feature"]).max() < 1e-9 * max(1.0, np.abs(ref["feature"]).max())


@pytest.mark.skipif(not pyoracle.model_available(), reason="oracle/_ref model shim not built")
def test_omega_receptive_fields_match_reference_model():
    """SMP_omega_physics: insertion-ordered fields cut to max_field by dropping whole outer distance shells."""
    rng = np.random.default_rng(3)
    for V, L, C, F, mf in ((9, 2, 8, 3, 5), (14, 3, 4, 4, 7), (12, 3, 8, 2, 12)):
        adj = (molecular_adjacency(V, rng, self_loops=False) > 0).astype(np.int32)
        feat = rng.uniform(0, 1, (V, F))
        params = rng.uniform(-0.1, 0.1, pyoracle.smp_omega_num_params(L, C, F))
        ref = pyoracle.ref_smp_omega_physics(adj, feat, mf, L, C, params, 1.0)
        gt = GraphTables(adj, feat, L, kind="omega", max_field=mf)
        assert gt.phi == ref["phi"]
        assert max(len(f) for f in gt.phi[L]) <= mf


@pytest.mark.skipif(not pyoracle.model_available(), reason="oracle/_ref model shim not built")
def test_plain_smp_omega_fields_and_feature_match_reference_model():
    """SMP_omega (SMP_omega.h:476-531): SMP_beta's WL features and ranking, receptive fields cut to max_field by distance
    then rank and ordered by rank.  The native tables (CCN_GRAPH_OMEGA_WL), the numpy restatement and the unmodified model
    agree on every field, and the oracle operators driven by the tables reproduce the model's graph feature."""
    rng = np.random.default_rng(31)
    for V, L, C, F, D, mf in ((9, 2, 2, 4, 2, 5), (14, 3, 2, 4, 1, 6), (12, 2, 3, 3, 2, 12), (16, 3, 2, 4, 2, 4)):
        adj = (molecular_adjacency(V, rng, self_loops=False) > 0).astype(np.int32)
        feat = np.eye(F)[rng.integers(0, F, V)]
        params = rng.uniform(-0.05, 0.05, pyoracle.smp_beta_num_params(L, C, F, D))
        ref = pyoracle.ref_smp_omega(adj, feat, mf, L, C, D, params, 1.0)
        native = GraphTables(adj, feat, L, D, kind="omega_wl", max_field=mf)
        restated = GraphTables(adj, feat, L, D, kind="omega_wl", max_field=mf, native=False)
        assert native.phi == ref["phi"] == restated.phi
        assert max(len(f) for f in native.phi[L]) <= mf
        for l in range(L):
            for v in range(V):
                a, b = native.levels[l][v], restated.levels[l][v]
                assert a["n"] == b["n"] and a["src"] == b["src"] and a["m"] == b["m"]
                assert np.array_equal(a["pos"], b["pos"]) and np.array_equal(a["adj"], b["adj"])
        feat_o = oracle_feature(native, params, L, C, F, D)
        assert np.abs(feat_o - ref["feature"]).max() < 1e-9 * max(1.0, np.abs(ref["feature"]).max())


@pytest.mark.skipif(not pyoracle.model_available(), reason="oracle/_ref model shim not built")
def test_pairgraphs_fixture_is_what_the_reference_model_gives():
    """tests/golden/smp_omega_pairgraphs.npz (the GPU test's target) regenerated from the unmodified SMP_omega_pairgraphs."""
    import os

    from tests.conftest import GOLDEN
    from tests.golden.make_golden import pairgraphs_inputs

    g = np.load(os.path.join(GOLDEN, "smp_omega_pairgraphs.npz"))
    L, C, F, mf, params, ex = pairgraphs_inputs()
    assert (L, C, F, mf) == (int(g["L"]), int(g["C"]), int(g["F"]), int(g["max_field"])) and np.array_equal(params, g["params"])
    for i, (adj, feat, a2, f2, target) in enumerate(ex):
        assert np.array_equal(adj, g["adj%d" % i]) and np.array_equal(a2, g["ladj%d" % i])
        out = pyoracle.ref_smp_omega_pairgraphs(adj, feat, a2, f2, mf, L, C, params, target)
        assert np.allclose(out["feature"], g["feature%d" % i], rtol=1e-12, atol=0)
        assert np.allclose(out["grads"], g["grads%d" % i], rtol=1e-10, atol=1e-14)
        assert abs(out["loss"] - float(g["loss%d" % i])) <= 1e-12 * abs(out["loss"])
        # the loss is the squared error of the prediction (SquaredLoss.h:50-58)
        assert abs(0.5 * (out["predict"] - target) ** 2 - out["loss"]) <= 1e-12 * max(1.0, out["loss"])
