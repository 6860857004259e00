"""Whole-model parity on the GPU: SMP_beta (GraphFlow/SMP_beta.h, BASELINE.json config 2's model) through the batched
B200 path (graphflow_b200/model.py: promotion gather -> 18-way contraction -> tensor-core mix per level, one launch set
per level for all vertices of all graphs) against the unmodified reference model: graph feature (SMP_beta::Feature),
loss and every parameter gradient after one forward/backward (what BatchLearn sums, SMP_beta.h:757-768)."""
import os

import numpy as np
import pytest
import torch

from oracle import pyoracle
from tests.conftest import GOLDEN
from tests.util import molecular_adjacency

pytestmark = pytest.mark.gpu
TOL = 1e-4


def blocks(flat, L, C, F, D):
    """Split a flat parameter(-gradient) vector into the reference's blocks H, (K_l, b_l)..., W."""
    sizes = [C * F * (D + 1)] + [s for _ in range(L) for s in (18 * C * C, C)] + [C]
    out, off = [], 0
    for s in sizes:
        out.append(flat[off:off + s])
        off += s
    return out


def check(model, graphs, targets, refs, L, C, F, D):
    tb = model.tables(graphs)
    gf, loss, grads = model.forward_backward(tb, targets)
    gf, loss, grads = gf.cpu().numpy(), loss.cpu().numpy(), grads.cpu().numpy().astype(np.float64)
    for i, r in enumerate(refs):
        assert np.abs(gf[i] - r["feature"]).max() <= TOL * np.abs(r["feature"]).max(), (i, gf[i], r["feature"])
        assert abs(loss[i] - r["loss"]) <= 1e-3 * max(1.0, abs(r["loss"]))
    want = sum(r["grads"] for r in refs)  # BatchLearn sums the per-example gradients (SumGradients.h:45-67)
    for got_b, want_b in zip(blocks(grads, L, C, F, D), blocks(want, L, C, F, D)):
        assert np.abs(got_b - want_b).max() <= TOL * np.abs(want_b).max()
    return tb


def test_smp_beta_golden_batch_of_three_graphs():
    from graphflow_b200.model import SMPBetaB200

    g = np.load(os.path.join(GOLDEN, "smp_beta_model.npz"))
    L, C, F, D = int(g["L"]), int(g["C"]), int(g["F"]), int(g["D"])
    model = SMPBetaB200(L, C, F, D)
    model.set_flat_params(g["params"])
    graphs = [(g["adj%d" % i], g["feat%d" % i]) for i in range(3)]
    refs = [{"feature": g["feature%d" % i], "loss": float(g["loss%d" % i]), "grads": g["grads%d" % i]} for i in range(3)]
    tb = check(model, graphs, [float(g["target%d" % i]) for i in range(3)], refs, L, C, F, D)
    assert len(tb.levels[-1]) > 1  # several size buckets were exercised
    assert tb.contractions == L * sum(a.shape[0] for a, _ in graphs)  # one contraction per (graph, vertex, level)


@pytest.mark.skipif(not pyoracle.model_available(), reason="oracle/_ref model shim not shipped")
def test_smp_beta_c32_fused_and_tensor_core_path():
    """C = 32: the fused contraction kernels and the tcgen05 mix, ragged receptive fields up to 14 vertices."""
    from graphflow_b200.model import SMPBetaB200

    rng = np.random.default_rng(11)
    L, C, F, D = 2, 32, 5, 2
    params = rng.uniform(-1, 1, pyoracle.smp_beta_num_params(L, C, F, D)) * 0.02
    graphs, refs, targets = [], [], []
    for V in (14, 9):
        adj = (molecular_adjacency(V, rng, self_loops=False) > 0).astype(np.int32)
        feat = np.eye(F)[rng.integers(0, F, V)]
        graphs.append((adj, feat))
        targets.append(float(V))
        refs.append(pyoracle.ref_smp_beta(adj, feat, L, C, D, params, float(V)))
    model = SMPBetaB200(L, C, F, D)
    model.set_flat_params(params)
    check(model, graphs, targets, refs, L, C, F, D)


@pytest.mark.skipif(not pyoracle.model_available(), reason="oracle/_ref model shim not shipped")
def test_smp_2d_ver8_model():
    """BASELINE config 4's model (SMP_2D_ver8: the mix is CustomMatMulTensor with K_l stored [C, 18 C]), 2 levels."""
    from graphflow_b200.model import SMPBetaB200

    rng = np.random.default_rng(12)
    L, C, F, D = 2, 32, 5, 2
    params = rng.uniform(-1, 1, pyoracle.smp_beta_num_params(L, C, F, D)) * 0.02
    graphs, refs, targets = [], [], []
    for V in (10, 13):
        adj = (molecular_adjacency(V, rng, self_loops=False) > 0).astype(np.int32)
        feat = np.eye(F)[rng.integers(0, F, V)]
        graphs.append((adj, feat))
        targets.append(float(V))
        refs.append(pyoracle.ref_smp_2d_ver8(adj, feat, L, C, D, params, float(V)))
    model = SMPBetaB200(L, C, F, D, k_transposed=True)
    model.set_flat_params(params)
    check(model, graphs, targets, refs, L, C, F, D)


@pytest.mark.skipif(not pyoracle.model_available(), reason="oracle/_ref model shim not shipped")
@pytest.mark.parametrize("L,C,max_field,sizes", [(2, 8, 5, (9, 6)), (3, 64, 8, (14, 11))])
def test_smp_omega_physics_model(L, C, max_field, sizes):
    """BASELINE config 3's model: channel widths halve per level (64 -> 32 -> 16 -> 8: fused and generic contraction kernels,
    tensor-core and SIMT mix), fields limited to max_field members, every level feeds the hidden-layer read-out."""
    from graphflow_b200.model import CCNModelB200

    rng = np.random.default_rng(13 + C)
    F = 4
    params = rng.uniform(-1, 1, pyoracle.smp_omega_num_params(L, C, F)) * (0.15 if C == 8 else 0.03)
    graphs, refs, targets = [], [], []
    for V in sizes:
        adj = (molecular_adjacency(V, rng, self_loops=False) > 0).astype(np.int32)
        feat = rng.uniform(0, 1, (V, F))
        graphs.append((adj, feat))
        targets.append(float(V))
        refs.append(pyoracle.ref_smp_omega_physics(adj, feat, max_field, L, C, params, float(V)))
    model = CCNModelB200("omega", L, C, F, max_field=max_field)
    model.set_flat_params(params)
    tb = model.tables(graphs)
    gf, loss, grads = model.forward_backward(tb, targets)
    gf, loss, grads = gf.cpu().numpy(), loss.cpu().numpy(), grads.cpu().numpy().astype(np.float64)
    for i, r in enumerate(refs):
        assert np.abs(gf[i] - r["feature"]).max() <= TOL * np.abs(r["feature"]).max()
        assert abs(loss[i] - r["loss"]) <= 1e-3 * max(1.0, abs(r["loss"]))
    want = sum(r["grads"] for r in refs)
    off = 0
    for shp in model.shapes:  # per parameter block, normalised by the block's largest gradient
        k = int(np.prod(shp))
        assert np.abs(grads[off:off + k] - want[off:off + k]).max() <= TOL * max(np.abs(want[off:off + k]).max(), 1e-12), shp
        off += k


def test_cuda_graph_step_matches_eager_and_tracks_parameter_updates():
    """`capture_step`: the whole forward+backward of a ragged batch as one CUDA graph gives the eager results, and a replay
    after an in-place parameter update gives the eager results at the new parameters."""
    from graphflow_b200.model import CCNModelB200
    from tests.util import molecular_adjacency

    L, C, F, D = 2, 32, 5, 2
    rng = np.random.default_rng(21)
    graphs = []
    for V in (9, 12, 7, 12, 10):
        adj = (molecular_adjacency(V, rng, self_loops=False) > 0).astype(np.int32)
        graphs.append((adj, np.eye(F)[rng.integers(0, F, V)]))
    targets = [float(a.shape[0]) for a, _ in graphs]
    model = CCNModelB200("beta", L, C, F, n_depth=D)
    model.set_flat_params(rng.uniform(-1, 1, model.num_params()) * 0.05)
    tb = model.tables(graphs)
    step = model.capture_step(tb, targets)
    for trial in range(2):
        gf_e, loss_e, g_e = (t.clone() for t in model.forward_backward(tb, targets))
        gf_g, loss_g, g_g = step()
        torch.cuda.synchronize()
        assert (gf_g - gf_e).abs().max().item() <= 1e-5 * gf_e.abs().max().item()
        assert (loss_g - loss_e).abs().max().item() <= 1e-5 * loss_e.abs().max().item()
        assert (g_g - g_e).abs().max().item() <= 1e-4 * g_e.abs().max().item()
        model.set_flat_params_device(model.get_flat_params() * 1.01 + 0.001)   # in place: the graph reads the same tensors


def test_graph_feature_is_permutation_invariant():
    """The reference's tests/test_graph_permutation_invariant.cpp:105-172 on this path: an Erdos-Renyi graph with random 0/1
    vertex features and a randomly relabelled copy give the same graph feature (there: `Difference in norm l1`)."""
    from graphflow_b200.model import CCNModelB200

    V, F, L, C, D = 14, 6, 3, 16, 2
    rng = np.random.default_rng(0)
    up = np.triu(rng.integers(0, 2, (V, V)), 1)
    adj = (up + up.T).astype(np.int32)
    feat = rng.integers(0, 2, (V, F)).astype(np.float64)
    perm = rng.permutation(V)
    adj2, feat2 = adj[np.ix_(perm, perm)], feat[perm]
    model = CCNModelB200("beta", L, C, F, n_depth=D)
    model.set_flat_params(rng.uniform(-1, 1, model.num_params()) * 0.1)
    f1, f2 = model.Feature((adj, feat)), model.Feature((adj2, feat2))
    assert np.abs(f1).max() > 0
    assert np.abs(f1 - f2).sum() <= 1e-4 * np.abs(f1).sum()
    if pyoracle.model_available():                              # and it is the reference's feature
        ref = pyoracle.ref_smp_beta(adj, feat, L, C, D, model.get_flat_params().cpu().numpy().astype(np.float64), 0.0)
        assert np.abs(f1 - ref["feature"]).max() <= 1e-4 * np.abs(ref["feature"]).max()


@pytest.mark.skipif(not pyoracle.model_available(), reason="oracle/_ref model shim not shipped")
@pytest.mark.parametrize("L,C,D,max_field,sizes", [(2, 8, 2, 5, (9, 12)), (3, 32, 1, 6, (14, 10))])
def test_plain_smp_omega_model(L, C, D, max_field, sizes):
    """SMP_omega (SMP_omega.h): SMP_beta's wiring on receptive fields cut to max_field members (distance, then WL rank) and
    ordered by rank: graph feature, loss and every parameter gradient against the unmodified model."""
    from graphflow_b200.model import CCNModelB200

    rng = np.random.default_rng(5 + C)
    F = 4
    params = rng.uniform(-1, 1, pyoracle.smp_beta_num_params(L, C, F, D)) * (0.15 if C == 8 else 0.04)
    graphs, refs, targets = [], [], []
    for V in sizes:
        adj = (molecular_adjacency(V, rng, self_loops=False) > 0).astype(np.int32)
        feat = np.eye(F)[rng.integers(0, F, V)]
        graphs.append((adj, feat))
        targets.append(float(V) / 4)
        refs.append(pyoracle.ref_smp_omega(adj, feat, max_field, L, C, D, params, float(V) / 4))
    model = CCNModelB200("omega_wl", L, C, F, n_depth=D, max_field=max_field)
    model.set_flat_params(params)
    tb = model.tables(graphs)
    assert max(b["n_max"] for lv in tb.levels for b in lv) <= max_field
    gf, loss, grads = model.forward_backward(tb, targets)
    gf, loss, grads = gf.cpu().numpy(), loss.cpu().numpy(), grads.cpu().numpy().astype(np.float64)
    for i, r in enumerate(refs):
        assert np.abs(gf[i] - r["feature"]).max() <= TOL * np.abs(r["feature"]).max()
        assert abs(loss[i] - r["loss"]) <= 1e-3 * max(1.0, abs(r["loss"]))
    want = sum(r["grads"] for r in refs)
    off = 0
    for shp in model.shapes:
        k = int(np.prod(shp))
        assert np.abs(grads[off:off + k] - want[off:off + k]).max() <= TOL * max(np.abs(want[off:off + k]).max(), 1e-12), shp
        off += k


def _check_pairgraphs(model, pairs, targets, refs):
    tbs = model.tables(pairs)
    gf, loss, grads = model.forward_backward(tbs, targets)
    gf, loss, grads = gf.cpu().numpy(), loss.cpu().numpy(), grads.cpu().numpy().astype(np.float64)
    for i, r in enumerate(refs):
        assert np.abs(gf[i] - r["feature"]).max() <= TOL * np.abs(r["feature"]).max()
        assert abs(loss[i] - r["loss"]) <= 1e-3 * max(1.0, abs(r["loss"]))
    want = sum(r["grads"] for r in refs)
    off = 0
    for p in model.params:  # per parameter block, normalised by the block's largest gradient
        k = p.numel()
        assert np.abs(grads[off:off + k] - want[off:off + k]).max() <= TOL * max(np.abs(want[off:off + k]).max(), 1e-12), tuple(p.shape)
        off += k
    assert off == want.size


def test_smp_omega_pairgraphs_golden():
    """SMP_omega_pairgraphs (SMP_omega_pairgraphs.h): the path run on the graph and on its line graph with separate parameters,
    level features concatenated level by level, two hidden layers -- graph feature, loss and every parameter gradient of a batch
    of two examples against the committed outputs of the unmodified reference model."""
    from graphflow_b200.model import PairGraphsModelB200

    g = np.load(os.path.join(GOLDEN, "smp_omega_pairgraphs.npz"))
    L, C, F, mf = int(g["L"]), int(g["C"]), int(g["F"]), int(g["max_field"])
    model = PairGraphsModelB200(L, C, F, F, mf)
    assert model.num_params() == g["params"].size
    model.set_flat_params(g["params"])
    pairs = [((g["adj%d" % i], g["feat%d" % i]), (g["ladj%d" % i], g["lfeat%d" % i])) for i in range(2)]
    refs = [{"feature": g["feature%d" % i], "loss": float(g["loss%d" % i]), "grads": g["grads%d" % i]} for i in range(2)]
    _check_pairgraphs(model, pairs, [float(g["target%d" % i]) for i in range(2)], refs)
    assert abs(model.Predict(pairs[0]) - float(g["predict0"])) <= 1e-4 * max(1.0, abs(float(g["predict0"])))


@pytest.mark.skipif(not pyoracle.model_available(), reason="oracle/_ref model shim not shipped")
def test_smp_omega_pairgraphs_c64_fused_path():
    """C = 64 -> 32 -> 16: the fused contraction kernels and the tensor-core mix on both trunks, fields limited to 8 members."""
    from graphflow_b200.model import PairGraphsModelB200
    from tests.util import line_graph

    rng = np.random.default_rng(77)
    L, C, F, mf = 2, 64, 4, 8
    params = rng.uniform(-1, 1, pyoracle.smp_omega_pairgraphs_num_params(L, C, F, F)) * 0.03
    pairs, refs, targets = [], [], []
    for V in (12, 10):
        adj = (molecular_adjacency(V, rng, self_loops=False) > 0).astype(np.int32)
        feat = rng.uniform(0, 1, (V, F))
        a2, f2 = line_graph(adj, feat)
        pairs.append(((adj, feat), (a2, f2)))
        targets.append(float(V) / 4)
        refs.append(pyoracle.ref_smp_omega_pairgraphs(adj, feat, a2, f2, mf, L, C, params, float(V) / 4))
    model = PairGraphsModelB200(L, C, F, F, mf)
    model.set_flat_params(params)
    _check_pairgraphs(model, pairs, targets, refs)
    before, after = model.BatchLearn(pairs, targets, 1e-5)  # one Adam step: the parameters moved, the loss is finite
    assert abs(before - sum(r["loss"] for r in refs)) <= 1e-3 * max(1.0, before)
    assert np.isfinite(after) and after != before
