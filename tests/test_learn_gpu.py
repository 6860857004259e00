"""The reference's "does it learn" integration test (tests/test_SMP_beta.cpp:70-146, 174-201): overfit the
four hard-coded molecules CH4, NH3, H2O, C2H4 (target = number of atoms) with SMP_beta -- here through the batched B200
path with the gradients the kernels produce and a plain Adam update.  Same graphs, features (one-hot atom type C/H/N/O),
model sizes (10 channels, nDepth 5; the reference uses 1 level, 2 are used here so that a non-trivial contraction runs)."""
import numpy as np
import pytest
import torch

from oracle import pyoracle

pytestmark = pytest.mark.gpu

ATOMS = {"C": 0, "H": 1, "N": 2, "O": 3}
MOLECULES = {
    "CH4": (["C", "H", "H", "H", "H"], [(0, 1), (0, 2), (0, 3), (0, 4)]),
    "NH3": (["N", "H", "H", "H"], [(0, 1), (0, 2), (0, 3)]),
    "H2O": (["O", "H", "H"], [(0, 1), (0, 2)]),
    "C2H4": (["C", "H", "H", "C", "H", "H"], [(0, 1), (0, 2), (0, 3), (3, 4), (3, 5)]),
}


def molecule(name):
    labels, edges = MOLECULES[name]
    V = len(labels)
    adj = np.zeros((V, V), np.int32)
    for u, v in edges:
        adj[u, v] = adj[v, u] = 1
    feat = np.zeros((V, 4))
    for i, a in enumerate(labels):
        feat[i, ATOMS[a]] = 1.0
    return adj, feat, float(V)


def test_overfits_the_four_reference_molecules():
    from graphflow_b200.model import SMPBetaB200

    L, C, F, D = 2, 10, 4, 5
    rng = np.random.default_rng(0)
    data = [molecule(n) for n in ("CH4", "NH3", "H2O", "C2H4")]
    graphs, targets = [(a, f) for a, f, _ in data], [t for _, _, t in data]
    model = SMPBetaB200(L, C, F, D)
    flat = rng.uniform(-1, 1, model.num_params()) * 0.1
    model.set_flat_params(flat)
    tb = model.tables(graphs)

    # step 0 agrees with the reference model on every molecule (loss and feature), when the shim is available
    gf, loss, grads = model.forward_backward(tb, targets)
    if pyoracle.model_available():
        for i, (a, f, t) in enumerate(data):
            ref = pyoracle.ref_smp_beta(a, f, L, C, D, flat, t)
            assert abs(loss[i].item() - ref["loss"]) <= 1e-3 * max(1.0, ref["loss"])
    first = loss.sum().item()

    # Adam on the flat parameter vector (lr 1e-3 as in the reference test, batch = the four molecules)
    p = torch.from_numpy(flat.astype(np.float32)).cuda()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    lr, b1, b2, eps = 3e-3, 0.9, 0.999, 1e-8
    last = first
    for step in range(1, 301):
        model.set_flat_params(p.cpu().numpy())
        _, loss, g = model.forward_backward(tb, targets)
        g = g / len(targets)
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        p = p - lr * (m / (1 - b1 ** step)) / ((v / (1 - b2 ** step)).sqrt() + eps)
        last = loss.sum().item()
    assert np.isfinite(last)
    assert last < 0.02 * first, (first, last)
    pred_err = (2 * loss).sqrt().max().item()  # |predict - target| of the worst molecule
    assert pred_err < 0.5, pred_err


@pytest.mark.skipif(not pyoracle.model_available(), reason="oracle/_ref model shim not built")
def test_training_trajectory_matches_reference_batchlearn():
    """Ten epochs of the reference's own `SMP_beta::BatchLearn` (SMP_beta.h:745-772: summed gradients, then
    `Adam::Learn(lr, nBatch)` with its per-element bias correction) on the four molecules, against the batched B200 path +
    the device-side `ccn_adam_step`: same loss before every epoch and the same parameters at the end."""
    from graphflow_b200 import optim
    from graphflow_b200.model import SMPBetaB200

    L, C, F, D, epochs, lr = 2, 6, 4, 3, 10, 1e-3
    rng = np.random.default_rng(4)
    data = [molecule(n) for n in ("CH4", "NH3", "H2O", "C2H4")]
    graphs, targets = [(a, f) for a, f, _ in data], [t for _, _, t in data]
    model = SMPBetaB200(L, C, F, D)
    flat = rng.uniform(-1, 1, model.num_params()) * 0.1
    want_losses, want_params = pyoracle.ref_smp_beta_batchlearn(graphs, targets, L, C, D, flat, epochs, lr)

    model.set_flat_params(flat)
    tb = model.tables(graphs)
    p = model.get_flat_params()
    opt = optim.Adam(model.ctx, p)
    got_losses = []
    for _ in range(epochs):
        model.set_flat_params_device(p)
        _, loss, g = model.forward_backward(tb, targets)
        got_losses.append(loss.sum().item())
        opt.learn(g.contiguous(), lr, len(targets))
    got_losses = np.array(got_losses)
    assert np.abs(got_losses - want_losses[:, 0]).max() <= 1e-3 * want_losses[:, 0].max(), (got_losses, want_losses[:, 0])
    assert want_losses[-1, 0] < want_losses[0, 0]                      # it is learning
    moved = np.abs(want_params - flat).max()
    assert np.abs(p.cpu().numpy() - want_params).max() < 0.02 * moved  # same trajectory, not merely the same direction


@pytest.mark.skipif(not pyoracle.model_available(), reason="oracle/_ref model shim not built")
def test_model_api_batchlearn_predict_checkpoint(tmp_path):
    """The reference model's own call sequence (tests/test_SMP_beta.cpp:174-201): BatchLearn per epoch, Predict, save_model,
    load_model into a second network, Predict again -- through CCNModelB200, against the reference's BatchLearn returns."""
    from graphflow_b200.model import SMPBetaB200

    L, C, F, D, epochs, lr = 1, 10, 4, 5, 6, 1e-3                  # the reference test's sizes: 1 level, 10 channels, nDepth 5
    rng = np.random.default_rng(6)
    data = [molecule(n) for n in ("CH4", "NH3", "H2O", "C2H4")]
    graphs, targets = [(a, f) for a, f, _ in data], [t for _, _, t in data]
    train = SMPBetaB200(L, C, F, D)
    flat = rng.uniform(-1, 1, train.num_params()) * 0.1
    train.set_flat_params(flat)
    want, want_params = pyoracle.ref_smp_beta_batchlearn(graphs, targets, L, C, D, flat, epochs, lr)
    tb = train.tables(graphs)
    for e in range(epochs):
        before, after = train.BatchLearn(graphs, targets, lr, tb=tb)
        assert abs(before - want[e, 0]) <= 1e-3 * max(1.0, want[e, 0]) and abs(after - want[e, 1]) <= 1e-3 * max(1.0, want[e, 1])
    path = str(tmp_path / "SMP_beta.dat")
    train.save_model(path)
    test = SMPBetaB200(L, C, F, D)
    test.load_model(path)
    for g, t in zip(graphs, targets):
        assert abs(test.Predict(g) - train.Predict(g)) < 1e-3      # 6 significant digits survive the text file
    ref_loaded = pyoracle.ref_checkpoint_roundtrip(8, L, C, F, D, want_params, load_path=path)
    assert np.abs(ref_loaded - want_params).max() < 0.02 * np.abs(want_params - flat).max()
