"""The reference's parameter updates on the device (`ccn_adam_step`, `ccn_momentum_step`, graphflow_b200/optim.py) against
the fp64 restatement of Adam.h / Momentum.h that tests/test_optim_checkpoint_cpu.py pins to the compiled reference, and
the text checkpoint through the model class."""
import numpy as np
import pytest
import torch

from graphflow_b200 import optim
from oracle import pyoracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import graphflow_b200

    c = graphflow_b200.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("n_batch", [4, None])
def test_adam_matches_reference_restatement(ctx, n_batch):
    """n_batch=4: Adam::Learn(alpha, nBatch) with its per-element bias correction (Adam.h:108-137); None: Learn(alpha)."""
    rng = np.random.default_rng(11)
    n, steps = 30011, 6                       # 180 k element updates: beta2^k spans 1 .. ~1e-78
    values = rng.uniform(-1, 1, n)
    grads = rng.uniform(-1, 1, (steps, n))
    want = pyoracle.adam_reference_restatement(values.astype(np.float32), grads.astype(np.float32), 1e-3, n_batch)
    p = torch.from_numpy(values.astype(np.float32)).cuda()
    opt = optim.Adam(ctx, p)
    for s in range(steps):
        opt.learn(torch.from_numpy(grads[s].astype(np.float32)).cuda(), 1e-3, n_batch)
    got = p.cpu().numpy().astype(np.float64)
    assert np.abs(got - want).max() < 2e-6
    assert np.abs(got - values).max() > 1e-3   # the parameters really moved


def test_adam_quirk_is_reproduced_not_smoothed_over(ctx):
    """With the per-element powers the first elements of the first step move ~10x further than a textbook Adam step."""
    n = 1000
    p = torch.zeros(n, device="cuda")
    g = torch.ones(n, device="cuda")
    optim.Adam(ctx, p).learn(g, 1e-3, 1)
    got = p.cpu().numpy().astype(np.float64)
    want = pyoracle.adam_reference_restatement(np.zeros(n), np.ones((1, n)), 1e-3, 1)
    assert np.abs(got - want).max() < 1e-6
    # element i: m^ = 0.1 / (1 - 0.9^(i+1)), v^ = 0.001 / (1 - 0.999^(i+1))  ->  step != alpha except for i = 0
    assert abs(got[0] + 1e-3) < 1e-6 and abs(got[-1] + 1e-3) > 1e-4


def test_momentum_and_sgd(ctx):
    rng = np.random.default_rng(12)
    n, steps = 5000, 4
    values, grads = rng.uniform(-1, 1, n), rng.uniform(-1, 1, (steps, n))
    for gamma, cls in ((0.9, optim.Momentum), (0.0, optim.SGD)):
        p = torch.from_numpy(values.astype(np.float32)).cuda()
        opt = cls(ctx, p)
        ref_p, mom = values.copy(), np.zeros(n)
        for s in range(steps):
            opt.learn(torch.from_numpy(grads[s].astype(np.float32)).cuda(), 0.01, 8)
            mom = gamma * mom + 0.01 * grads[s] / 8          # Momentum.h:59-66; SGD.h:44-50 when gamma = 0
            ref_p -= mom
        assert np.abs(p.cpu().numpy() - ref_p).max() < 1e-6
    if pyoracle.model_available():
        want = pyoracle.ref_optimizer("momentum", values, grads, 1234, 0.01, 8)
        p = torch.from_numpy(values.astype(np.float32)).cuda()
        opt = optim.Momentum(ctx, p)
        for s in range(steps):
            opt.learn(torch.from_numpy(grads[s].astype(np.float32)).cuda(), 0.01, 8)
        assert np.abs(p.cpu().numpy() - want).max() < 1e-6


def test_model_checkpoint_roundtrip(ctx, tmp_path):
    from graphflow_b200.model import CCNModelB200

    L, C, F, D = 2, 4, 3, 2
    model = CCNModelB200("beta", L, C, F, n_depth=D, ctx=ctx)
    rng = np.random.default_rng(13)
    flat = rng.uniform(-1, 1, model.num_params()).astype(np.float32)
    model.set_flat_params(flat)
    path = str(tmp_path / "model.dat")
    model.save_model(path)
    other = CCNModelB200("beta", L, C, F, n_depth=D, ctx=ctx)
    other.load_model(path)
    got = other.get_flat_params().cpu().numpy()
    assert np.allclose(got, flat, rtol=5e-6, atol=0)          # the format keeps 6 significant digits
    if pyoracle.model_available():                             # the reference's own load_model reads the same numbers
        loaded = pyoracle.ref_checkpoint_roundtrip(8, L, C, F, D, flat.astype(np.float64), load_path=path)
        assert np.array_equal(loaded.astype(np.float32), got)
