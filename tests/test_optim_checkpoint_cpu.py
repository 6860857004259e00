"""Host-side pieces of the "next" rows of SURVEY.md section 8(f): the reference's checkpoint text format and the
restatement of its Adam (incl. the per-element bias-correction quirk of `Learn(alpha, nBatch)`), pinned against the
compiled reference (oracle/_ref, SMP_beta::save_model / load_model, Adam.h, Momentum.h)."""

import numpy as np
import pytest

from graphflow_b200 import checkpoint
from oracle import pyoracle

needs_ref = pytest.mark.skipif(not pyoracle.model_available(), reason="oracle/_ref model shim not built")
L, C, F, D, VMAX = 2, 4, 3, 2, 8


def _params(seed=0):
    rng = np.random.default_rng(seed)
    n = pyoracle.smp_beta_num_params(L, C, F, D)
    p = rng.uniform(-1, 1, n) * 10.0 ** rng.integers(-8, 6, n)   # many magnitudes: fixed and scientific notation
    p[:6] = [0.0, -0.0, 1.0, -1.0, 123456789.0, 1e-5]
    return p


def test_value_formatting_follows_the_default_ostream_rules():
    assert checkpoint.format_value(0.5) == "0.5"
    assert checkpoint.format_value(1.0) == "1"
    assert checkpoint.format_value(123456789.0) == "1.23457e+08"
    assert checkpoint.format_value(1e-5) == "1e-05"
    assert checkpoint.format_value(-0.000123456789) == "-0.000123457"
    assert checkpoint.format_value(-0.0) == "-0"


def test_checkpoint_roundtrip(tmp_path):
    p = _params()
    path = str(tmp_path / "model.dat")
    checkpoint.save_model(path, p)
    q = checkpoint.load_model(path, p.size, np.float64)
    assert np.allclose(q, p, rtol=5e-6, atol=0)          # 6 significant digits
    checkpoint.save_model(path, q)
    assert np.array_equal(checkpoint.load_model(path, p.size, np.float64), q)   # idempotent after the first rounding
    with pytest.raises(ValueError):
        checkpoint.load_model(path, p.size + 1)


@needs_ref
def test_checkpoint_files_are_interchangeable_with_the_reference(tmp_path):
    p = _params(1)
    ours, theirs = str(tmp_path / "ours.dat"), str(tmp_path / "theirs.dat")
    checkpoint.save_model(ours, p)
    loaded_by_reference = pyoracle.ref_checkpoint_roundtrip(VMAX, L, C, F, D, p, save_path=theirs, load_path=ours)
    assert open(ours, "rb").read() == open(theirs, "rb").read()            # byte-identical writers
    assert np.array_equal(loaded_by_reference, checkpoint.load_model(theirs, p.size, np.float64))


@needs_ref
@pytest.mark.parametrize("mode,n_batch", [("adam_batch", 4), ("adam", None)])
def test_adam_restatement_matches_reference(mode, n_batch):
    rng = np.random.default_rng(2)
    n0, n1, steps = 37, 91, 5
    values = rng.uniform(-1, 1, n0 + n1)
    grads = rng.uniform(-1, 1, (steps, n0 + n1))
    want = pyoracle.ref_optimizer(mode, values, grads, n0, 1e-3, n_batch or 1)
    got = pyoracle.adam_reference_restatement(values, grads, 1e-3, n_batch)
    assert np.abs(got - want).max() < 1e-12
    if mode == "adam_batch":   # the quirk is real: a per-step bias correction gives visibly different parameters
        plain = pyoracle.adam_reference_restatement(values, grads / n_batch, 1e-3, None)
        assert np.abs(plain - want).max() > 1e-4
