"""The header-compatible C++ operators (include/graphflow_b200/ccn_ops_b200.h) against the unmodified reference CPU
operators, in C++ through the reference's own Entity / forward() / backward() API (tests/cpp/test_facade.cpp; the
procedures are those of the reference's tests/test_RisiContraction_18_gpu.cu and tests/test_MatMul_gpu.cu).  The
binaries are built by tests/cpp/Makefile (from __graft_entry__.build()) against both reference trees and travel to
the GPU box."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "cpp", "_build")


def run(binary, *args):
    path = os.path.join(BUILD, binary)
    assert os.path.exists(path), "%s missing: run `make -C tests/cpp` (needs the reference tree)" % path
    return subprocess.run([path] + [str(a) for a in args], capture_output=True, text=True, timeout=600)


@pytest.mark.gpu
@pytest.mark.parametrize("binary", ["test_facade_f32", "test_facade_f64"])
def test_facade_all_scenarios(binary):
    res = run(binary)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "facade failures=0" in res.stdout
    assert "FAIL" not in res.stdout


@pytest.mark.gpu
def test_facade_reference_test_sizes():
    """The reference's own harness sizes: `test_RisiContraction_18_gpu N C` (README usage) at the BASELINE shapes."""
    for n, c in ((8, 4), (24, 32), (32, 64)):
        res = run("test_facade_f32", "contract", n, c)
        assert res.returncode == 0, res.stdout + res.stderr


def test_facade_fails_loudly_without_gpu():
    """No CPU fallback behind the C++ classes: without a usable device the first op aborts with the C-ABI error."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    if not os.path.exists(os.path.join(BUILD, "test_facade_f32")):
        pytest.skip("facade binary not built (reference tree absent)")
    res = run("test_facade_f32", "contract", 4, 2)
    assert res.returncode != 0
    assert "ccn_ctx_create failed" in res.stderr


@pytest.mark.gpu
def test_cpp_model_facade_matches_reference_model():
    """ccn_b200::SMP_beta (include/graphflow_b200/SMP_beta_b200.h) against the unmodified ::SMP_beta built from the same
    srand() seed: identical parameters, Feature / Predict / getLoss within 1e-4, five epochs of BatchLearn with the same loss
    pairs and final parameters, and a checkpoint of ours read by the reference's load_model (tests/cpp/test_model.cpp)."""
    res = run("test_model_f64")
    assert res.returncode == 0, res.stdout + res.stderr
    assert "model failures=0" in res.stdout
    assert "FAIL" not in res.stdout
    for scenario in ("omega_Predict", "physics_BatchLearn_parameters", "pairgraphs_BatchLearn_parameters", "ver8_same_seed_parameters"):  # every facade ran
        assert scenario in res.stdout, scenario
    launches = [int(ln.split("=")[1]) for ln in res.stdout.splitlines() if ln.startswith("model kernel_launches=")]
    assert launches and min(launches) > 0          # the device path ran (there is no CPU path to fall back to)
