"""CPU checks of the drop-in boundary: the shared library builds for sm_100a, loads without a GPU, exports every
symbol include/ccn_b200.h declares (and nothing the header does not), and fails loudly -- not silently -- when no
device is present."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ccn_b200.h")


@pytest.fixture(scope="module")
def lib_path():
    from graphflow_b200 import build

    return build.build()


def header_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"CCN_API\s+[\w\s\*]+?\b(ccn_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = header_symbols()
    for must in ("ccn_ctx_create", "ccn_contract18_forward", "ccn_contract18_backward", "ccn_mix_forward",
                 "ccn_mix_backward", "ccn_contract18_forward_backward_host"):
        assert must in syms


def test_library_exports_exactly_the_header(lib_path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path], text=True)
    exported = sorted(s for s in re.findall(r" T (\w+)", out) if s.startswith("ccn_"))
    assert exported == header_symbols()


def test_python_binding_matches_header(lib_path):
    from graphflow_b200 import _lib

    assert sorted(_lib.SIGNATURES) == header_symbols()
    lib = _lib.load()
    assert lib.ccn_abi_version() == 2
    assert lib.ccn_status_string(0) == b"ok" and lib.ccn_status_string(-4) == b"no usable sm_100 device"


def test_sass_contains_tma_bulk_copies(lib_path):
    sass = subprocess.check_output(["cuobjdump", "-sass", lib_path], text=True)
    assert "UBLKCP" in sass  # cp.async.bulk -> the TMA engine streams T into shared memory
    assert "sm_100a" in subprocess.check_output(["cuobjdump", "-lelf", lib_path], text=True)


def test_no_silent_fallback_without_gpu(lib_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import graphflow_b200 as gf

    lib = gf.load()
    h = ctypes.c_void_p()
    assert lib.ccn_ctx_create(ctypes.byref(h), 0) == -4 and not h.value  # CCN_ERR_NO_DEVICE
    with pytest.raises(gf.CCNError):
        gf.Context(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "graphflow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), "%s mentions the oracle" % f
