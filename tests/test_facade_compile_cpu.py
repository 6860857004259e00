"""Compile-only checks of the C++ facade header against the UNMODIFIED reference trees (no GPU needed; skipped where the
reference tree is not mounted): the `-DCCN_B200_DROP_IN` spelling, under which a translation unit written against
GraphFlow_gpu's class names (`RisiContraction_18_gpu`, `MatMul_gpu`) compiles unchanged, for the double and the float tree."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "GraphFlow")) or shutil.which("g++") is None,
                    reason="needs the reference tree and g++")
@pytest.mark.parametrize("tree", ["GraphFlow", "GraphFlow_32bit"])
def test_drop_in_spelling_compiles(tree):
    src = os.path.join(ROOT, "tests", "cpp", "dropin_compile_check.cpp")
    r = subprocess.run(["g++", "-std=c++11", "-fsyntax-only", "-w", "-I" + os.path.join(REF, tree), "-I" + os.path.join(ROOT, "include"), src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
