"""Build the C-ABI shared library graphflow_b200/libccn_b200.so with nvcc for sm_100a (in-tree, so it travels
to the GPU box with the gpurun snapshot).  Usage: python -m graphflow_b200.build [--force] [--verbose]"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libccn_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(p) > t for p in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    flags = [f for f in FLAGS if f != "--use_fast_math=false"]
    cmd = [NVCC] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + sources()
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("nvcc failed building libccn_b200.so")
    if verbose:
        print(res.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
