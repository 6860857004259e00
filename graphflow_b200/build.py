"""Build the C-ABI shared library graphflow_b200/libccn_b200.so with nvcc for sm_100a (in-tree, so it travels
to the GPU box with the gpurun snapshot).  Every .cu is compiled to its own object (in parallel, only when it or a
header / generated table it may include is newer), then linked.  Usage: python -m graphflow_b200.build [--force] [--verbose]"""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libccn_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def shared_deps():
    """Everything a translation unit may include: kernel headers, generated tables and their generators, the C-ABI
    header and the C++ class headers."""
    inc = os.path.join(HERE, "..", "include")
    pats = [os.path.join(CSRC, "*.cuh"), os.path.join(CSRC, "*.inc"), os.path.join(CSRC, "gen", "*.py"),
            os.path.join(inc, "*.h"), os.path.join(inc, "**", "*.h")]
    out = []
    for p in pats:
        out += glob.glob(p, recursive=True)
    return out


def _newest(paths):
    return max([os.path.getmtime(p) for p in paths] or [0.0])


def stale():
    if not os.path.exists(LIB):
        return True
    return _newest(sources() + shared_deps()) > os.path.getmtime(LIB)


def _compile(src, obj, verbose):
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return src, res.returncode, res.stdout


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    dep_t = _newest(shared_deps())
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or max(os.path.getmtime(src), dep_t) > os.path.getmtime(obj):
            jobs.append((src, obj))
    # objects of sources that no longer exist must not be linked
    for old in glob.glob(os.path.join(OBJ, "*.o")):
        if old not in objs:
            os.remove(old)
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as pool:
        for src, rc, out in pool.map(lambda j: _compile(j[0], j[1], verbose), jobs):
            if rc != 0:
                sys.stderr.write(out)
                raise RuntimeError("nvcc failed on %s" % os.path.basename(src))
            if verbose:
                print(out)
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("nvcc failed linking libccn_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
