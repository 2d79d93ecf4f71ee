"""Build libb2nav.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

The shared library is the product: hand-written CUDA kernels plus the extern "C" layer declared in
include/b2nav.h.  It is git-ignored but travels to the GPU box with the working tree.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libb2nav.so")

# source -> extra flags.  The RBPF unit is built with -fmad=false: its cell indices and resampling ancestors must
# round like the reference's x86-64 build (no fused multiply-add), see csrc/rbpf_kernels.cuh.
SOURCES = {"api_common.cu": [], "mppi_api.cu": [], "rbpf_api.cu": ["-fmad=false"], "icp_api.cu": ["-fmad=false"]}
OBJ_DIR = os.path.join(LIB_DIR, "obj")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libb2nav cannot be built")


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "b2nav.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed building libb2nav.so")
    if verbose:
        sys.stderr.write(proc.stderr)


def build(force=False, verbose=False):
    """Compile every CUDA source (one object each, in parallel) and link ros-turtlebot-navigation_b200/lib/libb2nav.so."""
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "b2nav.h"))
    newest_header = max(os.path.getmtime(x) for x in headers)
    jobs, objs = [], []
    for src, extra in SOURCES.items():
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(path), newest_header):
            cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
            jobs.append(cmd)
    import concurrent.futures
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, len(jobs))) as ex:
        list(ex.map(lambda c: _run(c, verbose), jobs))
    _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs + ["-lnccl"], verbose)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
