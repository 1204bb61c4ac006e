"""Build libibgs_b200.so (sm_100a) in-tree with nvcc.  No torch headers are involved: the library is a
plain C-ABI CUDA shared object (include/ibgs_b200.h) that Python loads with ctypes.

    python -m ibgs_b200.build [--force] [--verbose]
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT_DIR, "libibgs_b200.so")
SOURCES = ["api.cu", "host_api.cu", "preprocess.cu", "binning.cu", "sort.cu", "textures.cu", "render_forward.cu",
           "render_backward.cu", "preprocess_backward.cu", "knn.cu", "prologue.cu", "ssim.cu", "adam.cu", "color_features.cu", "nhwc_ops.cu", "depth_normal.cu", "densify_stats.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "..", "include", "ibgs_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# -fmad=true / precise div+sqrt are nvcc defaults, i.e. exactly the reference's flags (DPR/setup.py:21-29);
# do NOT add -use_fast_math: bit-exact radii / tile keys depend on IEEE division and sqrt.
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _digest():
    h = hashlib.sha256()
    for p in [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _compile(src, verbose):
    obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
    cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    with open(obj + ".ptxas.log", "w") as f:
        f.write(r.stderr)
    return obj


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp = os.path.join(OUT_DIR, "build.sha256")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB
    if not os.path.exists(NVCC):
        if os.path.exists(LIB):
            return LIB  # GPU box without a toolkit would still use the shipped .so
        raise RuntimeError(f"nvcc not found at {NVCC} and no prebuilt {LIB}")
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
