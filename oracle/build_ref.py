"""Build the UNMODIFIED reference CUDA extensions into oracle/_ref/ (test infrastructure only).

The sources are compiled where they lie under /root/reference (never copied into this repo);
only the resulting .so files land in oracle/_ref/, which is git-ignored but travels to the GPU
box with gpurun.  The reference does not compile as shipped on gcc 13 / CUDA 12.9 (missing
<cstdint> in cuda_rasterizer/rasterizer_impl.h:24,40-65 and <cfloat> in simple_knn.cu:90,154);
one nvcc `-include` flag each fixes that without touching a source line.

Usage:  python oracle/build_ref.py [dpr] [knn]
Only tests/, __graft_entry__.smoke() and bench.py (--impl reference / cpu_baseline) may load
what this script builds; the product path never does.
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("IBGS_REFERENCE_ROOT", "/root/reference")


def build(which):
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils.cpp_extension import load

    if which == "dpr":
        d = os.path.join(REF, "submodules", "diff-plane-rasterization")
        name, bdir = "ref_dpr_C", os.path.join(OUT, "dpr")
        srcs = [f"{d}/cuda_rasterizer/rasterizer_impl.cu", f"{d}/cuda_rasterizer/forward.cu",
                f"{d}/cuda_rasterizer/backward.cu", f"{d}/rasterize_points.cu", f"{d}/ext.cpp"]
        kw = dict(extra_include_paths=[f"{d}/third_party/glm/"],
                  extra_cuda_cflags=["-include", "cstdint"])
    elif which == "knn":
        k = os.path.join(REF, "submodules", "simple-knn")
        name, bdir = "ref_knn_C", os.path.join(OUT, "knn")
        srcs = [f"{k}/spatial.cu", f"{k}/simple_knn.cu", f"{k}/ext.cpp"]
        kw = dict(extra_cuda_cflags=["-include", "cfloat"])
    else:
        raise SystemExit(f"unknown target {which}")
    if not os.path.isdir(REF):
        print(f"[build_ref] {REF} absent (GPU box?) - using prebuilt {bdir} if present")
        return
    os.makedirs(bdir, exist_ok=True)
    t0 = time.time()
    load(name=name, sources=srcs, build_directory=bdir, verbose=False, is_python_module=False, **kw)
    print(f"[build_ref] {name} built in {time.time() - t0:.0f}s -> {bdir}/{name}.so")


if __name__ == "__main__":
    for w in (sys.argv[1:] or ["dpr", "knn"]):
        build(w)
