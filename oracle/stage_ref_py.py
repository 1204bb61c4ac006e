"""Stage the reference's UNCHANGED Python glue next to the prebuilt reference extensions (test / bench infrastructure).

The drop-in claim of this repo is "gaussian_renderer/__init__.py, train.py and render.py run unchanged on top of
ibgs_b200".  To test that claim on the GPU box (which never sees /root/reference) the files that call the rasterizer
are copied VERBATIM -- no edits -- into baseline/_ref/py/, which is git-ignored (the history stays free of reference
sources) but NOT gpurun-ignored, so it travels to the box like oracle/_ref/*.so does.

    python oracle/stage_ref_py.py          # run where /root/reference exists; __graft_entry__.build() calls it

Staged:  gaussian_renderer/__init__.py, scene/*.py, utils/*.py, arguments/__init__.py, color_aggregation_network.py,
train.py (read by nobody; kept so the harness's train step can be diffed against it on the box) and the reference's own
autograd wrapper submodules/diff-plane-rasterization/diff_plane_rasterization/__init__.py (under py_ref_dpr/, used
only by the reference arm together with oracle/_ref/dpr/ref_dpr_C.so).
Only tests/ and bench.py read what this script stages; nothing under ibgs_b200/ does.
"""
import glob
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("IBGS_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(ROOT, "baseline", "_ref")

FILES = (["gaussian_renderer/__init__.py", "arguments/__init__.py", "color_aggregation_network.py", "train.py"]
         + ["scene/" + f for f in ("__init__.py", "gaussian_model.py", "cameras.py", "app_model.py",
                                   "dataset_readers.py", "colmap_loader.py")])
WRAPPER = "submodules/diff-plane-rasterization/diff_plane_rasterization/__init__.py"


def staged():
    return os.path.exists(os.path.join(OUT, "py", "gaussian_renderer", "__init__.py"))


def stage():
    if not os.path.isdir(REF):
        print(f"[stage_ref_py] {REF} absent (GPU box?) - using what is already staged under {OUT}")
        return staged()
    files = list(FILES) + [os.path.relpath(p, REF) for p in sorted(glob.glob(os.path.join(REF, "utils", "*.py")))]
    manifest = {}
    for rel in files:
        dst = os.path.join(OUT, "py", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    dst = os.path.join(OUT, "py_ref_dpr", "diff_plane_rasterization", "__init__.py")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copyfile(os.path.join(REF, WRAPPER), dst)
    manifest[WRAPPER] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "sha256": manifest}, f, indent=1)
    print(f"[stage_ref_py] {len(manifest)} files staged verbatim under {OUT}")
    return True


if __name__ == "__main__":
    stage()
