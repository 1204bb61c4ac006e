"""Generates tests/golden/*.npz on a GPU box from the UNMODIFIED reference CUDA extension (oracle/_ref).

    gpurun -- python tests/golden/make_golden.py        (writes into gpurun_out/golden/, copy to tests/golden/)

The reference ships no golden vectors; these fixtures are outputs of the reference
itself on the deterministic `tiny` / `cfg1` synthetic scenes (ibgs_b200/synthetic.py), including the
source-view depth maps (rendered with the reference's own depth-only pass).  tests/test_oracle_golden.py
checks the CPU oracle against them without a GPU; tests/test_gpu_oracle.py checks the CUDA path.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ibgs_b200 import synthetic as S  # noqa: E402
from oracle import ref_ext  # noqa: E402
import ibgs_testutil as U  # noqa: E402


def ref_src_depths(sc, bl=4):
    out = []
    for i in range(sc["nb_src"]):
        cam = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in S.src_view(sc, i).items()}
        H, W = sc["H"], sc["W"]
        sc1 = dict(sc)
        sc1.update(nb_src=1, ref_to_src_list=torch.zeros((1, 16), device="cuda"),
                   src_images=torch.zeros((1, 3, H * W), device="cuda"),
                   src_rendered_depths=torch.zeros((1, 1, H * W), device="cuda"),
                   src_cam_pos=torch.zeros((1, 3), device="cuda"))
        fw = ref_ext.forward(sc1, render_geo=False, render_depth_only=True, buffer_length=bl, cam=cam)
        out.append(fw["depth"])
    return torch.stack(out, 0).contiguous()


def npy(t):
    return t.detach().cpu().numpy()


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    for name, thr in (("tiny", 0.05), ("cfg1", 0.01)):
        sc = U.scene_to_device(S.make_scene(name))
        sc["src_rendered_depths"] = ref_src_depths(sc)
        cot = {k: v.cuda() for k, v in S.cotangents(sc).items()}
        P, N = sc["P"], sc["H"] * sc["W"]
        fw = ref_ext.forward(sc, render_geo=True, depth_error_threshold=thr)
        gr = ref_ext.backward(sc, fw, cot, render_geo=True)
        img = ref_ext.decode_image(fw["img"], N)
        geom = ref_ext.decode_geom(fw["geom"], P)
        d = dict(thr=np.float32(thr), src_rendered_depths=npy(sc["src_rendered_depths"]),
                 num_rendered=np.int64(fw["num_rendered"]), radii=npy(fw["radii"]),
                 tiles_touched=npy(geom["tiles_touched"]), n_contrib=npy(img["n_contrib"]),
                 final_T=npy(img["final_T"]), mask=npy(fw["mask"]).astype(np.uint8))
        keep_full = name == "tiny"
        for k in ("color", "normal", "depth", "cam_feat", "warped", "min_depth_diff", "camera_ray"):
            a = npy(fw[k])
            d["out_" + k] = a if keep_full else a[:, ::4, ::4].copy()   # cfg1: sub-sampled to keep the fixture small
        for k, v in gr.items():
            if k in ("colors", "cov3D"):
                continue
            a = npy(v)
            d["grad_" + k] = a if keep_full else a[::5].copy()
        # colour-only and depth-only variants
        H, W = sc["H"], sc["W"]
        sc1 = dict(sc)
        sc1.update(nb_src=1, ref_to_src_list=torch.zeros((1, 16), device="cuda"),
                   src_images=torch.zeros((1, 3, H * W), device="cuda"),
                   src_rendered_depths=torch.zeros((1, 1, H * W), device="cuda"),
                   src_cam_pos=torch.zeros((1, 3), device="cuda"))
        fc = ref_ext.forward(sc1, render_geo=False)
        d["color_only"] = npy(fc["color"]) if keep_full else npy(fc["color"])[:, ::4, ::4].copy()
        for bl in (1, 3, 4):
            fd = ref_ext.forward(sc1, render_geo=False, render_depth_only=True, buffer_length=bl)
            a = npy(fd["depth"])
            d[f"depth_only_bl{bl}"] = a if keep_full else a[:, ::4, ::4].copy()
        np.savez_compressed(os.path.join(outdir, f"ref_{name}.npz"), **d)
        print(name, "R", int(fw["num_rendered"]), "mask frac", float(fw["mask"].float().mean()))
    if ref_ext.available("knn"):
        g = torch.Generator().manual_seed(11)
        pts = torch.randn((3000, 3), generator=g) * torch.tensor([3.0, 1.0, 0.3])
        out = ref_ext.dist2(pts.cuda())
        np.savez_compressed(os.path.join(outdir, "ref_knn.npz"), points=pts.numpy(), dist2=npy(out))
        print("knn ok")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
