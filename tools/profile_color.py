"""One forward+backward of the colour-aggregation fast path at cfg3's image size (dev tool; the command ncu wraps for
color_features_forward_kernel / color_features_backward_kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import colorfeat_ref as CR
from ibgs_b200 import color_aggregation as CA
H, W = 822, 1237
torch.manual_seed(0)
mlp = torch.nn.Sequential(torch.nn.Linear(7, 32), torch.nn.ReLU(), torch.nn.Linear(32, 32), torch.nn.ReLU()).cuda()
pkg = CR.random_render_pkg(H, W, seed=1, device="cuda")
for _ in range(3):
    leaves = {k: pkg[k].clone().requires_grad_(True) for k in ("render", "warped_image")}
    x = CA.color_features(leaves["warped_image"], pkg["cam_feat"], leaves["render"], pkg["camera_ray"].view(3, H, W), mlp, 3)
    x.float().square().mean().backward()
torch.cuda.synchronize()
print("done")
