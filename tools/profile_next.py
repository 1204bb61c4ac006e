"""One pass over the section-8f fast paths at the benchmark size (dev tool; the command ncu wraps for the kernels
ssim_forward_kernel / ssim_backward_kernel / adam_step_kernel / preprocess_depth_batch_kernel / prologue kernels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from ibgs_b200 import synthetic as S
import ibgs_b200.depth_batch as DB
import ibgs_b200.loss_utils as LU
import ibgs_testutil as U
import adam_bench
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3_1080p"
sc = U.scene_to_device(S.make_scene(name))
cams = []
for i in range(sc["nb_src"]):
    cam = S.src_view(sc, i)
    cams.append({k: (v.cuda() if torch.is_tensor(v) else v) for k, v in cam.items()})
st = DB.DepthBatchSettings(sc["H"], sc["W"], sc["tanfovx"], sc["tanfovy"], 1.0, torch.stack([c["viewmatrix"] for c in cams]),
                           torch.stack([c["projmatrix"] for c in cams]), 4)
for _ in range(2):
    DB.render_depth_batch(st, sc["means3D"], sc["opacities"], scales=sc["scales"], rotations=sc["rotations"],
                          normals=sc["normals_world"], camera_centers=torch.stack([c["campos"] for c in cams]))
g = torch.Generator().manual_seed(0)
gt = torch.rand((3, sc["H"], sc["W"]), generator=g).cuda()
img = (gt + 0.05 * torch.randn(gt.shape, generator=g).cuda()).clamp(0, 1).requires_grad_(True)
for _ in range(2):
    img.grad = None
    (1.0 - LU.ssim(img, gt)).backward()
adam_bench.measure(sc["P"], iters=1)
torch.cuda.synchronize()
print("done")
