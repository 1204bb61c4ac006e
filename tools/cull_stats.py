"""How tight is the per-warp cull?  (dev tool, GPU)  For a sample of tiles: fraction of (sub-tile, Gaussian)
pairs surviving the bbox test, fraction of survivors with >=1 pixel passing alpha>=1/255, lane efficiency."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from ibgs_b200 import synthetic as S
import ibgs_b200.diff_plane_rasterization as dpr
import ibgs_testutil as U
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3_1080p"
sc = U.scene_to_device(S.make_scene(name))
outs, _, state = U.ours_forward_backward(dpr, sc, None, render_geo=False)
st = U.decode_ours(state)
W, H = sc["W"], sc["H"]
gx = (W + 15) // 16
rng = st["ranges"].long()
lens = (rng[:, 1] - rng[:, 0])
print("tiles", rng.shape[0], "R", state["num_rendered"], "mean len", lens.float().mean().item(), "max len", lens.max().item())
g = torch.Generator(device="cpu").manual_seed(0)
tiles = torch.randperm(rng.shape[0], generator=g)[:300].tolist()
tot_pairs = surv = surv_hit = lanes_hit = 0
exact_surv = 0
for t in tiles:
    a, b = rng[t].tolist()
    if b <= a: continue
    ids = st["point_list"][a:b].long()
    xy = st["means2D"][ids]; co = st["conic_opacity"][ids]; tau = st["cull_tau"][ids]
    tx, ty = (t % gx) * 16, (t // gx) * 16
    for wsub in range(8):
        x0, y0 = tx + (wsub & 1) * 8, ty + (wsub >> 1) * 4
        gxx, gyy, A, B, C = xy[:, 0], xy[:, 1], co[:, 0], co[:, 1], co[:, 2]
        cxp, cyp = gxx.clamp(x0, x0 + 7), gyy.clamp(y0, y0 + 3)
        ddx, ddy = gxx - cxp, gyy - cyp
        d2 = gyy - (gyy + B * ddx / C).clamp(y0, y0 + 3)
        q1 = 0.5 * (A * ddx * ddx + C * d2 * d2) + B * ddx * d2
        d1 = gxx - (gxx + B * ddy / A).clamp(x0, x0 + 7)
        q2 = 0.5 * (A * d1 * d1 + C * ddy * ddy) + B * d1 * ddy
        big = torch.full_like(q1, 3e38)
        qmin = torch.minimum(torch.where(ddx != 0, q1, big), torch.where(ddy != 0, q2, big))
        qmin = torch.where((ddx == 0) & (ddy == 0), torch.zeros_like(qmin), qmin)
        keep = ~(qmin > tau)
        px = torch.arange(x0, x0 + 8, device="cuda").float().repeat(4)
        py = torch.arange(y0, y0 + 4, device="cuda").float().repeat_interleave(8)
        dx = xy[:, 0:1] - px[None]; dy = xy[:, 1:2] - py[None]
        power = -0.5 * (co[:, 0:1] * dx * dx + co[:, 2:3] * dy * dy) - co[:, 1:2] * dx * dy
        alpha = torch.clamp(co[:, 3:4] * torch.exp(power), max=0.99)
        hit = (power <= 0) & (alpha >= 1.0 / 255.0)
        anyhit = hit.any(1)
        tot_pairs += ids.numel(); surv += int(keep.sum()); surv_hit += int((keep & anyhit).sum())
        lanes_hit += int(hit[keep].sum()); exact_surv += int(anyhit.sum())
        assert int((anyhit & ~keep).sum()) == 0, "cull rejected a contributing Gaussian!"
print(f"(warp,Gaussian) pairs {tot_pairs}: bbox survivors {surv/tot_pairs:.3f}, survivors with a hit {surv_hit/tot_pairs:.3f} "
      f"(exact-test floor {exact_surv/tot_pairs:.3f}); hit lanes per survivor {lanes_hit/max(surv,1):.1f}/32")
