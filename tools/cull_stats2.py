"""Survivor statistics for different warp sub-tile shapes (dev tool, GPU): how many (sub-tile, Gaussian) pairs pass
the exact alpha>=1/255 test, and how many lanes they light up, for 8x4 (one warp), 4x4 (half-warp) and 8x8."""
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
g = torch.Generator(device="cpu").manual_seed(0)
tiles = torch.randperm(rng.shape[0], generator=g)[:400].tolist()
ncon = st["n_contrib"].view(H, W)
tot = 0
acc = {}
for t in tiles:
    a, b = rng[t].tolist()
    if b <= a: continue
    ids = st["point_list"][a:b].long()
    xy = st["means2D"][ids]; co = st["conic_opacity"][ids]
    tx, ty = (t % gx) * 16, (t // gx) * 16
    px = torch.arange(tx, tx + 16, device="cuda").float()
    py = torch.arange(ty, ty + 16, device="cuda").float()
    dx = xy[:, 0:1, None] - px[None, None, :]            # [n,1,16]
    dy = xy[:, 1:2, None] - py[None, :, None]            # [n,16,1]
    power = -0.5 * (co[:, 0, None, None] * dx * dx + co[:, 2, None, None] * dy * dy) - co[:, 1, None, None] * dx * dy
    alpha = torch.clamp(co[:, 3, None, None] * torch.exp(power), max=0.99)
    hit = (power <= 0) & (alpha >= 1.0 / 255.0)           # [n,16(y),16(x)]
    # restrict to contributors before each pixel's last contributor (what forward/backward actually walk)
    idx = torch.arange(1, ids.numel() + 1, device="cuda")[:, None, None]
    yy = torch.arange(ty, ty + 16, device="cuda").clamp(max=H - 1); xx = torch.arange(tx, tx + 16, device="cuda").clamp(max=W - 1)
    last = ncon[yy][:, xx][None]
    hit = hit & (idx <= last)
    tot += ids.numel()
    for (sh, sw) in ((4, 8), (4, 4), (8, 8), (2, 8), (4, 16)):
        hh = hit.view(-1, 16 // sh, sh, 16 // sw, sw).any(dim=4).any(dim=2)   # [n, ny, nx] any hit in sub-tile
        lanes = hit.view(-1, 16 // sh, sh, 16 // sw, sw).sum(dim=(2, 4))
        d = acc.setdefault((sh, sw), [0, 0])
        d[0] += int(hh.sum()); d[1] += int(lanes.sum())
print("tile instances sampled", tot)
for k, (s, l) in acc.items():
    print(f"sub-tile {k[1]}x{k[0]}: survivors per instance {s/tot:.3f}, lit lanes per survivor {l/max(s,1):.2f} of {k[0]*k[1]}")
