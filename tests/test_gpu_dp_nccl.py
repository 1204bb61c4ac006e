"""GPU, world_size 2, NCCL: the sharded step on real devices -- the arena summed over
2 ranks (each renders its round-robin share of a 4-view batch through the CUDA rasterizer) equals the arena a single
process accumulates over the same 4 views.  Skipped on boxes with fewer than 2 GPUs."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VIEWS = 4
KEYS = ("means3D", "shs", "opacities", "scales", "rotations")


def _render_views(view_ids, device):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    from ibgs_b200 import parallel as PL
    from ibgs_b200 import synthetic as S
    import ibgs_b200.diff_plane_rasterization as dpr
    import ibgs_testutil as U
    sc_cpu = S.make_scene("cfg1")
    sc = U.scene_to_device(sc_cpu, device)
    sc["src_rendered_depths"] = U.render_src_depths(dpr, sc)
    cot = {k: v.to(device) for k, v in S.cotangents(sc_cpu).items()}
    leaf = {k: sc[k].detach().clone().requires_grad_(True) for k in KEYS}
    arena = PL.GradArena({k: tuple(v.shape) for k, v in leaf.items()}, device=device)
    for k, v in leaf.items():
        v.grad = arena.views[k]
    w2c = sc_cpu["w2c"].double().numpy()
    for gid in view_ids:
        rng = np.random.default_rng(1000 + gid)
        D = S._rigid(S._rot_axis_angle(rng.normal(size=3), np.radians(rng.uniform(0.0, 2.0))), rng.uniform(-0.1, 0.1, 3))
        cam = S.make_camera(D @ w2c, sc["W"], sc["H"])
        cam = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in cam.items()}
        sc_v = dict(sc)
        sc_v.update({k: cam[k] for k in ("viewmatrix", "projmatrix", "campos", "tanfovx", "tanfovy")})
        sc_v["all_map"] = S.all_map_for_view(sc["means3D"], sc["normals_world"], cam["viewmatrix"], cam["campos"]).contiguous()
        rs = U.make_settings(dpr, sc_v, render_geo=True)
        z = torch.zeros_like(sc["means3D"])
        res = dpr.GaussianRasterizer(rs)(means3D=leaf["means3D"], means2D=z, means2D_abs=z, opacities=leaf["opacities"],
                                         shs=leaf["shs"], scales=leaf["scales"], rotations=leaf["rotations"],
                                         all_map=sc_v["all_map"])
        torch.autograd.backward([res[0], res[2], res[3], res[5]], [cot["color"], cot["normal"], cot["depth"], cot["warped"]])
    return arena


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    device = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    sys.path.insert(0, ROOT)
    from ibgs_b200 import parallel as PL
    arena = _render_views(PL.shard_views(VIEWS, rank, world), device)
    arena.all_reduce()
    torch.cuda.synchronize()
    torch.save(arena.flat.cpu(), os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_gradient_equals_single_process_sum(tmp_path):
    import torch.multiprocessing as mp
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(r0, r1)
    single = _render_views(list(range(VIEWS)), torch.device("cuda", 0)).flat.cpu()
    rel = ((r0 - single).double().norm() / single.double().norm()).item()
    assert rel <= 1e-3, rel   # float atomics: summation order differs run to run (as in the reference)
