"""GPU, world_size 2, NCCL: one whole data-parallel training step through every fast path of this repo --
fused parameter prologue (ibgs_b200.fused) -> rasterizer with split SH tensors -> L1 + fused SSIM loss
(ibgs_b200.loss_utils) -> backward into the ArenaAdam gradient arena -> ONE all-reduce -> one-launch Adam step
(ibgs_b200.optim).  Two ranks, each rendering its round-robin share of a 4-view batch, must end with bit-identical
parameters, and those must match the single-process step over the same 4 views.  The single-process half also runs
on a 1-GPU box; the 2-rank half is skipped there."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VIEWS = 4
LRS = {"xyz": 1.6e-4, "f_dc": 0.0025, "f_rest": 0.0025 / 20, "opacity": 0.05, "scaling": 0.005, "rotation": 0.001,
       "normal": 0.001, "offset": 1.6e-5}


def _step(view_ids, device, distributed):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    from ibgs_b200 import synthetic as S
    from ibgs_b200.fused import gaussian_prologue
    from ibgs_b200.loss_utils import ssim
    from ibgs_b200.optim import ArenaAdam
    import ibgs_b200.diff_plane_rasterization as dpr
    import ibgs_testutil as U
    sc_cpu = S.make_scene("cfg1")
    sc = U.scene_to_device(sc_cpu, device)
    sc["src_rendered_depths"] = U.render_src_depths(dpr, sc)
    P = sc["P"]
    # raw GaussianModel parameters whose activations reproduce the synthetic scene (scene/gaussian_model.py:127-147)
    op = sc["opacities"].clamp(1e-4, 1 - 1e-4)
    raw = {"xyz": sc["means3D"], "f_dc": sc["shs"][:, :1, :].contiguous(), "f_rest": sc["shs"][:, 1:, :].contiguous(),
           "opacity": torch.log(op / (1 - op)), "scaling": torch.log(sc["scales"]), "rotation": sc["rotations"],
           "normal": sc["normals_world"], "offset": torch.zeros((P, 1), device=device)}
    opt = ArenaAdam(raw, LRS)
    before = opt.flat_params.clone()
    pr = opt.params
    g = torch.Generator().manual_seed(11)
    gt = torch.rand((3, sc["H"], sc["W"]), generator=g).to(device)
    cot = {k: v.to(device) for k, v in S.cotangents(sc_cpu).items()}
    w2c = sc_cpu["w2c"].double().numpy()
    for gid in view_ids:
        rng = np.random.default_rng(1000 + gid)
        D = S._rigid(S._rot_axis_angle(rng.normal(size=3), np.radians(rng.uniform(0.0, 2.0))), rng.uniform(-0.1, 0.1, 3))
        cam = S.make_camera(D @ w2c, sc["W"], sc["H"])
        cam = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in cam.items()}
        opacity, scales, rotations, all_map = gaussian_prologue(
            pr["xyz"], pr["opacity"], pr["scaling"], pr["rotation"], pr["f_dc"], pr["f_rest"], pr["normal"], pr["offset"],
            cam["viewmatrix"], cam["campos"], concat_sh=False)
        sc_v = dict(sc)
        sc_v.update({k: cam[k] for k in ("viewmatrix", "projmatrix", "campos", "tanfovx", "tanfovy")})
        rs = U.make_settings(dpr, sc_v, render_geo=True)
        z = torch.zeros_like(sc["means3D"])
        res = dpr.GaussianRasterizer(rs)(means3D=pr["xyz"], means2D=z, means2D_abs=z, opacities=opacity, shs=pr["f_dc"],
                                         shs_rest=pr["f_rest"], scales=scales, rotations=rotations, all_map=all_map)
        image = res[0]
        loss = 0.8 * (image - gt).abs().mean() + 0.2 * (1.0 - ssim(image, gt))          # train.py:302-305
        loss = loss + 1e-3 * ((res[2] * cot["normal"]).mean() + (res[3] * cot["depth"]).mean() + (res[5] * cot["warped"]).mean())
        loss.backward()
    if distributed:
        opt.all_reduce_grads()
    grads = opt.flat_grads.clone()
    opt.step(grad_scale=1.0 / VIEWS, zero_grads=True)
    torch.cuda.synchronize(device)
    return before.cpu(), opt.flat_params.detach().cpu(), grads.cpu()


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    device = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    sys.path.insert(0, ROOT)
    from ibgs_b200 import parallel as PL
    _, after, grads = _step(PL.shard_views(VIEWS, rank, world), device, True)
    torch.save((after, grads), os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_single_process_step_moves_every_group():
    before, after, grads = _step(list(range(VIEWS)), torch.device("cuda", 0), False)
    assert torch.isfinite(after).all() and torch.isfinite(grads).all()
    assert (grads != 0).float().mean().item() > 0.3
    assert (after != before).float().mean().item() > 0.3


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_step_equals_single_process_step(tmp_path):
    import torch.multiprocessing as mp
    port = 29900 + (os.getpid() % 90)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    (a0, g0), (a1, g1) = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(g0, g1) and torch.equal(a0, a1)          # identical reduced gradients -> identical parameters
    before, single, gs = _step(list(range(VIEWS)), torch.device("cuda", 0), False)
    rel_g = ((g0 - gs).double().norm() / gs.double().norm()).item()
    assert rel_g <= 1e-3, rel_g                                   # float atomics: summation order differs
    # Adam's first step is ~ lr * sign(g): elements whose gradient is at the noise level may flip, the rest agree
    upd0, upd1 = (a0 - before).double(), (single - before).double()
    assert ((upd0 - upd1).norm() / upd1.norm()).item() <= 5e-2
