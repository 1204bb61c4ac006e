"""CPU, world_size 2, gloo: `ibgs_b200.parallel.GaussianDataParallel` -- the data-parallel pieces beyond the gradient
all-reduce: parameter / gradient arenas re-seated under an unchanged GaussianModel-like object
and its torch.optim.Adam, densification-statistics reduction incl. the max_radii2D maximum
(scene/gaussian_model.py:600-604, train.py:400-405), rendered_depth_list exchange (train.py:299), identical densify
decisions on every rank (seeded torch.normal, gaussian_model.py:562-566), re-attach after the tensors were replaced."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P, K = 37, 9           # odd Gaussian count on purpose: arena groups must still start on aligned boundaries


class FakeGaussians:
    """The attributes GaussianDataParallel touches, with the reference's names (scene/gaussian_model.py:56-76,218-240)."""

    def __init__(self, seed=0, n=P):
        g = torch.Generator().manual_seed(seed)
        mk = lambda *s: torch.nn.Parameter(torch.randn(*s, generator=g))
        self._xyz, self._features_dc, self._features_rest = mk(n, 3), mk(n, 1, 3), mk(n, K - 1, 3)
        self._opacity, self._scaling, self._rotation = mk(n, 1), mk(n, 3), mk(n, 4)
        self._normal, self._offset = mk(n, 3), mk(n, 1)
        self.xyz_gradient_accum = torch.zeros(n, 1)
        self.xyz_gradient_accum_abs = torch.zeros(n, 1)
        self.denom = torch.zeros(n, 1)
        self.denom_abs = torch.zeros(n, 1)
        self.max_radii2D = torch.zeros(n)
        names = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation", "normal", "offset")
        attrs = ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation", "_normal", "_offset")
        self.optimizer = torch.optim.Adam([{"params": [getattr(self, a)], "lr": 0.01, "name": n_}
                                           for n_, a in zip(names, attrs)], lr=0.0, eps=1e-15)


def _view_loss(gm, extra, view_id):
    g = torch.Generator().manual_seed(100 + view_id)
    loss = 0.0
    for a in ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation", "_normal", "_offset"):
        p = getattr(gm, a)
        loss = loss + (p * torch.randn(p.shape, generator=g)).sum() + 0.5 * (p * p).sum() * (view_id + 1) * 1e-2
    loss = loss + (extra.weight * torch.randn(extra.weight.shape, generator=g)).sum()
    return loss


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ibgs_b200 import parallel as PL
    gm = FakeGaussians(seed=0)
    extra = torch.nn.Linear(3, 2)
    torch.manual_seed(1)
    with torch.no_grad():
        extra.weight.copy_(torch.randn(2, 3))
        extra.bias.zero_()
    old_params = [gm._xyz, gm._rotation]
    dp = PL.GaussianDataParallel(gm, extra_modules=(extra, None))
    assert gm._xyz is old_params[0] and gm._rotation is old_params[1]          # same Parameter objects (optimizer keeps them)
    for name, (off, n) in dp.offsets.items():
        assert off % PL.ALIGN_FLOATS == 0, name
    assert gm._rotation.data_ptr() % 16 == 0 and gm._rotation.grad.data_ptr() % 16 == 0
    dp.broadcast_parameters(0)
    views = list(range(6))
    mine = PL.shard_views(len(views), rank, world)
    for v in mine:
        _view_loss(gm, extra, views[v]).backward()                        # accumulates in place into the arenas
    assert gm._xyz.grad.data_ptr() == dp.grad_views["xyz"].data_ptr()
    dp.all_reduce_grads(views_total=len(views))
    grads = {k: g.clone() for k, g in dp.grad_views.items()}
    grads["extra_w"] = extra.weight.grad.clone()

    # densification statistics: each rank saw different views
    g = torch.Generator().manual_seed(50 + rank)
    for _ in range(3):
        m = torch.rand(P, generator=g) < 0.5
        gm.xyz_gradient_accum[m] += torch.rand(int(m.sum()), 1, generator=g)
        gm.xyz_gradient_accum_abs[m] += torch.rand(int(m.sum()), 1, generator=g)
        gm.denom[m] += 1
        gm.denom_abs[m] += 1
        gm.max_radii2D[m] = torch.max(gm.max_radii2D[m], torch.rand(int(m.sum()), generator=g) * 30)
    local_stats = {k: getattr(gm, k).clone() for k in PL.STAT_SUMS + ("max_radii2D",)}
    dp.sync_densification_stats()
    stats = {k: getattr(gm, k).clone() for k in PL.STAT_SUMS + ("max_radii2D",)}
    # a second sync without new local increments must not change anything (increments are relative to the last sync)
    dp.sync_densification_stats()
    for k in stats:
        assert torch.equal(stats[k], getattr(gm, k)), k

    # rendered-depth cache: rank r rendered views r and r + 2 of 4
    cache = torch.zeros(4, 1, 5, 7)
    idx = [rank, rank + 2]
    for i in idx:
        cache[i] = float(10 * rank + i + 1)
    dp.sync_depth_cache(cache, idx)

    # optimizer step on the reduced gradients, then "densification" with the shared seed: identical replicas
    gm.optimizer.step()
    dp.zero_grad()
    assert float(dp.flat_grads.abs().sum()) == 0.0 and gm._xyz.grad.data_ptr() == dp.grad_views["xyz"].data_ptr()
    dp.seed_for_densification(iteration=700)
    noise = torch.normal(mean=torch.zeros(P, 3), std=torch.ones(P, 3))
    keep = stats["denom"].squeeze(1) > 0                                  # decision taken from the SYNCED statistics
    for a in ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation", "_normal", "_offset"):
        t = getattr(gm, a).detach()[keep]
        if a == "_xyz":
            t = t + 0.01 * noise[keep]
        setattr(gm, a, torch.nn.Parameter(t.clone()))
    dp.attach()                                                           # tensors were replaced: new arenas
    same = dp.check_replicas_identical()
    torch.save(dict(grads=grads, stats=stats, local_stats=local_stats, cache=cache, same=same,
                    params=dp.flat_params.clone(), n=int(keep.sum())), os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_gaussian_data_parallel_world2(tmp_path):
    world = 2
    port = 29800 + (os.getpid() % 150)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    # gradients: every rank holds the mean over all six views == single-process result
    gm = FakeGaussians(seed=0)
    extra = torch.nn.Linear(3, 2)
    torch.manual_seed(1)
    with torch.no_grad():
        extra.weight.copy_(torch.randn(2, 3))
    for v in range(6):
        _view_loss(gm, extra, v).backward()
    want = {"xyz": gm._xyz.grad, "f_dc": gm._features_dc.grad, "f_rest": gm._features_rest.grad,
            "opacity": gm._opacity.grad, "scaling": gm._scaling.grad, "rotation": gm._rotation.grad,
            "normal": gm._normal.grad, "offset": gm._offset.grad, "extra_w": extra.weight.grad}
    for k, w in want.items():
        assert torch.equal(r0["grads"][k], r1["grads"][k]), k
        assert torch.allclose(r0["grads"][k], w / 6.0, rtol=1e-5, atol=1e-6), k
    # statistics: sums add, max_radii2D takes the maximum, identical on both ranks
    from ibgs_b200 import parallel as PL
    for k in PL.STAT_SUMS:
        assert torch.equal(r0["stats"][k], r1["stats"][k]), k
        assert torch.allclose(r0["stats"][k], r0["local_stats"][k] + r1["local_stats"][k]), k
    assert torch.equal(r0["stats"]["max_radii2D"], torch.max(r0["local_stats"]["max_radii2D"], r1["local_stats"]["max_radii2D"]))
    assert torch.equal(r0["stats"]["max_radii2D"], r1["stats"]["max_radii2D"])
    # depth cache: both ranks hold all four entries
    assert torch.equal(r0["cache"], r1["cache"])
    assert [float(r0["cache"][i].flatten()[0]) for i in range(4)] == [1.0, 12.0, 3.0, 14.0]
    # replicas identical after the seeded "densification" and re-attach
    assert r0["same"] and r1["same"] and r0["n"] == r1["n"]
    assert torch.equal(r0["params"], r1["params"])
