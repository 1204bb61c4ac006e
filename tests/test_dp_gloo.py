"""CPU, world_size 2, gloo: the data-parallel host logic (view sharding + flat per-Gaussian gradient
all-reduce).  The per-rank gradients come from the CPU oracle standing in for the
rasterizer; the check is that the reduced arena equals the single-process sum over the same view batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _view_grads(view_id):
    from ibgs_b200 import synthetic as S
    from oracle import oracle as O
    sc = S.make_scene("tiny", seed=100 + view_id, P=300, W=48, H=32)
    fw = O.forward(sc, render_geo=False)
    gr = O.backward(sc, fw, S.cotangents(sc, seed=view_id), render_geo=False)
    return {k: torch.from_numpy(np.ascontiguousarray(gr[k], dtype=np.float32)) for k in
            ("means3D", "sh", "opacities", "scales", "rotations")}


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ibgs_b200 import parallel as PL
    views = list(range(6))
    mine = PL.shard_views(len(views), rank, world)
    arena = PL.GradArena({"means3D": (300, 3), "sh": (300, 9, 3), "opacities": (300, 1), "scales": (300, 3),
                          "rotations": (300, 4)}, device="cpu")
    for v in mine:
        arena.accumulate(_view_grads(views[v]))
    arena.all_reduce()
    torch.save({k: t.clone() for k, t in arena.views.items()}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_view_sharding_and_gradient_allreduce(tmp_path):
    from ibgs_b200 import parallel as PL
    # sharding: disjoint, complete, balanced
    for n, w in ((8, 2), (7, 4), (3, 8), (16, 8)):
        parts = [PL.shard_views(n, r, w) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    world = 2
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    want = None
    for v in range(6):
        g = _view_grads(v)
        want = g if want is None else {k: want[k] + g[k] for k in g}
    for k in want:
        assert torch.equal(r0[k], r1[k])                      # every rank ends with the same reduced gradient
        assert torch.allclose(r0[k], want[k], rtol=1e-5, atol=1e-6), k
