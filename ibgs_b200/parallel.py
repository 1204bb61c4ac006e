"""Data-parallel host logic for multi-GPU training around the rasterizer.

The reference is single-GPU, batch = 1 view (train.py:277-292); views are independent units whose gradients add, so the
path shards naturally over camera views with replicated Gaussians and NO data-path collective.  What has to cross ranks:

  * `GradArena` / `GaussianDataParallel.all_reduce_grads`   the per-Gaussian parameter gradients, ONE flat float32
        arena (xyz 3 + f_dc 3 + f_rest 3(K-1) + opacity 1 + scaling 3 + rotation 4 + normal 3 + offset 1 floats per
        Gaussian; groups start on 256-byte boundaries) summed with a single NCCL all-reduce per step, and one small
        bucket for everything else that trains (AppModel, colour-aggregation network);
  * `sync_densification_stats`   the statistics `densify_and_prune` decides on (scene/gaussian_model.py:600-604,
        train.py:400-405): gradient-norm accumulators and denominators add over views -> SUM of each rank's increment
        since the last sync; `max_radii2D` -> MAX.  Needed only right before a densification (every 100 iterations),
        not every step -- the screen-space gradients themselves (6 floats per Gaussian) never travel;
  * `sync_depth_cache`           `scene.rendered_depth_list[idx]` is overwritten by whoever renders view idx
        (train.py:299) and read by every later view that has idx as a neighbour: the entries a step's views produced
        are all-gathered so every rank keeps the same cache;
  * `seed_for_densification`     clone / split draw from torch's global RNG (gaussian_model.py:498-502,562-566): with
        identical statistics and an identical seed every rank takes the identical decision, so the replicas never
        need a parameter broadcast after densification (`check_replicas_identical` asserts it in tests).

torch.distributed is the plumbing (backend "nccl" on GPUs over NVLink / NVSwitch, "gloo" in the CPU tests).
"""
from collections import OrderedDict

import torch
import torch.distributed as dist

ALIGN_FLOATS = 64   # arena groups start on 256-byte boundaries (the kernels use 16-byte vector accesses on them)


def _world():
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def _rank():
    return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0


def shard_views(num_views, rank, world_size):
    """Indices of the views rank `rank` renders: round-robin, balanced to within one view."""
    return list(range(rank, num_views, world_size))


def _layout(shapes, align=ALIGN_FLOATS):
    offs, off = OrderedDict(), 0
    for name, shape in shapes.items():
        n = int(torch.Size(shape).numel())
        off = (off + align - 1) // align * align
        offs[name] = (off, n)
        off += n
    return offs, off


class GradArena:
    def __init__(self, shapes, device="cuda", dtype=torch.float32):
        self.shapes = OrderedDict(shapes)
        self.offsets, total = _layout(self.shapes)
        self.flat = torch.zeros(total, dtype=dtype, device=device)
        self.views = OrderedDict()
        for name, shape in self.shapes.items():
            off, n = self.offsets[name]
            self.views[name] = self.flat[off:off + n].view(*shape)

    def zero_(self):
        self.flat.zero_()

    def accumulate(self, grads):
        for k, g in grads.items():
            if k in self.views and g is not None:
                self.views[k].add_(g.view_as(self.views[k]))

    def all_reduce(self, async_op=False, upto=None):
        """Sum over ranks; one collective for the whole arena (or its first `upto` groups: everything before the
        screen-space gradient carriers, which only feed local densification statistics)."""
        if _world() > 1:
            buf = self.flat
            if upto is not None:
                off, _ = self.offsets[upto]
                buf = self.flat[:off]
            return dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=async_op)
        return None

    @property
    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()


# the eight per-Gaussian parameter tensors of GaussianModel, in the order of its Adam groups
# (scene/gaussian_model.py:227-236)
GAUSSIAN_PARAMS = OrderedDict([("xyz", "_xyz"), ("f_dc", "_features_dc"), ("f_rest", "_features_rest"),
                               ("opacity", "_opacity"), ("scaling", "_scaling"), ("rotation", "_rotation"),
                               ("normal", "_normal"), ("offset", "_offset")])
STAT_SUMS = ("xyz_gradient_accum", "xyz_gradient_accum_abs", "denom", "denom_abs")


class GaussianDataParallel:
    """View-sharded data parallelism around an (unchanged) reference `GaussianModel`.

    `attach()` re-seats the storage of the eight parameter tensors -- the SAME nn.Parameter objects the model's
    torch.optim.Adam groups hold, so optimizer state survives -- as views into one flat parameter arena, and gives every
    parameter a `.grad` that is a view into one flat gradient arena.  Autograd then accumulates a view batch in place
    and the exchange is one all-reduce with no pack / unpack copies.  After `densify_and_prune` has replaced the
    tensors, call `attach()` again.
    """

    def __init__(self, gaussians, extra_modules=(), average=True):
        self.gaussians = gaussians
        self.extra_modules = [m for m in extra_modules if m is not None]
        self.average = average
        self.attach()
        self.mark_stats_synced()

    # ---- arenas --------------------------------------------------------------------------------------------------
    def attach(self):
        gm = self.gaussians
        params = OrderedDict((name, getattr(gm, attr)) for name, attr in GAUSSIAN_PARAMS.items())
        first = next(iter(params.values()))
        offs, total = _layout(OrderedDict((k, tuple(p.shape)) for k, p in params.items()))
        self.offsets = offs
        self.flat_params = torch.zeros(total, dtype=first.dtype, device=first.device)
        self.flat_grads = torch.zeros(total, dtype=first.dtype, device=first.device)
        self.grad_views = OrderedDict()
        for name, p in params.items():
            off, n = offs[name]
            self.flat_params[off:off + n].copy_(p.detach().reshape(-1))
            p.data = self.flat_params[off:off + n].view(p.shape)        # same Parameter object, new storage
            g = self.flat_grads[off:off + n].view(p.shape)
            if p.grad is not None:
                g.copy_(p.grad)
            p.grad = g
            self.grad_views[name] = g
        self.params = params
        extra = [p for m in self.extra_modules for p in m.parameters() if p.requires_grad]
        self.extra_params = extra
        n_extra = sum(p.numel() for p in extra)
        self.extra_flat = torch.zeros(max(n_extra, 1), dtype=first.dtype, device=first.device) if extra else None
        off = 0
        for p in extra:
            g = self.extra_flat[off:off + p.numel()].view(p.shape)
            if p.grad is not None:
                g.copy_(p.grad)
            p.grad = g
            off += p.numel()
        return self

    def zero_grad(self):
        """Replacement for `optimizer.zero_grad(set_to_none=True)` (train.py:424-425,429): set_to_none would drop the
        arena views; the arenas are cleared in place instead and every .grad is re-pointed if a caller dropped it."""
        self.flat_grads.zero_()
        for name, p in self.params.items():
            if p.grad is None or p.grad.data_ptr() != self.grad_views[name].data_ptr():
                p.grad = self.grad_views[name]
        if self.extra_flat is not None:
            self.extra_flat.zero_()
            off = 0
            for p in self.extra_params:
                g = self.extra_flat[off:off + p.numel()].view(p.shape)
                if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                    p.grad = g
                off += p.numel()

    def all_reduce_grads(self, views_total=None, async_op=False):
        """Sum (or mean over `views_total` views when average=True) of every trainable gradient over the ranks:
        one collective for the per-Gaussian arena, one for the small rest."""
        handles = []
        if _world() > 1:
            handles.append(dist.all_reduce(self.flat_grads, op=dist.ReduceOp.SUM, async_op=async_op))
            if self.extra_flat is not None:
                handles.append(dist.all_reduce(self.extra_flat, op=dist.ReduceOp.SUM, async_op=async_op))
        if self.average and views_total and views_total > 1:
            if async_op:
                for h in handles:
                    if h is not None:
                        h.wait()
            self.flat_grads.mul_(1.0 / views_total)
            if self.extra_flat is not None:
                self.extra_flat.mul_(1.0 / views_total)
        return handles

    def broadcast_parameters(self, src=0):
        """Replicate rank `src`'s parameters (start of training / after loading a checkpoint)."""
        if _world() > 1:
            dist.broadcast(self.flat_params, src=src)
            for p in self.extra_params:
                dist.broadcast(p.data, src=src)

    # ---- densification statistics ----------------------------------------------------------------------------------
    def mark_stats_synced(self):
        gm = self.gaussians
        self._stat_base = {k: getattr(gm, k).clone() for k in STAT_SUMS if torch.is_tensor(getattr(gm, k, None))
                           and getattr(gm, k).numel()}

    def sync_densification_stats(self):
        """Make the statistics densify_and_prune reads identical on every rank: sums of the increments since the last
        sync (one collective over a [4, P] buffer) and the maximum of max_radii2D (gaussian_model.py:600-604,
        train.py:400-403).  Call right before densify_and_prune / after add_densification_stats of the step."""
        gm = self.gaussians
        if _world() > 1 and self._stat_base:
            names = list(self._stat_base)
            delta = torch.stack([(getattr(gm, k) - self._stat_base[k]).reshape(-1) for k in names])
            dist.all_reduce(delta, op=dist.ReduceOp.SUM)
            for i, k in enumerate(names):
                getattr(gm, k).copy_((self._stat_base[k].reshape(-1) + delta[i]).view_as(getattr(gm, k)))
            if torch.is_tensor(getattr(gm, "max_radii2D", None)) and gm.max_radii2D.numel():
                dist.all_reduce(gm.max_radii2D, op=dist.ReduceOp.MAX)
        self.mark_stats_synced()

    @staticmethod
    def seed_for_densification(iteration, base_seed=0):
        """Same RNG stream on every rank for the `torch.normal` draws of densify_and_split (gaussian_model.py:562-566)."""
        torch.manual_seed(base_seed * 1_000_003 + int(iteration))

    # ---- rendered-depth cache ----------------------------------------------------------------------------------------
    @staticmethod
    def sync_depth_cache(rendered_depth_list, local_indices):
        """All ranks end up with the same `scene.rendered_depth_list`: every rank contributes the entries its views of
        this step overwrote (train.py:299).  `local_indices` must have the same length on every rank (pad by repeating
        an index when the view batch is ragged)."""
        if _world() == 1:
            return
        idx = torch.as_tensor(list(local_indices), dtype=torch.long, device=rendered_depth_list.device)
        mine = rendered_depth_list[idx].contiguous()
        all_idx = [torch.empty_like(idx) for _ in range(_world())]
        all_val = [torch.empty_like(mine) for _ in range(_world())]
        dist.all_gather(all_idx, idx)
        dist.all_gather(all_val, mine)
        for r in range(_world()):           # rank order: a view rendered twice in one step keeps the highest rank's
            rendered_depth_list[all_idx[r]] = all_val[r]

    def check_replicas_identical(self):
        """True iff every rank holds bit-identical parameters (test helper; one MAX + one MIN all-reduce)."""
        if _world() == 1:
            return True
        hi, lo = self.flat_params.clone(), self.flat_params.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        return bool(torch.equal(hi, lo))
