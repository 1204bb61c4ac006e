"""Data-parallel host logic for multi-GPU training around the rasterizer.

The reference is single-GPU, batch = 1 view (train.py:277-292); views are independent units whose
gradients add, so the path shards naturally over camera views with replicated Gaussians:

  * `shard_views`  round-robin assignment of a view batch to ranks (no data-path collective);
  * `GradArena`    ONE flat float32 buffer holding every per-Gaussian gradient (xyz 3 + SH 3K + opacity 1 +
                   scale 3 + rotation 4 + ... floats per Gaussian); parameters' .grad can be views into it, so
                   the per-step exchange is a single NCCL all-reduce over NVLink with no pack/unpack copy.

torch.distributed is the plumbing (backend "nccl" on GPUs, "gloo" in the CPU tests).
"""
from collections import OrderedDict

import torch
import torch.distributed as dist


def shard_views(num_views, rank, world_size):
    """Indices of the views rank `rank` renders: round-robin, balanced to within one view."""
    return list(range(rank, num_views, world_size))


class GradArena:
    def __init__(self, shapes, device="cuda", dtype=torch.float32):
        self.shapes = OrderedDict(shapes)
        sizes = [int(torch.Size(s).numel()) for s in self.shapes.values()]
        self.flat = torch.zeros(sum(sizes), dtype=dtype, device=device)
        self.views = OrderedDict()
        off = 0
        for (name, shape), n in zip(self.shapes.items(), sizes):
            self.views[name] = self.flat[off:off + n].view(*shape)
            off += n

    def zero_(self):
        self.flat.zero_()

    def accumulate(self, grads):
        for k, g in grads.items():
            if k in self.views and g is not None:
                self.views[k].add_(g.view_as(self.views[k]))

    def all_reduce(self, async_op=False):
        """Sum over ranks; one collective for the whole arena."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        return None

    @property
    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()
