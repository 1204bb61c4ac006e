"""One-launch Adam over flat arenas -- optional fast path for
`gaussians.optimizer.step(); gaussians.optimizer.zero_grad()` (train.py:422-424).

The reference builds `torch.optim.Adam(l, lr=0.0, eps=1e-15)` over eight per-Gaussian parameter groups with their own
learning rates (scene/gaussian_model.py:227-240) and reschedules two of them every iteration through
`optimizer.param_groups` (:251-262).  `ArenaAdam` keeps that surface -- `param_groups` is a list of dicts with `name`,
`lr` and `params` -- but lays parameters, gradients and both moments out as four flat float32 arenas of one layout, so
that (1) the data-parallel exchange is one all-reduce of the gradient arena (ibgs_b200/parallel.py) and (2) the update
is one CUDA launch that touches every value once (`ibgs_adam_step`, csrc/adam.cu).  The parameters handed back are
`torch.nn.Parameter` views into the arena whose `.grad` are views into the gradient arena: autograd accumulates into
them in place.

Densification replaces every per-Gaussian tensor and its optimizer state (scene/gaussian_model.py:362-596); after it,
build a new ArenaAdam from the new tensors with `ArenaAdam.from_state(...)` carrying the moments over.
There is no CPU or PyTorch fallback: CUDA tensors only.
"""
import ctypes as C
from collections import OrderedDict

import torch

from . import _native as N

ARENA_ALIGN = 64   # floats: groups start on 256-byte boundaries


class ArenaAdam:
    def __init__(self, named_tensors, lrs, betas=(0.9, 0.999), eps=1e-15, exp_avg=None, exp_avg_sq=None, step=0):
        """named_tensors: ordered {name: initial value (CUDA float32 tensor)}; lrs: {name: learning rate}."""
        named_tensors = OrderedDict(named_tensors)
        if not (1 <= len(named_tensors) <= N.ADAM_MAX_GROUPS):
            raise ValueError(f"1..{N.ADAM_MAX_GROUPS} parameter groups, got {len(named_tensors)}")
        first = next(iter(named_tensors.values()))
        if not first.is_cuda:
            raise RuntimeError("ibgs_b200.optim: parameters must be CUDA tensors (there is no CPU path)")
        self.device = first.device
        sizes = [int(t.numel()) for t in named_tensors.values()]
        # every group starts on a 256-byte boundary of the arenas: the rasterizer / prologue kernels read and write some
        # of these tensors (rotations and their gradients) with 16-byte vector accesses, and after densification the
        # Gaussian count is arbitrary.  The pad words are zero in all four arenas and belong to no group.
        starts, total = [], 0
        for n in sizes:
            total = (total + ARENA_ALIGN - 1) // ARENA_ALIGN * ARENA_ALIGN
            starts.append(total)
            total += n
        f32 = dict(dtype=torch.float32, device=self.device)
        self.flat_params = torch.zeros(total, **f32)
        self.flat_grads = torch.zeros(total, **f32)
        self.exp_avg = torch.zeros(total, **f32)
        self.exp_avg_sq = torch.zeros(total, **f32)
        self.betas, self.eps, self.step_count = (float(betas[0]), float(betas[1])), float(eps), int(step)
        self.params, self.grads, self.param_groups, self._ranges = OrderedDict(), OrderedDict(), [], OrderedDict()
        for (name, t), n, off in zip(named_tensors.items(), sizes, starts):
            sl = slice(off, off + n)
            self.flat_params[sl].copy_(t.detach().reshape(-1))
            p = torch.nn.Parameter(self.flat_params[sl].view(t.shape), requires_grad=True)
            p.grad = self.flat_grads[sl].view(t.shape)
            self.params[name], self.grads[name] = p, p.grad
            self._ranges[name] = (off, n)
            self.param_groups.append({"params": [p], "lr": float(lrs[name]), "name": name})
            if exp_avg is not None and name in exp_avg:
                self.exp_avg[sl].copy_(exp_avg[name].reshape(-1))
                self.exp_avg_sq[sl].copy_(exp_avg_sq[name].reshape(-1))

    @classmethod
    def from_state(cls, named_tensors, lrs, exp_avg, exp_avg_sq, step, betas=(0.9, 0.999), eps=1e-15):
        """Rebuild after densification / pruning from the new per-group tensors and their carried-over moments."""
        return cls(named_tensors, lrs, betas=betas, eps=eps, exp_avg=exp_avg, exp_avg_sq=exp_avg_sq, step=step)

    # ---- densification: the counterparts of GaussianModel's optimizer surgery (scene/gaussian_model.py:362-470) --------
    # Each returns a NEW ArenaAdam (the arenas are re-laid-out; densification happens every 100 iterations, so this is
    # off the per-step path) and keeps step_count, betas, eps and every group's current learning rate.
    def _rebuild(self, tensors, exp_avg, exp_avg_sq):
        lrs = {g["name"]: g["lr"] for g in self.param_groups}
        return ArenaAdam(tensors, lrs, betas=self.betas, eps=self.eps, exp_avg=exp_avg, exp_avg_sq=exp_avg_sq,
                         step=self.step_count)

    def prune(self, keep_mask):
        """_prune_optimizer (scene/gaussian_model.py:377-395): rows (dim 0) of every group, and of both moments, where
        keep_mask is True.  prune_points passes `~mask`."""
        keep_mask = keep_mask.to(self.device)
        t = {n: p.detach()[keep_mask] for n, p in self.params.items()}
        m = {n: self.state(n)["exp_avg"][keep_mask] for n in self.params}
        v = {n: self.state(n)["exp_avg_sq"][keep_mask] for n in self.params}
        return self._rebuild(t, m, v)

    def extend(self, tensors_dict):
        """cat_tensors_to_optimizer (scene/gaussian_model.py:420-438): new rows appended to every group, their moments
        start at zero."""
        t, m, v = {}, {}, {}
        for n, p in self.params.items():
            ext = tensors_dict[n].to(device=self.device, dtype=torch.float32)
            t[n] = torch.cat((p.detach(), ext), dim=0)
            st = self.state(n)
            m[n] = torch.cat((st["exp_avg"], torch.zeros_like(ext)), dim=0)
            v[n] = torch.cat((st["exp_avg_sq"], torch.zeros_like(ext)), dim=0)
        return self._rebuild(t, m, v)

    def reset_state(self, name, tensor):
        """replace_tensor_to_optimizer (scene/gaussian_model.py:362-375; the opacity reset): group `name` takes the
        new values in place and its moments restart from zero."""
        off, n = self._ranges[name]
        self.flat_params[off:off + n].copy_(tensor.detach().reshape(-1))
        self.exp_avg[off:off + n].zero_()
        self.exp_avg_sq[off:off + n].zero_()
        return self.params[name]

    def state(self, name):
        off, n = self._ranges[name]
        shape = self.params[name].shape
        return {"step": self.step_count, "exp_avg": self.exp_avg[off:off + n].view(shape),
                "exp_avg_sq": self.exp_avg_sq[off:off + n].view(shape)}

    def all_reduce_grads(self, async_op=False):
        """Sum of the gradient arena over ranks: one collective."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.all_reduce(self.flat_grads, op=dist.ReduceOp.SUM, async_op=async_op)
        return None

    def zero_grad(self, set_to_none=False):
        # the gradients stay views into the arena (set_to_none would detach them from it): always zero in place
        self.flat_grads.zero_()
        for name, p in self.params.items():
            if p.grad is None or p.grad.data_ptr() != self.grads[name].data_ptr():
                p.grad = self.grads[name]

    def step(self, grad_scale=1.0, zero_grads=False, skip=()):
        """One Adam update of every group with its current `param_groups[i]['lr']`; with zero_grads the gradient arena
        is cleared in the same pass (step + zero_grad of train.py:422-424 in one launch).
        `skip`: names of groups that received NO gradient this iteration.  torch.optim.Adam leaves a parameter whose
        .grad is None untouched (no moment decay, no update); the arena always holds a gradient (zeros), so the caller
        says which groups to leave out -- they are simply not part of the launch."""
        for name, p in self.params.items():
            if p.grad is None or p.grad.data_ptr() != self.grads[name].data_ptr():
                raise RuntimeError(f"ArenaAdam: .grad of '{name}' no longer aliases the gradient arena "
                                   "(use zero_grad() of this class, not set_to_none)")
        self.step_count += 1
        a = N.IbgsAdamArgs()
        a.params, a.grads = self.flat_params.data_ptr(), self.flat_grads.data_ptr()
        a.exp_avg, a.exp_avg_sq = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        groups = [g for g in self.param_groups if g["name"] not in skip]
        a.num_groups = len(groups)
        for i, g in enumerate(groups):
            off, n = self._ranges[g["name"]]
            a.groups[i].offset, a.groups[i].count, a.groups[i].lr = off, n, float(g["lr"])
        a.beta1, a.beta2, a.eps = self.betas[0], self.betas[1], self.eps
        a.step, a.grad_scale, a.zero_grads = self.step_count, float(grad_scale), int(bool(zero_grads))
        with torch.cuda.device(self.device):
            N.check(N.lib.ibgs_adam_step(C.byref(a), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                    "ibgs_adam_step")
