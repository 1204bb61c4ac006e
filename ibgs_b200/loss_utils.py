"""Fused SSIM for the IBGS loss step -- same names, arguments and return values as the
SSIM functions of the reference's `utils/loss_utils.py`:

    ssim(img1, img2, window_size=11, size_average=True)                       reference :34-65
    compute_photometric_ssim(img1, img2, window_size=11, size_average=True)   reference :67-90
    ssim2(img1, img2, window_size=11)                                         reference :92-117

`train.py` calls them six times per iteration (:302 rendered image, :330 once per warped source view, :355 aggregated
image); in the reference each call is five depthwise 11x11 cuDNN convolutions plus ~15 elementwise kernels forward
and about twice that in autograd's backward.  Here the SSIM map is ONE CUDA launch and its backward ONE launch
(`ibgs_ssim_forward` / `ibgs_ssim_backward`, csrc/ssim.cu).  Gradients flow to img1, img2 or both, like autograd's.
Only the reference's window (11 taps, sigma 1.5) is implemented; any other `window_size` raises.  CUDA tensors only:
there is no CPU or PyTorch fallback.
"""
import ctypes as C

import torch

from . import _native as N


class _SsimMap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img1, img2):
        if not (img1.is_cuda and img2.is_cuda):
            raise RuntimeError("ibgs_b200.loss_utils: images must be CUDA tensors (there is no CPU path)")
        if img1.shape != img2.shape or img1.dim() < 2:
            raise RuntimeError(f"ssim needs two images of one shape [...,H,W], got {tuple(img1.shape)} and {tuple(img2.shape)}")
        device = img1.device
        x = img1.detach().to(torch.float32).contiguous()
        y = img2.detach().to(device=device, dtype=torch.float32).contiguous()
        H, W = x.shape[-2], x.shape[-1]
        planes = x.numel() // (H * W) if H * W else 0
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        # the kernels are written for "img1 differentiable" (3 partial planes) or "both" (4): a differentiable img2
        # alone is the mirrored call -- the map is symmetric in its two arguments
        swap = need2 and not need1
        if swap:
            x, y = y, x
        nparts = 0 if not (need1 or need2) else (4 if (need1 and need2) else 3)
        out = torch.empty_like(x)
        parts = torch.empty((nparts,) + tuple(x.shape), dtype=torch.float32, device=device)
        a = N.IbgsSsimArgs()
        a.planes, a.height, a.width = planes, H, W
        a.img1, a.img2, a.ssim_map = x.data_ptr(), y.data_ptr(), out.data_ptr()
        if nparts:
            a.dm_dmu1, a.dm_de11, a.dm_de12 = parts[0].data_ptr(), parts[1].data_ptr(), parts[2].data_ptr()
        if nparts == 4:
            a.dm_dmu2 = parts[3].data_ptr()
        if planes:
            with torch.cuda.device(device):
                N.check(N.lib.ibgs_ssim_forward(C.byref(a), C.c_void_p(torch.cuda.current_stream(device).cuda_stream)),
                        "ibgs_ssim_forward")
        ctx.swap, ctx.nparts = swap, nparts
        ctx.dtypes = (img1.dtype, img2.dtype)
        ctx.save_for_backward(x, y, parts)
        return out.to(img1.dtype)

    @staticmethod
    def backward(ctx, g):
        x, y, parts = ctx.saved_tensors
        device = x.device
        H, W = x.shape[-2], x.shape[-1]
        planes = x.numel() // (H * W) if H * W else 0
        # a cotangent that is one broadcast scalar (the `.mean()` case: autograd hands over an expanded 0-dim tensor) stays
        # a single device float -- no materialised plane, no host read-back
        if g.numel() and all(s == 0 for s in g.stride()):
            gten, is_scalar = g.as_strided((1,), (1,)).to(torch.float32), 1
        else:
            gten, is_scalar = g.to(torch.float32).contiguous(), 0
        d1 = torch.empty_like(x)
        d2 = torch.empty_like(x) if ctx.nparts == 4 else None
        a = N.IbgsSsimArgs()
        a.planes, a.height, a.width = planes, H, W
        a.img1, a.img2 = x.data_ptr(), y.data_ptr()
        a.dm_dmu1, a.dm_de11, a.dm_de12 = parts[0].data_ptr(), parts[1].data_ptr(), parts[2].data_ptr()
        if ctx.nparts == 4:
            a.dm_dmu2 = parts[3].data_ptr()
            a.dL_dimg2 = d2.data_ptr()
        a.dL_dmap, a.dL_dmap_is_scalar, a.dL_dmap_scale = gten.data_ptr(), is_scalar, 1.0
        a.dL_dimg1 = d1.data_ptr()
        if planes:
            with torch.cuda.device(device):
                N.check(N.lib.ibgs_ssim_backward(C.byref(a), C.c_void_p(torch.cuda.current_stream(device).cuda_stream)),
                        "ibgs_ssim_backward")
        if ctx.swap:
            return None, d1.to(ctx.dtypes[1])
        return d1.to(ctx.dtypes[0]), (d2.to(ctx.dtypes[1]) if d2 is not None else None)


def ssim_map(img1, img2):
    """The per-pixel SSIM map of the reference's `_ssim` (utils/loss_utils.py:46-60), shape of the inputs."""
    return _SsimMap.apply(img1, img2)


def _check_window(window_size):
    if window_size != 11:
        raise NotImplementedError("ibgs_b200.loss_utils implements the reference's 11-tap window only")


def ssim(img1, img2, window_size=11, size_average=True):
    """utils/loss_utils.py:34-65."""
    _check_window(window_size)
    m = ssim_map(img1, img2)
    if size_average:
        return m.mean()
    return m.mean(1).mean(1).mean(1)


def compute_photometric_ssim(img1, img2, window_size=11, size_average=True):
    """utils/loss_utils.py:67-90: the map itself unless size_average."""
    _check_window(window_size)
    m = ssim_map(img1, img2)
    return m.mean() if size_average else m


def ssim2(img1, img2, window_size=11):
    """utils/loss_utils.py:92-117."""
    _check_window(window_size)
    return ssim_map(img1, img2).mean(0)
