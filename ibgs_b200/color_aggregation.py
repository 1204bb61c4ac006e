"""Optional fast path for the colour-aggregation step.

    from ibgs_b200.color_aggregation import fuse_color        # instead of color_aggregation_network.fuse_color

`fuse_color` has the reference's signature and returns the reference's dict (color_aggregation_network.py:156-250); the
`color_aggregation_network` argument is the reference's own, unchanged `ColorFusionResidualNet` -- its parameters stay
the leaves the caller's optimizer updates.  What changes is how the numbers are produced:

  * feature assembly (:196-206) + per_view_mlp + view aggregation + the cat that builds the conv decoder's input
    (:121-131) are ONE hand-written CUDA kernel per direction (csrc/color_features.cu through the C ABI), writing the
    decoder's input directly as NHWC with 40 channels (38 + 2 zero), bf16 when `precision="bf16"`;
  * the conv decoder (ConvDecoderAE, :6-68) runs through cuDNN on that layout with every channel count rounded up to a
    multiple of 8 (38 -> 40, 19 -> 24, 9 -> 16): zero-padded copies of the module's weights made inside each layer's
    autograd function, whose backward un-pads the weight gradient again, so gradients reach the unpadded parameters.  The reference's shapes (38 / 19 / 9 / 76 channels, NCHW fp32) do
    not meet the tensor-core kernels' alignment and run on SIMT / tf32 implicit-GEMM fall-backs: 10.0 ms forward+backward
    at 1237x822; padded NHWC bf16: 6.3 ms, fp32 (tf32) 6.7 ms (tools/colornet_bench.py, profiles/NOTES.md);
  * the exposure affine (:136-153) solves the same least-squares problem through its 4x4 normal equations in float64
    from masked sums instead of boolean-gathering the valid pixels (no host sync, no (N_valid x 4) matrices).

precision="fp32" keeps float32 storage (cuDNN may use tf32 in the convolutions if torch allows it); "bf16" (the
default, what the reference's own `enable_mix_precision` flag does at test time, render.py:137) stores activations in
bf16.  residual_resolution_scale != 1 and per_view_feat_dim != 32 are not covered by the kernel: NotImplementedError
(use the reference's function for those).  There is no CPU path.
"""
import ctypes as C

import torch
import torch.nn.functional as F

from . import _native as N

CHANNEL_PITCH = 40
_MODES = {"mean": 0, "max": 1}


def _f32c(t):
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class _ColorFeatures(torch.autograd.Function):
    """(warped, cam_feat, rendered, camera_ray, per_view_mlp weights) -> conv-decoder input (1, 40, H, W), channels_last."""

    @staticmethod
    def forward(ctx, warped, cam_feat, rendered, camera_ray, w1, b1, w2, b2, n_views, mode, bf16):
        if not rendered.is_cuda:
            raise RuntimeError("ibgs_b200.color_aggregation: tensors must be CUDA tensors (there is no CPU path)")
        _, H, W = rendered.shape
        dev = rendered.device
        ins = [_f32c(t) for t in (warped, cam_feat, rendered, camera_ray, w1, b1, w2, b2)]
        if ins[0].numel() < n_views * 3 * H * W or ins[1].numel() < n_views * 4 * H * W:
            raise ValueError("warped / cam_feat hold fewer than n_views views")
        if tuple(ins[4].shape) != (32, 7) or tuple(ins[6].shape) != (32, 32):
            raise NotImplementedError("the fused kernel covers per_view_mlp = Linear(7,32), Linear(32,32)")
        buf = torch.empty((1, H, W, CHANNEL_PITCH), dtype=torch.bfloat16 if bf16 else torch.float32, device=dev)
        a = N.IbgsColorFeatArgs()
        a.height, a.width, a.n_views, a.mode, a.channel_pitch, a.bf16 = H, W, n_views, mode, CHANNEL_PITCH, int(bf16)
        for name, t in zip(("warped", "cam_feat", "rendered", "camera_ray", "w1", "b1", "w2", "b2"), ins):
            setattr(a, name, t.data_ptr())
        a.cnn_input = buf.data_ptr()
        with torch.cuda.device(dev):
            N.check(N.lib.ibgs_color_features_forward(C.byref(a), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                    "ibgs_color_features_forward")
        ctx.save_for_backward(*ins)
        ctx.cfg = (H, W, n_views, mode, bool(bf16))
        ctx.shapes = (warped.shape, rendered.shape)
        return buf.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        H, W, n_views, mode, bf16 = ctx.cfg
        ins = ctx.saved_tensors
        dev = ins[2].device
        g = g.to(torch.bfloat16 if bf16 else torch.float32).permute(0, 2, 3, 1).contiguous()   # NHWC storage
        need = ctx.needs_input_grad
        fopt = dict(dtype=torch.float32, device=dev)
        d_warped = None
        if need[0]:
            d_warped = torch.zeros(ctx.shapes[0], **fopt)      # views past n_views get no gradient
        d_rendered = torch.empty(ctx.shapes[1], **fopt) if need[2] else None
        dw = torch.zeros(32 * 7 + 32 + 32 * 32 + 32, **fopt)
        a = N.IbgsColorFeatArgs()
        a.height, a.width, a.n_views, a.mode, a.channel_pitch, a.bf16 = H, W, n_views, mode, CHANNEL_PITCH, int(bf16)
        for name, t in zip(("warped", "cam_feat", "rendered", "camera_ray", "w1", "b1", "w2", "b2"), ins):
            setattr(a, name, t.data_ptr())
        a.g_cnn_input = g.data_ptr()
        a.d_warped = d_warped.data_ptr() if d_warped is not None else None
        a.d_rendered = d_rendered.data_ptr() if d_rendered is not None else None
        base = dw.data_ptr()
        a.d_w1, a.d_b1, a.d_w2, a.d_b2 = base, base + 4 * 224, base + 4 * 256, base + 4 * (256 + 1024)
        with torch.cuda.device(dev):
            N.check(N.lib.ibgs_color_features_backward(C.byref(a), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                    "ibgs_color_features_backward")
        return (d_warped, None, d_rendered, None, dw[:224].view(32, 7), dw[224:256], dw[256:1280].view(32, 32),
                dw[1280:1312], None, None, None)


def color_features(warped, cam_feat, rendered, camera_ray, per_view_mlp, n_views, mode="mean", bf16=True):
    """Conv-decoder input of ColorFusionResidualNet for one view: (1, 40, H, W) channels_last (see the module docstring).
    warped (>= n_views*3, H, W) / cam_feat (>= n_views*4, H, W) as the rasterizer returns them, rendered / camera_ray (3, H, W),
    per_view_mlp the reference's nn.Sequential(Linear(7,32), ReLU, Linear(32,32), ReLU)."""
    l1, l2 = per_view_mlp[0], per_view_mlp[2]
    return _ColorFeatures.apply(warped, cam_feat, rendered, camera_ray, l1.weight, l1.bias, l2.weight, l2.bias,
                                int(n_views), _MODES[mode], bool(bf16))


# ---- conv decoder on tensor-core-aligned shapes ------------------------------------------------------------------------
def _pad8(c):
    return (c + 7) // 8 * 8


def _nhwc(t):
    """(1, C, H, W) tensor -> its (H, W, C)-contiguous storage view (copying only if it is not channels_last already)."""
    v = t.permute(0, 2, 3, 1)
    return v if v.is_contiguous() else v.contiguous()


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _fast_glue_ok(t):
    return t.is_cuda and t.dim() == 4 and t.shape[0] == 1 and t.shape[1] % 8 == 0 and t.dtype in (torch.bfloat16, torch.float32)


class _MaxPool2(torch.autograd.Function):
    """nn.MaxPool2d(2) on an NHWC tensor (csrc/nhwc_ops.cu): same values and gradient routing as torch."""

    @staticmethod
    def forward(ctx, x):
        xv = _nhwc(x)
        _, H, W, Cn = xv.shape
        y = torch.empty((1, H // 2, W // 2, Cn), dtype=x.dtype, device=x.device)
        idx = torch.empty((H // 2, W // 2, Cn), dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            N.check(N.lib.ibgs_nhwc_maxpool2_forward(xv.data_ptr(), y.data_ptr(), idx.data_ptr(), H, W, Cn,
                                                     int(x.dtype == torch.bfloat16), _stream(x.device)), "ibgs_nhwc_maxpool2_forward")
        ctx.save_for_backward(idx)
        ctx.shape = (H, W, Cn)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        H, W, Cn = ctx.shape
        gv = _nhwc(g)
        gx = torch.empty((1, H, W, Cn), dtype=g.dtype, device=g.device)
        with torch.cuda.device(g.device):
            N.check(N.lib.ibgs_nhwc_maxpool2_backward(gv.data_ptr(), idx.data_ptr(), gx.data_ptr(), H, W, Cn,
                                                      int(g.dtype == torch.bfloat16), _stream(g.device)), "ibgs_nhwc_maxpool2_backward")
        return gx.permute(0, 3, 1, 2)


class _UpsampleNearest(torch.autograd.Function):
    """F.interpolate(x, size=(Ho, Wo), mode="nearest") on an NHWC tensor (csrc/nhwc_ops.cu)."""

    @staticmethod
    def forward(ctx, x, Ho, Wo):
        xv = _nhwc(x)
        _, Hi, Wi, Cn = xv.shape
        y = torch.empty((1, Ho, Wo, Cn), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            N.check(N.lib.ibgs_nhwc_upsample_cat_forward(xv.data_ptr(), None, y.data_ptr(), Hi, Wi, Ho, Wo, Cn, 0,
                                                         int(x.dtype == torch.bfloat16), _stream(x.device)),
                    "ibgs_nhwc_upsample_cat_forward")
        ctx.shape = (Hi, Wi, Ho, Wo, Cn)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        Hi, Wi, Ho, Wo, Cn = ctx.shape
        gv = _nhwc(g)
        ga = torch.empty((1, Hi, Wi, Cn), dtype=g.dtype, device=g.device)
        with torch.cuda.device(g.device):
            N.check(N.lib.ibgs_nhwc_upsample_backward(gv.data_ptr(), ga.data_ptr(), Hi, Wi, Ho, Wo, Cn, Cn,
                                                      int(g.dtype == torch.bfloat16), _stream(g.device)),
                    "ibgs_nhwc_upsample_backward")
        return ga.permute(0, 3, 1, 2), None, None


def max_pool2(x):
    """nn.MaxPool2d(2)(x) for a (1, C, H, W) channels_last CUDA tensor with C % 8 == 0 (bf16 / float32)."""
    if not _fast_glue_ok(x):
        raise RuntimeError("ibgs_b200.color_aggregation.max_pool2: (1, C % 8 == 0, H, W) CUDA bf16 / float32 tensors only")
    return _MaxPool2.apply(x)


def upsample_nearest(x, size):
    """F.interpolate(x, size=size, mode="nearest") for a (1, C, H, W) channels_last CUDA tensor with C % 8 == 0."""
    if not _fast_glue_ok(x):
        raise RuntimeError("ibgs_b200.color_aggregation.upsample_nearest: (1, C % 8 == 0, H, W) CUDA bf16 / float32 tensors only")
    return _UpsampleNearest.apply(x, int(size[0]), int(size[1]))


def _pad_wb(w, b, segs, dtype):
    """Zero-padded copies (no autograd graph) of a conv's weight / bias: output channels and every input segment rounded up
    to a multiple of 8 channels, weight in channels_last, both in `dtype`."""
    co, _, kh, kw = w.shape
    cop, cip = _pad8(co), sum(_pad8(c) for c in segs)
    wp = torch.zeros((cop, cip, kh, kw), dtype=dtype, device=w.device).contiguous(memory_format=torch.channels_last)
    at = atp = 0
    for c in segs:
        wp[:co, atp:atp + c].copy_(w[:, at:at + c])
        at += c
        atp += _pad8(c)
    bp = torch.zeros(cop, dtype=dtype, device=w.device)
    bp[:co].copy_(b)
    return wp, bp


def _unpad_w(gwp, co, segs, dtype):
    parts, atp = [], 0
    for c in segs:
        parts.append(gwp[:co, atp:atp + c])
        atp += _pad8(c)
    g = parts[0] if len(parts) == 1 else torch.cat(parts, 1)
    return g.to(dtype).contiguous()


class _ConvLayer(torch.autograd.Function):
    """One conv (+ bias) (+ ReLU) layer of the decoder on channel-padded NHWC activations, taking the module's UNPADDED
    weight / bias: the padding is done inside (a handful of copies, no autograd nodes), the forward is cuDNN's fused
    conv + bias + ReLU (or conv + bias), the backward is this repo's ReLU-mask + bias-gradient kernel followed by aten's
    convolution_backward for the data and weight gradients, un-padded again before they are returned."""

    @staticmethod
    def forward(ctx, x, w, b, segs, padding, relu):
        wp, bp = _pad_wb(w, b, segs, x.dtype)
        padding = list(padding)
        if relu:
            y = torch.cudnn_convolution_relu(x, wp, bp, [1, 1], padding, [1, 1], 1)
            ctx.save_for_backward(x, wp, y)
        else:
            y = F.conv2d(x, wp, bp, padding=padding)
            ctx.save_for_backward(x, wp)
        ctx.cfg = (tuple(segs), padding, bool(relu), w.shape[0], w.dtype, b.dtype)
        return y

    @staticmethod
    def backward(ctx, g):
        segs, padding, relu, co, wdt, bdt = ctx.cfg
        if relu:
            x, wp, y = ctx.saved_tensors
            # ReLU mask + bias gradient in one pass (csrc/nhwc_ops.cu); g is read in place even when it is a channel slice
            # of a concatenation's gradient (pitch = channels of the wider buffer)
            _, Cn, H, W = y.shape
            gv = g.permute(0, 2, 3, 1)
            st = gv.stride()
            sliced = (g.dtype == y.dtype and st[3] == 1 and st[2] % 8 == 0 and st[2] >= Cn and st[1] == W * st[2]
                      and g.data_ptr() % 16 == 0)
            if not sliced:
                gv = gv.to(y.dtype).contiguous()
                st = gv.stride()
            yv = _nhwc(y)
            gm = torch.empty((1, H, W, Cn), dtype=y.dtype, device=y.device)
            db = torch.zeros(Cn, dtype=torch.float32, device=y.device)
            with torch.cuda.device(y.device):
                N.check(N.lib.ibgs_nhwc_relu_bias_backward(gv.data_ptr(), st[2], yv.data_ptr(), gm.data_ptr(), db.data_ptr(),
                                                           H * W, Cn, int(y.dtype == torch.bfloat16), _stream(y.device)),
                        "ibgs_nhwc_relu_bias_backward")
            gx, gwp, _ = torch.ops.aten.convolution_backward(gm.permute(0, 3, 1, 2), x, wp, None, [1, 1], padding, [1, 1], False,
                                                             [0, 0], 1, [ctx.needs_input_grad[0], True, False])
        else:
            x, wp = ctx.saved_tensors
            gx, gwp, db = torch.ops.aten.convolution_backward(g.to(x.dtype), x, wp, [wp.shape[0]], [1, 1], padding, [1, 1], False,
                                                              [0, 0], 1, [ctx.needs_input_grad[0], True, True])
        return gx, _unpad_w(gwp, co, segs, wdt), db[:co].to(bdt), None, None, None


class _Cat2(torch.autograd.Function):
    """torch.cat([a, b], dim=1) of two (1, C, H, W) NHWC tensors in one pass at copy rate (csrc/nhwc_ops.cu, identity
    mapping of the upsample+cat kernel); the backward hands out the two channel slices as views, like torch."""

    @staticmethod
    def forward(ctx, a, b):
        av, bv = _nhwc(a), _nhwc(b)
        _, H, W, Ca = av.shape
        Cb = bv.shape[3]
        out = torch.empty((1, H, W, Ca + Cb), dtype=a.dtype, device=a.device)
        with torch.cuda.device(a.device):
            N.check(N.lib.ibgs_nhwc_upsample_cat_forward(av.data_ptr(), bv.data_ptr(), out.data_ptr(), H, W, H, W, Ca, Cb,
                                                         int(a.dtype == torch.bfloat16), _stream(a.device)),
                    "ibgs_nhwc_upsample_cat_forward")
        ctx.ca = Ca
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        return g[:, :ctx.ca], g[:, ctx.ca:]


def cat2(a, b):
    if not (_fast_glue_ok(a) and _fast_glue_ok(b) and a.dtype == b.dtype and a.shape[2:] == b.shape[2:]):
        raise RuntimeError("ibgs_b200.color_aggregation.cat2: two (1, C % 8 == 0, H, W) CUDA tensors of one dtype and size")
    return _Cat2.apply(a, b)


def conv_decoder(net, x):
    """ConvDecoderAE.forward (color_aggregation_network.py:51-68) of the unchanged module `net` on a channel-padded
    NHWC input x (1, 40, H, W); returns the (1, 3, H, W) residual in x's dtype."""
    h = net.enc1[0].in_channels
    def conv(t, seq, segs, relu=True):
        c = seq if isinstance(seq, torch.nn.Conv2d) else seq[0]
        return _ConvLayer.apply(t, c.weight, c.bias, tuple(segs), tuple(c.padding), relu)

    e1 = conv(x, net.enc1, [h])
    p1 = max_pool2(e1)
    e2 = conv(p1, net.enc2, [h])
    p2 = max_pool2(e2)
    bottleneck = conv(p2, net.enc3, [h // 2])
    u2 = upsample_nearest(bottleneck, e2.shape[-2:])
    u2 = conv(u2, net.up2_conv, [h // 4])
    d2 = conv(cat2(u2, e2), net.dec2, [h // 2, h // 2])
    u1 = upsample_nearest(d2, e1.shape[-2:])
    u1 = conv(u1, net.up1_conv, [h // 2])
    d1 = conv(cat2(u1, e1), net.dec1, [h, h])
    fused = conv(cat2(d1, x), net.fuse_input, [h, h])
    return conv(fused, net.final, [h], relu=False)[:, :3]


def compute_exposure_affine_matrix(I_s_warp, I_r, valid_mask):
    """color_aggregation_network.py:136-153: the affine colour map A (3x4) minimising |A [I_r; 1] - I_s_warp| over the
    valid pixels, applied to the whole image.  Solved from the 4x4 normal equations (float64, masked sums)."""
    with torch.no_grad():
        _, H, W = I_r.shape
        m = valid_mask[0].reshape(1, -1).to(torch.float64)
        X = torch.cat([I_r.detach().reshape(3, -1).double(), torch.ones((1, H * W), dtype=torch.float64, device=I_r.device)], 0)
        Xm = X * m
        XtX = Xm @ X.T                                          # (4, 4)
        XtY = Xm @ I_s_warp.detach().reshape(3, -1).double().T  # (4, 3)
        A = torch.linalg.lstsq(XtX, XtY).solution               # tolerates a rank-deficient system like the reference's lstsq
        affine_matrix = A.T.to(I_r.dtype)                       # (3, 4)
    transformed = torch.einsum("ij,jhw->ihw", affine_matrix[:, :3], I_r) + affine_matrix[:, 3].view(3, 1, 1)
    return transformed, affine_matrix


def fuse_color(render_pkg, color_aggregation_network, iter_count, burn_start, burn_end, iteration, opts, precision="bf16"):
    """Drop-in for color_aggregation_network.fuse_color (:156-250); same arguments, same result dict."""
    if color_aggregation_network is None:
        return None
    if precision not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    if opts.residual_resolution_scale != 1:
        raise NotImplementedError("ibgs_b200.color_aggregation.fuse_color covers residual_resolution_scale == 1")
    net = color_aggregation_network
    if net.per_view_feat_dim != 32:
        raise NotImplementedError("ibgs_b200.color_aggregation.fuse_color covers per_view_feat_dim == 32")

    if iter_count is None or burn_start is None or burn_end is None:
        burned_in_gauss = 1.0
    else:
        burned_in_gauss = max(0.0, min(1.0, (iter_count - burn_start) / (burn_end - burn_start)))
        burned_in_gauss = (burned_in_gauss + 1) / 2

    det = (lambda t: t.detach()) if burned_in_gauss < 1.0 else (lambda t: t)     # :171-177
    rendered_image = det(render_pkg["render"])
    _, H, W = rendered_image.shape
    warped = det(render_pkg["warped_image"])
    cam_feat = det(render_pkg["cam_feat"])
    min_depth_diff = det(render_pkg["min_depth_diff"])
    camera_ray_world = det(render_pkg["camera_ray"]).view(3, H, W)
    warped4 = warped.view(-1, 3, H, W)
    warped_image_list = warped4.permute(2, 3, 0, 1)

    if opts.enable_exposure_correction:
        use_first_src_mask = render_pkg["use_first_src_frame_mask"]
        first_warped_image = warped4[0] * use_first_src_mask
        rendered_image, _ = compute_exposure_affine_matrix(first_warped_image, rendered_image, use_first_src_mask == 1)

    # number of leading source views that produced any warped colour (:191-194; one host read, as in the reference)
    nb_valid_warp_level = torch.count_nonzero(warped4.sum(dim=(1, 2, 3))).item()
    nb_valid_warp_level = min(nb_valid_warp_level, opts.nb_visible_src_frames)
    if nb_valid_warp_level == 0:
        return None
    warped_image_list = warped_image_list[:, :, :nb_valid_warp_level]
    valid_warp_mask = (min_depth_diff < 0.999).float()

    x = color_features(warped, cam_feat, rendered_image, camera_ray_world, net.per_view_mlp, nb_valid_warp_level,
                       mode=net.feat_aggregate_mode, bf16=(precision == "bf16"))
    residual = conv_decoder(net.conv_decoder, x)[0].float()
    image_pred = burned_in_gauss * rendered_image + residual
    return {
        "image_pred": image_pred,
        "warped_image_list": warped_image_list,
        "residual": residual,
        "valid_warp_mask": valid_warp_mask,
        "burned_in_gauss": burned_in_gauss,
        "nb_valid_warp_level": nb_valid_warp_level,
    }
