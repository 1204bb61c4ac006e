"""Shared helpers for the parity tests: move a synthetic scene to the GPU, run THIS implementation
through its public API (ibgs_b200.diff_plane_rasterization, i.e. through the C ABI), decode its state."""
import torch

from ibgs_b200 import synthetic as S

TENSOR_KEYS = ("means3D", "scales", "rotations", "opacities", "shs", "all_map", "bg", "ref_to_src_list",
               "src_cam_pos", "src_images", "viewmatrix", "projmatrix", "campos", "normals_world")


def scene_to_device(scene, device="cuda"):
    sc = dict(scene)
    for k in TENSOR_KEYS:
        sc[k] = scene[k].to(device).contiguous()
    return sc


def make_settings(dpr, sc, render_geo=True, render_depth_only=False, buffer_length=4,
                  depth_error_threshold=0.01, debug=False, cam=None):
    H, W = sc["H"], sc["W"]
    dev = sc["means3D"].device
    cam = cam or sc
    if render_geo:
        r2s, scp, simg, sdep, nb = (sc["ref_to_src_list"], sc["src_cam_pos"], sc["src_images"],
                                    sc["src_rendered_depths"], sc["nb_src"])
    else:  # what gaussian_renderer builds when no source views are used (__init__.py:89-92,270-275)
        nb = 1
        r2s = torch.zeros((1, 16), device=dev)
        simg = torch.zeros((1, 3, H * W), device=dev)
        sdep = torch.zeros((1, 1, H * W), device=dev)
        scp = torch.zeros((1, 3), device=dev)
    return dpr.GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=sc["bg"],
        scale_modifier=1.0, viewmatrix=cam["viewmatrix"], projmatrix=cam["projmatrix"],
        ref_to_src_list=r2s, src_cam_pos=scp, src_images=simg, src_rendered_depths=sdep, nb_src_images=nb,
        buffer_length=buffer_length, depth_error_threshold=depth_error_threshold, sh_degree=sc["sh_degree"],
        campos=cam["campos"], prefiltered=False, render_geo=render_geo, render_depth_only=render_depth_only,
        debug=debug)


def render_src_depths(dpr, sc, buffer_length=4):
    """src_rendered_depths[i] = depth-only render from source pose i (gaussian_renderer.render_depth)."""
    dev = sc["means3D"].device
    out = []
    for i in range(sc["nb_src"]):
        cam = S.src_view(sc, i)
        cam = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in cam.items()}
        rs = make_settings(dpr, sc, render_geo=False, render_depth_only=True, buffer_length=buffer_length, cam=cam)
        with torch.no_grad():
            z = torch.zeros_like(sc["means3D"])
            res = dpr.GaussianRasterizer(rs)(means3D=sc["means3D"], means2D=z, means2D_abs=z,
                                             opacities=sc["opacities"], shs=sc["shs"], scales=sc["scales"],
                                             rotations=sc["rotations"], all_map=cam["all_map"])
        out.append(res[3])
    return torch.stack(out, dim=0).contiguous() if out else torch.zeros((0, 1, sc["H"], sc["W"]), device=dev)


OUT_NAMES = ("color", "radii", "normal", "depth", "cam_feat", "warped", "min_depth_diff", "camera_ray", "mask")
GRAD_NAMES = ("means3D", "means2D", "means2D_abs", "sh", "opacities", "scales", "rotations", "all_map")


def ours_forward_backward(dpr, sc, cot=None, render_geo=True, render_depth_only=False, buffer_length=4,
                          depth_error_threshold=0.01, keep_state=True):
    """Runs our implementation through its public API.  Returns (outputs, grads or None, state)."""
    rs = make_settings(dpr, sc, render_geo, render_depth_only, buffer_length, depth_error_threshold)
    leaf = {k: sc[k].detach().clone().requires_grad_(cot is not None)
            for k in ("means3D", "shs", "opacities", "scales", "rotations", "all_map")}
    m2d = torch.zeros_like(sc["means3D"], requires_grad=cot is not None)
    m2d_abs = torch.zeros_like(sc["means3D"], requires_grad=cot is not None)
    dpr.KEEP_STATE = keep_state
    use_map = render_geo or render_depth_only
    res = dpr.GaussianRasterizer(rs)(means3D=leaf["means3D"], means2D=m2d, means2D_abs=m2d_abs,
                                     opacities=leaf["opacities"], shs=leaf["shs"], scales=leaf["scales"],
                                     rotations=leaf["rotations"], all_map=leaf["all_map"] if use_map else None)
    dpr.KEEP_STATE = False
    outs = dict(zip(OUT_NAMES, res))
    state = dict(dpr.LAST_STATE) if keep_state else {}
    grads = None
    if cot is not None:
        loss = (outs["color"] * cot["color"]).sum()
        if render_geo:
            loss = loss + (outs["normal"] * cot["normal"]).sum() + (outs["depth"] * cot["depth"]).sum() + \
                (outs["warped"] * cot["warped"]).sum()
        loss.backward()
        grads = dict(means3D=leaf["means3D"].grad, means2D=m2d.grad, means2D_abs=m2d_abs.grad,
                     sh=leaf["shs"].grad, opacities=leaf["opacities"].grad, scales=leaf["scales"].grad,
                     rotations=leaf["rotations"].grad,
                     all_map=leaf["all_map"].grad if use_map else None)
    outs = {k: v.detach() for k, v in outs.items()}
    return outs, grads, state


def decode_ours(state):
    """Decodes this implementation's state buffers with ibgs_state_layout (the C ABI's own description)."""
    from ibgs_b200 import _native as N   # (lazy: bench.py's reference arm imports this module without the library)
    P, H, W, R = state["P"], state["H"], state["W"], state["num_rendered"]
    Npix = H * W
    T = ((W + 15) // 16) * ((H + 15) // 16)
    g, b, im, sc = state["geom"], state["binning"], state["image"], state["scratch"]

    def view(buf, off, count, dtype):
        nb = count * torch.empty(0, dtype=dtype).element_size()
        return buf[off:off + nb].view(dtype)

    go, _ = N.state_layout(N.IBGS_BUF_GEOM, P)
    rec = view(g, go[0], 16 * P, torch.float32).view(P, 16)
    out = dict(
        means2D=rec[:, 0:2], conic_opacity=torch.cat([rec[:, 2:5], rec[:, 5:6]], dim=1), cull_tau=rec[:, 6],
        rgb=rec[:, 8:11], plane_d=rec[:, 11], plane_n=rec[:, 12:15],
        depths=view(g, go[1], P, torch.float32), tiles_touched=view(g, go[2], P, torch.int32),
        clamped=view(g, go[3], P, torch.uint8))
    io, _ = N.state_layout(N.IBGS_BUF_IMAGE, Npix, T)
    out.update(final_T=view(im, io[0], Npix, torch.float32), n_contrib=view(im, io[1], Npix, torch.int32),
               sum_w=view(im, io[2], Npix, torch.float32), low=view(im, io[3], Npix, torch.int32),
               high=view(im, io[4], Npix, torch.int32),
               valid_idx=view(im, io[5], 5 * Npix, torch.int32).view(5, Npix),
               valid_w=view(im, io[6], 5 * Npix, torch.float32).view(5, Npix),
               ranges=view(im, io[7], 2 * T, torch.int32).view(T, 2))
    out["point_list"] = view(b, 0, R, torch.int32)
    so, _ = N.state_layout(N.IBGS_BUF_SCRATCH, state.get("scratch_capacity", R) or R, T)
    tile_bits = N.lib.ibgs_sort_bits(T) - 32
    if tile_bits <= 16:   # tile ids are stored as uint16 up to 65536 tiles (binning.cu)
        out["tiles_unsorted"] = view(sc, so[0], R, torch.int16).int() & 0xFFFF
        out["tiles_sorted"] = view(sc, so[1], R, torch.int16).int() & 0xFFFF
    else:
        out["tiles_unsorted"] = view(sc, so[0], R, torch.int32)
        out["tiles_sorted"] = view(sc, so[1], R, torch.int32)
    if tile_bits <= 8:
        # single-pass tile sort (tiny images): the sorted tile ids are not stored; the ranges (compared with the reference's on their
        # own) say which tile every list position belongs to
        rg = out["ranges"].long()
        counts = rg[:, 1] - rg[:, 0]
        assert int(counts.sum()) == R and bool((rg[1:, 0][counts[1:] > 0] >= rg[:-1, 0].cummax(0).values[counts[1:] > 0]).all())
        out["tiles_sorted"] = torch.repeat_interleave(torch.arange(T, device=rg.device, dtype=torch.int32), counts)
    out["point_list_unsorted"] = view(sc, so[2], R, torch.int32)
    # the reference's 64-bit sort keys (tile id << 32 | depth bits, rasterizer_impl.cu:219-223) rebuilt from
    # this implementation's state: tile id of every instance + depth of the Gaussian it points to
    dbits = out["depths"].view(torch.int32).long()
    out["keys"] = (out["tiles_sorted"].long() << 32) | dbits[out["point_list"].long()]
    out["keys_unsorted"] = (out["tiles_unsorted"].long() << 32) | dbits[out["point_list_unsorted"].long()]
    return out


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    den = b.norm().item()
    return (a - b).norm().item() / den if den > 0 else (a - b).norm().item()
