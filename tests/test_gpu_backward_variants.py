"""GPU: both variants of the backward tile renderer (one / two pixels per lane, csrc/render_backward.cu) against the
float64 CPU oracle and against each other.  ibgs_backward chooses per view by the average tile-list length; the small
test scenes would only ever take the one-pixel path, so the variant is forced here (ibgs_set_backward_variant)."""
import numpy as np
import pytest
import torch

from ibgs_b200 import synthetic as S
import ibgs_testutil as U

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dpr():
    import ibgs_b200.diff_plane_rasterization as d
    return d


@pytest.fixture()
def variant():
    from ibgs_b200 import _native as N

    def force(v):
        N.check(N.lib.ibgs_set_backward_variant(v), "ibgs_set_backward_variant")
    yield force
    N.lib.ibgs_set_backward_variant(0)


def _trimmed_rel(a, b):
    a2, b2 = a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1)
    err = np.abs(a2 - b2).sum(1)
    keep = np.argsort(err)[: len(err) - max(1, len(err) // 100)]
    return np.linalg.norm(a2[keep] - b2[keep]) / max(np.linalg.norm(b2[keep]), 1e-30)


@pytest.mark.parametrize("geo", [True, False])
def test_both_variants_match_oracle_and_each_other(dpr, variant, geo):
    from oracle import oracle as O
    sc_cpu = S.make_scene("cfg1")
    sc = U.scene_to_device(sc_cpu)
    sc["src_rendered_depths"] = U.render_src_depths(dpr, sc)
    sc_cpu["src_rendered_depths"] = sc["src_rendered_depths"].cpu()
    cot_cpu = S.cotangents(sc_cpu)
    cot = {k: v.cuda() for k, v in cot_cpu.items()}
    fw = O.forward(sc_cpu, render_geo=geo)
    gr = O.backward(sc_cpu, fw, cot_cpu, render_geo=geo)
    got = {}
    for v in (1, 2):
        variant(v)
        _, grads, _ = U.ours_forward_backward(dpr, sc, cot, render_geo=geo)
        got[v] = {k: g.cpu().numpy().astype(np.float64) for k, g in grads.items() if g is not None}
        for k in U.GRAD_NAMES:
            if k not in got[v]:
                continue
            rel = _trimmed_rel(got[v][k], gr[k].reshape(got[v][k].shape))
            assert rel < 5e-3, f"variant {v}, {k}: {rel}"
    for k in got[1]:
        den = max(np.linalg.norm(got[1][k]), 1e-30)
        assert np.linalg.norm(got[1][k] - got[2][k]) / den < 1e-4, k   # only the summation order differs


def test_variant_knob_rejects_bad_values(variant):
    from ibgs_b200 import _native as N
    assert N.lib.ibgs_set_backward_variant(3) < 0 and "pixels_per_lane" in N.last_error()
    assert N.lib.ibgs_set_backward_variant(0) == 0


def test_ragged_image_and_buffer_lengths_with_two_pixels_per_lane(dpr, variant):
    """Image size not a multiple of the tile, buffer lengths 1..8 (MAXE 5 and 9 kernels), 5 source views."""
    variant(1)
    base = {}
    scs = {}
    for bl in (1, 3, 8):
        sc = U.scene_to_device(S.make_scene("tiny", W=83, H=45, nb_src=5 if bl == 8 else 4))
        sc["src_rendered_depths"] = U.render_src_depths(dpr, sc, buffer_length=bl)
        cot = {k: v.cuda() for k, v in S.cotangents(sc).items()}
        scs[bl] = (sc, cot)
        _, g, _ = U.ours_forward_backward(dpr, sc, cot, buffer_length=bl, depth_error_threshold=0.05)
        base[bl] = g
    variant(2)
    for bl, (sc, cot) in scs.items():
        _, g, _ = U.ours_forward_backward(dpr, sc, cot, buffer_length=bl, depth_error_threshold=0.05)
        for k in U.GRAD_NAMES:
            assert U.rel_l2(g[k], base[bl][k]) < 1e-4, (bl, k)


@pytest.fixture()
def fwd_variant():
    from ibgs_b200 import _native as N

    def force(v):
        N.check(N.lib.ibgs_set_forward_variant(v), "ibgs_set_forward_variant")
    yield force
    N.lib.ibgs_set_forward_variant(0)


@pytest.mark.parametrize("name,kw", [("cfg1", {}), ("tiny", dict(W=83, H=45))])
def test_forward_variants_agree(dpr, fwd_variant, name, kw):
    """One vs two pixels per lane in the forward tile renderer (by default only dense depth-only launches take the
    second one): per-pixel blend order is the same, so integer outputs are identical and float outputs agree to
    rounding (the two kernels contract a few shared products into FMAs differently) -- render_geo (buffer lengths
    1..8), colour-only and depth-only, ragged image sizes included."""
    sc = U.scene_to_device(S.make_scene(name, **kw))
    fwd_variant(1)
    sc["src_rendered_depths"] = U.render_src_depths(dpr, sc)
    modes = [dict(render_geo=True, buffer_length=bl) for bl in (1, 2, 3, 4, 5, 8)]
    modes += [dict(render_geo=False)] + [dict(render_geo=False, render_depth_only=True, buffer_length=bl) for bl in (1, 2, 3, 4, 7)]
    ref = []
    for m in modes:
        outs, _, state = U.ours_forward_backward(dpr, sc, None, **m)
        st = U.decode_ours(state)
        ref.append((outs, st["final_T"].clone(), st["n_contrib"].clone()))
    fwd_variant(2)
    for m, (o1, T1, n1) in zip(modes, ref):
        outs, _, state = U.ours_forward_backward(dpr, sc, None, **m)
        st = U.decode_ours(state)
        assert torch.equal(outs["radii"], o1["radii"])
        assert (st["n_contrib"] != n1).float().mean().item() < 1e-4, m       # a pair exactly on a threshold may flip
        assert (st["final_T"] - T1).abs().max().item() < 1e-5, m
        for k in U.OUT_NAMES:
            if k in ("radii", "mask"):
                continue
            d = (outs[k] - o1[k]).abs()
            assert (d > 1e-4).float().mean().item() < 1e-3, (m, k, d.max().item())
        assert (outs["mask"] != o1["mask"]).float().mean().item() < 1e-3, m
