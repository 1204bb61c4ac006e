"""CPU: the float64 SSIM oracle (oracle/ssim_oracle.py) against golden vectors produced by the reference's own
utils/loss_utils.py under torch autograd in float64 (tests/golden/make_ssim_golden.py)."""
import os

import numpy as np

from oracle import ssim_oracle as SO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ssim_ref.npz")


def test_oracle_matches_reference_python():
    g = np.load(GOLD)
    for tag in ("a", "b", "c"):
        x, y, cot = g[f"{tag}_x"], g[f"{tag}_y"], g[f"{tag}_cot"]
        m = SO.ssim_map(x, y)
        assert np.abs(m - g[f"{tag}_map"]).max() < 1e-12
        d1, d2 = SO.ssim_map_backward(x, y, cot)
        assert np.abs(d1 - g[f"{tag}_dx"]).max() < 1e-10 * max(1.0, np.abs(g[f"{tag}_dx"]).max())
        assert np.abs(d2 - g[f"{tag}_dy"]).max() < 1e-10 * max(1.0, np.abs(g[f"{tag}_dy"]).max())
        assert abs(SO.ssim(x, y) - g[f"{tag}_ssim"]) < 1e-13
        d1, d2 = SO.ssim_map_backward(x, y, 1.0 / x.size)
        assert np.abs(d1 - g[f"{tag}_ssim_dx"]).max() < 1e-12
        assert np.abs(d2 - g[f"{tag}_ssim_dy"]).max() < 1e-12
        assert np.abs(m.mean(0) - g[f"{tag}_ssim2"]).max() < 1e-12
    assert np.abs(SO.ssim(g["b_x"], g["b_y"], size_average=False) - g["b_ssim_per_image"]).max() < 1e-13


def test_oracle_properties():
    rng = np.random.default_rng(0)
    x = rng.random((2, 30, 41))
    # identical images -> SSIM 1 everywhere; symmetric in its arguments; gradient of the mean vanishes at x == y
    assert np.abs(SO.ssim_map(x, x) - 1.0).max() < 1e-12
    y = rng.random((2, 30, 41))
    assert np.abs(SO.ssim_map(x, y) - SO.ssim_map(y, x)).max() < 1e-13
    d1, d2 = SO.ssim_map_backward(x, x, 1.0)
    assert np.abs(d1 + d2).max() < 1e-9
    # finite-difference check of the analytic gradient
    cot = rng.normal(size=x.shape)
    d1, _ = SO.ssim_map_backward(x, y, cot)
    e = np.zeros_like(x)
    e[1, 7, 9] = 1e-6
    fd = ((SO.ssim_map(x + e, y) - SO.ssim_map(x - e, y)) * cot).sum() / 2e-6
    assert abs(fd - d1[1, 7, 9]) < 1e-5 * max(1.0, abs(fd))
