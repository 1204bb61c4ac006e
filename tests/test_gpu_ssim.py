"""GPU parity of the fused SSIM (csrc/ssim.cu through ibgs_b200.loss_utils -> C ABI) against the float64 oracle, the
golden vectors of the reference's utils/loss_utils.py, and -- at the benchmark's 1080p size -- the float32 torch
transcription of the same expressions.  Gates: map <= 2e-5 max-abs, gradients <= 1e-4 relative L2 (float32 arithmetic
against float64 truth; the torch float32 path itself sits at the same distance)."""
import os

import numpy as np
import pytest
import torch

import ibgs_testutil as U
from ssim_ref import torch_ssim_map

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ssim_ref.npz")


@pytest.fixture(scope="module")
def LU():
    import ibgs_b200.loss_utils as m
    return m


def _t(a, grad=False):
    return torch.from_numpy(np.ascontiguousarray(a)).float().cuda().requires_grad_(grad)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_matches_reference_golden(LU, tag):
    g = np.load(GOLD)
    x, y, cot = _t(g[f"{tag}_x"], True), _t(g[f"{tag}_y"], True), _t(g[f"{tag}_cot"])
    m = LU.compute_photometric_ssim(x, y, size_average=False)
    assert np.abs(m.detach().cpu().numpy() - g[f"{tag}_map"]).max() <= 2e-5
    (m * cot).sum().backward()
    assert U.rel_l2(x.grad.cpu(), torch.from_numpy(g[f"{tag}_dx"])) <= 1e-4
    assert U.rel_l2(y.grad.cpu(), torch.from_numpy(g[f"{tag}_dy"])) <= 1e-4
    # only img1 / only img2 differentiable (train.py:302: the ground truth carries no gradient; :330: img2 = the warp)
    for need in ((True, False), (False, True)):
        x1, y1 = _t(g[f"{tag}_x"], need[0]), _t(g[f"{tag}_y"], need[1])
        s = LU.ssim(x1, y1)
        assert abs(s.item() - float(g[f"{tag}_ssim"])) <= 1e-5
        s.backward()
        if need[0]:
            assert U.rel_l2(x1.grad.cpu(), torch.from_numpy(g[f"{tag}_ssim_dx"])) <= 1e-4 and y1.grad is None
        else:
            assert U.rel_l2(y1.grad.cpu(), torch.from_numpy(g[f"{tag}_ssim_dy"])) <= 1e-4 and x1.grad is None
    assert np.abs(LU.ssim2(_t(g[f"{tag}_x"]), _t(g[f"{tag}_y"])).cpu().numpy() - g[f"{tag}_ssim2"]).max() <= 2e-5
    if tag == "b":
        per = LU.ssim(_t(g["b_x"]), _t(g["b_y"]), size_average=False)
        assert np.abs(per.cpu().numpy() - g["b_ssim_per_image"]).max() <= 1e-5


@pytest.mark.parametrize("shape", [(3, 64, 64), (3, 33, 95), (1, 5, 7), (2, 3, 70, 31), (3, 32, 32), (4, 1, 1)])
def test_matches_oracle_on_ragged_shapes(LU, shape):
    from oracle import ssim_oracle as SO
    rng = np.random.default_rng(sum(shape))
    x = rng.random(shape)
    y = np.clip(0.6 * x + 0.4 * rng.random(shape), 0, 1)
    cot = rng.normal(size=shape)
    xt, yt = _t(x, True), _t(y, True)
    m = LU.ssim_map(xt, yt)
    assert np.abs(m.detach().cpu().numpy() - SO.ssim_map(x, y)).max() <= 2e-5
    (m * _t(cot)).sum().backward()
    d1, d2 = SO.ssim_map_backward(x, y, cot)
    assert U.rel_l2(xt.grad.cpu(), torch.from_numpy(d1)) <= 1e-4
    assert U.rel_l2(yt.grad.cpu(), torch.from_numpy(d2)) <= 1e-4


def test_full_size_against_torch_expressions_and_properties(LU):
    g = torch.Generator().manual_seed(3)
    low = torch.rand((1, 3, 136, 241), generator=g)
    gt = torch.nn.functional.interpolate(low, size=(1080, 1920), mode="bilinear", align_corners=True)[0].cuda()
    img = (gt + 0.1 * torch.randn((3, 1080, 1920), generator=g).cuda()).clamp(0, 1)
    a = img.clone().requires_grad_(True)
    b = img.clone().requires_grad_(True)
    s_ours = LU.ssim(a, gt)
    s_ref = torch_ssim_map(b, gt).mean()
    assert abs(s_ours.item() - s_ref.item()) <= 1e-5
    s_ours.backward()
    s_ref.backward()
    assert U.rel_l2(a.grad, b.grad) <= 1e-4
    # size-independent properties: SSIM(x, x) == 1 (exactly representable inputs keep float error tiny), symmetry
    m = LU.ssim_map(gt, gt)
    assert (m - 1.0).abs().max().item() <= 1e-4
    assert (LU.ssim_map(img, gt) - LU.ssim_map(gt, img)).abs().max().item() <= 1e-5


def test_errors(LU):
    x = torch.rand((3, 16, 16))
    with pytest.raises(RuntimeError):
        LU.ssim(x, x)                       # CPU tensors: no fallback
    with pytest.raises(RuntimeError):
        LU.ssim(x.cuda(), torch.rand((3, 16, 17)).cuda())
    with pytest.raises(NotImplementedError):
        LU.ssim(x.cuda(), x.cuda(), window_size=7)
    e = torch.zeros((0, 16, 16)).cuda()
    assert LU.ssim_map(e, e).shape == (0, 16, 16)
