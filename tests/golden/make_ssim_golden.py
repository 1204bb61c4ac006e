"""Generates tests/golden/ssim_ref.npz by running the REFERENCE's own utils/loss_utils.py (imported from
/root/reference, CPU, float64, torch autograd).  Run in the build container:  python tests/golden/make_ssim_golden.py"""
import importlib.util
import os

import numpy as np
import torch

REF = os.environ.get("IBGS_REFERENCE_ROOT", "/root/reference")
spec = importlib.util.spec_from_file_location("ref_loss_utils", os.path.join(REF, "utils", "loss_utils.py"))
L = importlib.util.module_from_spec(spec)
spec.loader.exec_module(L)

out = {}
g = torch.Generator().manual_seed(20251017)
for tag, shape in (("a", (3, 37, 53)), ("b", (2, 3, 20, 45)), ("c", (1, 9, 12))):
    x = torch.rand(shape, generator=g, dtype=torch.float64).requires_grad_(True)
    # a correlated second image (SSIM well away from 0) with smooth and flat regions
    y = (0.7 * x.detach() + 0.3 * torch.rand(shape, generator=g, dtype=torch.float64)).requires_grad_(True)
    with torch.no_grad():
        y[..., : shape[-2] // 3, :] = 0.25
    cot = torch.randn(shape, generator=g, dtype=torch.float64)
    m = L.compute_photometric_ssim(x, y, size_average=False)
    (m * cot).sum().backward()
    out[f"{tag}_x"], out[f"{tag}_y"], out[f"{tag}_cot"] = x.detach().numpy(), y.detach().numpy(), cot.numpy()
    out[f"{tag}_map"] = m.detach().numpy()
    out[f"{tag}_dx"], out[f"{tag}_dy"] = x.grad.numpy().copy(), y.grad.numpy().copy()
    x.grad = None
    y.grad = None
    s = L.ssim(x, y)
    s.backward()
    out[f"{tag}_ssim"] = s.detach().numpy()
    out[f"{tag}_ssim_dx"], out[f"{tag}_ssim_dy"] = x.grad.numpy().copy(), y.grad.numpy().copy()
    if len(shape) == 4:
        out[f"{tag}_ssim_per_image"] = L.ssim(x, y, size_average=False).detach().numpy()
    out[f"{tag}_ssim2"] = L.ssim2(x, y).detach().numpy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ssim_ref.npz"), **out)
print({k: v.shape for k, v in out.items()})
