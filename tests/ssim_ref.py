"""Test-only float32 torch transcription of the reference's SSIM (utils/loss_utils.py:24-60), for full-size GPU
comparisons where the numpy oracle would take minutes.  Not imported by the product."""
from math import exp

import torch
import torch.nn.functional as F


def _window(channel, device, dtype):
    g = torch.Tensor([exp(-(x - 11 // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
    g = (g / g.sum()).unsqueeze(1)
    w2 = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, 11, 11).contiguous().to(device=device, dtype=dtype)


def torch_ssim_map(img1, img2):
    channel = img1.size(-3)
    w = _window(channel, img1.device, img1.dtype)
    mu1 = F.conv2d(img1, w, padding=5, groups=channel)
    mu2 = F.conv2d(img2, w, padding=5, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img1 * img1, w, padding=5, groups=channel) - mu1_sq
    s2 = F.conv2d(img2 * img2, w, padding=5, groups=channel) - mu2_sq
    s12 = F.conv2d(img1 * img2, w, padding=5, groups=channel) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return ((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
