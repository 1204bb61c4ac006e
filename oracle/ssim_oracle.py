"""TEST INFRASTRUCTURE ONLY -- float64 numpy restatement of the reference's SSIM (utils/loss_utils.py:24-117) and of
its gradient, used as the checker for csrc/ssim.cu.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg may import this; the product path (ibgs_b200.loss_utils) never does.

Parity: PINNED on tests/golden/ssim_ref.npz, produced by importing the reference's own utils/loss_utils.py from
/root/reference and differentiating it with torch autograd in float64 on the CPU (tests/golden/make_ssim_golden.py).
"""
import numpy as np
from scipy.ndimage import correlate

C1, C2 = 0.01 ** 2, 0.03 ** 2


def window(window_size=11, sigma=1.5):
    """loss_utils.py:24-32, with torch's own float32 arithmetic (sum order, mm) so that the window is the reference's
    bit for bit; everything after it is float64 numpy."""
    import torch
    g = torch.Tensor([np.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    g = (g / g.sum()).unsqueeze(1)
    return g.mm(g.t()).float().numpy().astype(np.float64)


def _conv(planes, w):
    """F.conv2d(x, window, padding=5, groups=channel): depthwise, zero padding (loss_utils.py:47-55)."""
    out = np.empty_like(planes)
    for i in range(planes.shape[0]):
        out[i] = correlate(planes[i], w, mode="constant", cval=0.0)
    return out


def ssim_map(img1, img2, with_partials=False):
    """loss_utils.py:46-60 on [...,H,W] arrays; returns the map (and the partials of the map with respect to the five
    convolution outputs mu1, mu2, e11, e22, e12 when asked)."""
    shape = img1.shape
    x = np.asarray(img1, np.float64).reshape(-1, shape[-2], shape[-1])
    y = np.asarray(img2, np.float64).reshape(-1, shape[-2], shape[-1])
    w = window()
    mu1, mu2 = _conv(x, w), _conv(y, w)
    s1 = _conv(x * x, w) - mu1 * mu1
    s2 = _conv(y * y, w) - mu2 * mu2
    s12 = _conv(x * y, w) - mu1 * mu2
    A, B = 2 * mu1 * mu2 + C1, 2 * s12 + C2
    Cc, D = mu1 * mu1 + mu2 * mu2 + C1, s1 + s2 + C2
    m = (A * B) / (Cc * D)
    if not with_partials:
        return m.reshape(shape)
    dA, dB, dC, dD = B / (Cc * D), A / (Cc * D), -m / Cc, -m / D
    parts = dict(mu1=dA * 2 * mu2 + dB * (-2 * mu2) + dC * 2 * mu1 + dD * (-2 * mu1),
                 mu2=dA * 2 * mu1 + dB * (-2 * mu1) + dC * 2 * mu2 + dD * (-2 * mu2),
                 e11=dD, e22=dD, e12=2 * dB)
    return m.reshape(shape), parts, (x, y, w)


def ssim_map_backward(img1, img2, g):
    """(dL/dimg1, dL/dimg2) for a cotangent g of the map: the transpose of a zero-padded correlation with a symmetric
    window is the same correlation."""
    shape = img1.shape
    _, p, (x, y, w) = ssim_map(img1, img2, with_partials=True)
    g = np.broadcast_to(np.asarray(g, np.float64), shape).reshape(x.shape)
    d1 = _conv(g * p["mu1"], w) + 2 * x * _conv(g * p["e11"], w) + y * _conv(g * p["e12"], w)
    d2 = _conv(g * p["mu2"], w) + 2 * y * _conv(g * p["e22"], w) + x * _conv(g * p["e12"], w)
    return d1.reshape(shape), d2.reshape(shape)


def ssim(img1, img2, size_average=True):
    m = ssim_map(img1, img2)
    return m.mean() if size_average else m.mean(1).mean(1).mean(1)
