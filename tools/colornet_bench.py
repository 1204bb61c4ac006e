"""ColorFusionResidualNet's conv decoder (color_aggregation_network.py:6-68) forward+backward under different memory
formats / dtypes (dev tool; decides what ibgs_b200.color_aggregation runs).   usage: python tools/colornet_bench.py [H W]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref", "py"))
import color_aggregation_network as CAN  # noqa: E402

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (822, 1237)
torch.manual_seed(0)
net = CAN.ConvDecoderAE(38).cuda()
x0 = torch.randn(1, 38, H, W, device="cuda")


def run(name, fmt, dtype, iters=10):
    n = net.to(memory_format=fmt)
    x = x0.to(memory_format=fmt).clone().requires_grad_(True)

    def step():
        with torch.autocast("cuda", dtype=dtype, enabled=dtype is not None):
            y = n(x)
        y.float().square().mean().backward()
        return y
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        y = step()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) / iters:8.3f} ms fwd+bwd", flush=True)
    return y.detach().float()


torch.backends.cudnn.benchmark = False
ref = run("fp32 NCHW (reference)", torch.contiguous_format, None)
torch.backends.cudnn.benchmark = True
run("fp32 NCHW cudnn.benchmark", torch.contiguous_format, None)
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
y = run("fp32 NHWC tf32", torch.channels_last, None)
print("   max-abs vs fp32:", (y - ref).abs().max().item())
y = run("bf16 autocast NCHW", torch.contiguous_format, torch.bfloat16)
print("   max-abs vs fp32:", (y - ref).abs().max().item(), "ref max", ref.abs().max().item())
y = run("bf16 autocast NHWC", torch.channels_last, torch.bfloat16)
print("   max-abs vs fp32:", (y - ref).abs().max().item())
y = run("fp16 autocast NHWC", torch.channels_last, torch.float16)
print("   max-abs vs fp32:", (y - ref).abs().max().item())


# ---- channel counts padded to multiples of 8 (zero weights): same function, tensor-core-aligned shapes ----------------
import torch.nn.functional as F  # noqa: E402


def pad8(c):
    return (c + 7) // 8 * 8


class Padded(torch.nn.Module):
    """ConvDecoderAE with every channel count rounded up to a multiple of 8; the padded weights are rebuilt from the
    original parameters in every forward (differentiable), inputs to a conv after a concat are mapped segment-wise."""
    def __init__(self, net, hidden):
        super().__init__()
        self.net, self.h = net, hidden

    def w(self, conv, in_segments):
        wt, b = conv.weight, conv.bias
        co = wt.shape[0]
        parts, at = [], 0
        for c in in_segments:
            seg = wt[:, at:at + c]
            parts.append(F.pad(seg, (0, 0, 0, 0, 0, pad8(c) - c)))
            at += c
        wt = torch.cat(parts, 1)
        wt = F.pad(wt, (0, 0, 0, 0, 0, 0, 0, pad8(co) - co))
        return wt.contiguous(memory_format=torch.channels_last), F.pad(b, (0, pad8(co) - co))

    def conv(self, x, seq, segs, relu=True, pad=1):
        wt, b = self.w(seq if isinstance(seq, torch.nn.Conv2d) else seq[0], segs)
        y = F.conv2d(x, wt.to(x.dtype), b.to(x.dtype), padding=pad)
        return F.relu(y) if relu else y

    def forward(self, x):   # x: (1, pad8(h), H, W) channels_last
        n, h = self.net, self.h
        e1 = self.conv(x, n.enc1, [h])
        p1 = F.max_pool2d(e1, 2)
        e2 = self.conv(p1, n.enc2, [h])
        p2 = F.max_pool2d(e2, 2)
        bt = self.conv(p2, n.enc3, [h // 2])
        u2 = F.interpolate(bt, size=e2.shape[-2:], mode="nearest")
        u2 = self.conv(u2, n.up2_conv, [h // 4])
        d2 = self.conv(torch.cat([u2, e2], 1), n.dec2, [h // 2, h // 2])
        u1 = F.interpolate(d2, size=e1.shape[-2:], mode="nearest")
        u1 = self.conv(u1, n.up1_conv, [h // 2])
        d1 = self.conv(torch.cat([u1, e1], 1), n.dec1, [h, h])
        fused = self.conv(torch.cat([d1, x], 1), n.fuse_input, [h, h], pad=0)
        return self.conv(fused, n.final, [h], relu=False, pad=0)[:, :3]


def run_padded(name, dtype, iters=10):
    pn = Padded(net.to(memory_format=torch.contiguous_format), 38)
    xp = F.pad(x0, (0, 0, 0, 0, 0, 2)).contiguous(memory_format=torch.channels_last).clone().requires_grad_(True)

    def step():
        with torch.autocast("cuda", dtype=dtype, enabled=dtype is not None):
            y = pn(xp.to(dtype) if dtype else xp)
        y.float().square().mean().backward()
        return y
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        y = step()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) / iters:8.3f} ms fwd+bwd", flush=True)
    return y.detach().float()


for nm, dt in (("padded NHWC tf32", None), ("padded NHWC bf16", torch.bfloat16), ("padded NHWC fp16", torch.float16)):
    y = run_padded(nm, dt)
    print("   max-abs vs fp32:", (y - ref).abs().max().item())
from torch.profiler import profile, ProfilerActivity  # noqa: E402
pn = Padded(net, 38)
xp = F.pad(x0, (0, 0, 0, 0, 0, 2)).contiguous(memory_format=torch.channels_last).clone().requires_grad_(True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = pn(xp.to(torch.bfloat16))
    y.float().square().mean().backward()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=90))
