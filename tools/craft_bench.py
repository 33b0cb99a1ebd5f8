#!/usr/bin/env python
"""CRAFT detector forward on the B200: this repo's tensor-core path (ocrfa_craft_forward) beside the same
architecture as stock torch.nn modules (torchvision vgg16_bn features + DoubleConv blocks; cuDNN, eager) - the
library bar. Synthetic weights, input resident on the device, CUDA events. Measurement tooling, not the product.

    python tools/craft_bench.py [--height 736 --width 1280 --batch 1 --steps 20]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ocrfa_b200.craft import CRAFT_CONVS, CraftNet, synth_craft_state_dict  # noqa: E402


class DoubleConv(nn.Module):
    def __init__(self, in_ch, mid_ch, out_ch):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(in_ch + mid_ch, mid_ch, 1), nn.BatchNorm2d(mid_ch), nn.ReLU(inplace=True),
                                  nn.Conv2d(mid_ch, out_ch, 3, padding=1), nn.BatchNorm2d(out_ch), nn.ReLU(inplace=True))

    def forward(self, x):
        return self.conv(x)


class TorchCraft(nn.Module):
    """The reference architecture from stock modules (utils.py:48-98, blocks.py:27-82)."""

    def __init__(self):
        super().__init__()
        from torchvision import models
        f = models.vgg16_bn(weights=None).features
        self.s1, self.s2, self.s3, self.s4 = f[:12], f[12:19], f[19:29], f[29:39]
        self.s5 = nn.Sequential(nn.MaxPool2d(3, 1, 1), nn.Conv2d(512, 1024, 3, padding=6, dilation=6), nn.Conv2d(1024, 1024, 1))
        self.up1, self.up2, self.up3, self.up4 = DoubleConv(1024, 512, 256), DoubleConv(512, 256, 128), DoubleConv(256, 128, 64), DoubleConv(128, 64, 32)
        self.cls = nn.Sequential(nn.Conv2d(32, 32, 3, padding=1), nn.ReLU(inplace=True), nn.Conv2d(32, 32, 3, padding=1),
                                 nn.ReLU(inplace=True), nn.Conv2d(32, 16, 3, padding=1), nn.ReLU(inplace=True),
                                 nn.Conv2d(16, 16, 1), nn.ReLU(inplace=True), nn.Conv2d(16, 2, 1))

    def forward(self, x):
        a = self.s1(x); b = self.s2(a); c = self.s3(b); d = self.s4(c); e = self.s5(d)
        y = self.up1(torch.cat([e, d], 1))
        y = self.up2(torch.cat([F.interpolate(y, size=c.shape[2:], mode="bilinear", align_corners=False), c], 1))
        y = self.up3(torch.cat([F.interpolate(y, size=b.shape[2:], mode="bilinear", align_corners=False), b], 1))
        feat = self.up4(torch.cat([F.interpolate(y, size=a.shape[2:], mode="bilinear", align_corners=False), a], 1))
        return self.cls(feat).permute(0, 2, 3, 1), feat


def craft_flops(n, h, w):
    """2 * MACs of the 27 convolutions; `down` = how far below the input resolution each one runs."""
    down = [1, 1, 2, 2, 4, 4, 4, 8, 8, 8, 16, 16, 16, 16, 16, 16, 8, 8, 4, 4, 2, 2, 2, 2, 2, 2, 2]
    assert len(down) == len(CRAFT_CONVS)
    return sum(2.0 * n * (h // s) * (w // s) * cout * cin * k * k for (_, cin, cout, k, _), s in zip(CRAFT_CONVS, down))


def timed(fn, steps):
    with torch.no_grad():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=736)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    x = torch.randn(args.batch, 3, args.height, args.width, device=dev)
    flops = craft_flops(args.batch, args.height, args.width)
    out = {"workload": f"CRAFT forward, {args.batch} x 3 x {args.height} x {args.width}, synthetic weights", "gflop": flops / 1e9}
    sd = {k: torch.from_numpy(v) for k, v in synth_craft_state_dict(0).items()}
    for prec in ("bf16", "f16"):
        net = CraftNet(precision=prec)
        net.load_state_dict(sd)
        ms = timed(lambda: net(x), args.steps)
        out[f"ours_{prec}"] = {"ms": ms, "images_per_s": args.batch / (ms * 1e-3), "tflops": flops / (ms * 1e-3) / 1e12}
        del net
    torch.backends.cudnn.benchmark = True
    ref = TorchCraft().to(dev).eval()
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    ms = timed(lambda: ref(x), args.steps)
    out["torch_eager_tf32"] = {"ms": ms, "images_per_s": args.batch / (ms * 1e-3)}
    ref16 = ref.to(memory_format=torch.channels_last)
    x16 = x.contiguous(memory_format=torch.channels_last)

    def f16():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return ref16(x16)

    ms = timed(f16, args.steps)
    out["torch_eager_bf16_autocast_channels_last"] = {"ms": ms, "images_per_s": args.batch / (ms * 1e-3), "tflops": flops / (ms * 1e-3) / 1e12}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
