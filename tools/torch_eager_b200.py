#!/usr/bin/env python
"""The library bar: the reference ARCHITECTURE as stock torch.nn modules (cuDNN convolutions,
cuDNN bidirectional LSTM, cuBLAS linears) run eagerly on the B200 - what the reference's own `device="cuda"` path
executes (utils.py:692, blocks.py:422-447) with a current PyTorch. Forward + softmax/argmax only, random weights,
synthetic input resident on the device, CUDA-event timing. Measurement infrastructure, not part of the product.

    python tools/torch_eager_b200.py [--batch 256 --height 32 --width 512 --steps 20]
"""
from __future__ import annotations

import argparse
import json

import torch
import torch.nn as nn


def conv_bn_relu(cin, cout, k, s=1, p=0):
    return nn.Sequential(nn.Conv2d(cin, cout, k, s, p, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class Block(nn.Module):  # the reference's BlockB: conv-bn-relu, conv-bn, (+1x1 conv-bn shortcut), add, relu
    def __init__(self, cin, cout):
        super().__init__()
        self.c1 = nn.Conv2d(cin, cout, 3, 1, 1, bias=False)
        self.b1 = nn.BatchNorm2d(cout)
        self.c2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.b2 = nn.BatchNorm2d(cout)
        self.ds = None if cin == cout else nn.Sequential(nn.Conv2d(cin, cout, 1, 1, 0, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        r = x if self.ds is None else self.ds(x)
        y = torch.relu(self.b1(self.c1(x)))
        return torch.relu(self.b2(self.c2(y)) + r)


class BiLSTM(nn.Module):
    def __init__(self, i, h, o):
        super().__init__()
        self.rnn = nn.LSTM(i, h, bidirectional=True, batch_first=True)
        self.lin = nn.Linear(2 * h, o)

    def forward(self, x):
        self.rnn.flatten_parameters()
        y, _ = self.rnn(x)
        return self.lin(y)


class Recognizer(nn.Module):  # ResNet-29 [1,2,5,3] -> mean over H' -> 2 x BiLSTM -> Linear(185)
    def __init__(self, num_class=185):
        super().__init__()
        L = [conv_bn_relu(1, 32, 3, 1, 1), conv_bn_relu(32, 64, 3, 1, 1), nn.MaxPool2d(2, 2), Block(64, 128),
             conv_bn_relu(128, 128, 3, 1, 1), nn.MaxPool2d(2, 2), Block(128, 256), Block(256, 256),
             conv_bn_relu(256, 256, 3, 1, 1), nn.MaxPool2d(2, (2, 1), (0, 1)), Block(256, 512)]
        L += [Block(512, 512) for _ in range(4)] + [conv_bn_relu(512, 512, 3, 1, 1)] + [Block(512, 512) for _ in range(3)]
        L += [conv_bn_relu(512, 512, 2, (2, 1), (0, 1)), conv_bn_relu(512, 512, 2, 1, 0)]
        self.cnn = nn.Sequential(*L)
        self.seq = nn.Sequential(BiLSTM(512, 512, 512), BiLSTM(512, 512, 512))
        self.pred = nn.Linear(512, num_class)

    def forward(self, x):
        f = self.cnn(x)                                # [B,512,H',T]
        v = f.permute(0, 3, 1, 2).mean(dim=3)          # [B,T,512]
        return self.pred(self.seq(v))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--height", type=int, default=32)
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = Recognizer().to(dev).eval()
    x = torch.rand(args.batch, 1, args.height, args.width, device=dev) * 2 - 1
    out = {"workload": f"{args.batch} lines of {args.height}x{args.width}, forward + softmax + argmax, torch {torch.__version__} eager",
           "params": sum(p.numel() for p in model.parameters())}

    def timed(tag, fn):
        with torch.no_grad():
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out[tag] = {"ms_per_step": ms, "lines_per_s": args.batch / (ms * 1e-3)}

    def fwd(m, xx):
        return torch.softmax(m(xx).float(), dim=2).argmax(2)

    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    timed("fp32", lambda: fwd(model, x))
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    timed("tf32", lambda: fwd(model, x))
    m16 = Recognizer().to(dev).eval().to(memory_format=torch.channels_last)
    x16 = x.contiguous(memory_format=torch.channels_last)

    def fwd16():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return fwd(m16, x16)

    timed("bf16_autocast_channels_last", fwd16)
    mh = Recognizer().to(dev).eval().half().to(memory_format=torch.channels_last)
    xh = x.half().contiguous(memory_format=torch.channels_last)
    timed("fp16_weights_channels_last", lambda: fwd(mh, xh))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
