"""Seed-fixed synthetic weights and text-line images for the OCR-FA recognizer (data generation for tests and bench).

The reference ships no weights (arabic.pth is a network download, /root/reference/src/ocrfa/utils.py:408-412)
and no sample images, so parity is exercised with a deterministic recipe that both the reference
(via oracle/gen_golden.py) and the CUDA path load. PyTorch default init is degenerate for this network
(every line decodes to the same one-character string), so the recipe below is
variance-preserving and biased towards blank so that repeats / blanks / collapses are exercised.

numpy's PCG64 streams are used (not torch RNG) so the tensors are identical across torch versions.
State-dict keys and shapes follow /root/reference/src/ocrfa/modules/blocks.py:226-338, 422-438.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Tuple

import numpy as np

# (name, cin, cout, kh, kw) of every conv in ResNet(1, 512, BlockB, [1, 2, 5, 3]) in forward order
# /root/reference/src/ocrfa/modules/blocks.py:239-322


def resnet_conv_specs():
    specs = [("conv0_1", 1, 32, 3, 3), ("conv0_2", 32, 64, 3, 3)]
    inpl = 64
    for li, (planes, blocks) in enumerate(zip([128, 256, 512, 512], [1, 2, 5, 3]), start=1):
        for b in range(blocks):
            cin = inpl if b == 0 else planes
            specs.append((f"layer{li}.{b}.conv1", cin, planes, 3, 3))
            specs.append((f"layer{li}.{b}.conv2", planes, planes, 3, 3))
            if b == 0 and inpl != planes:
                specs.append((f"layer{li}.{b}.downsample.0", inpl, planes, 1, 1))
        inpl = planes
        if li < 4:
            specs.append((f"conv{li}", planes, planes, 3, 3))
    specs.append(("conv4_1", 512, 512, 2, 2))
    specs.append(("conv4_2", 512, 512, 2, 2))
    return specs


def bn_name_for(conv_name: str) -> str:
    """conv -> its BatchNorm module name (blocks.py:247-322, 200-203, 327-331)."""
    if conv_name.endswith("downsample.0"):
        return conv_name[:-1] + "1"
    head, _, tail = conv_name.rpartition(".")
    if tail.startswith("conv"):
        return (head + "." if head else "") + "bn" + tail[4:]
    raise ValueError(conv_name)


def synth_state_dict(seed: int = 0, num_class: int = 185, blank_bias: float = 3.0, prefix: str = ""):
    """Returns an OrderedDict[str, np.ndarray(float32)] with the reference Model's state-dict keys."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = OrderedDict()
    fe = prefix + "FeatureExtraction.ConvNet."

    def normal(shape, std):
        return (rng.standard_normal(shape, dtype=np.float32) * np.float32(std)).astype(np.float32)

    def uniform(shape, lo, hi):
        return (rng.random(shape, dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)

    for name, cin, cout, kh, kw in resnet_conv_specs():
        fan_in = cin * kh * kw
        sd[fe + name + ".weight"] = normal((cout, cin, kh, kw), math.sqrt(2.0 / fan_in))
        bn = fe + bn_name_for(name)
        sd[bn + ".weight"] = uniform((cout,), 0.5, 1.0)
        sd[bn + ".bias"] = normal((cout,), 0.1)
        sd[bn + ".running_mean"] = normal((cout,), 0.1)
        sd[bn + ".running_var"] = uniform((cout,), 0.8, 1.2)
        sd[bn + ".num_batches_tracked"] = np.array(1, dtype=np.int64)
    for li in range(2):
        sm = prefix + f"SequenceModeling.{li}."
        for suffix in ("", "_reverse"):
            sd[sm + "rnn.weight_ih_l0" + suffix] = normal((2048, 512), 1.0 / math.sqrt(512))
            sd[sm + "rnn.weight_hh_l0" + suffix] = normal((2048, 512), 1.0 / math.sqrt(512))
            sd[sm + "rnn.bias_ih_l0" + suffix] = normal((2048,), 0.05)
            sd[sm + "rnn.bias_hh_l0" + suffix] = normal((2048,), 0.05)
        sd[sm + "linear.weight"] = normal((512, 1024), 2.0 / math.sqrt(1024))
        sd[sm + "linear.bias"] = normal((512,), 0.05)
    pw = normal((num_class, 512), 12.0 / math.sqrt(512))
    pb = normal((num_class,), 0.1)
    pb[0] += np.float32(blank_bias)
    sd[prefix + "Prediction.weight"] = pw
    sd[prefix + "Prediction.bias"] = pb
    return sd


def synth_lines(n: int, h: int, w: int, seed: int = 1234) -> np.ndarray:
    """uint8 [n, h, w] text-like lines: light background, dark strokes, mild noise."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.empty((n, h, w), dtype=np.uint8)
    for i in range(n):
        bg = rng.uniform(200.0, 255.0)
        img = np.full((h, w), bg, dtype=np.float32)
        for _ in range(int(rng.integers(8, 41))):
            sw = int(rng.integers(2, 13))
            sh = int(rng.integers(max(2, h // 8), h + 1))
            x0 = int(rng.integers(0, max(1, w - sw + 1)))
            y0 = int(rng.integers(0, max(1, h - sh + 1)))
            img[y0:y0 + sh, x0:x0 + sw] = rng.uniform(0.0, 80.0)
        img += rng.standard_normal((h, w), dtype=np.float32) * np.float32(6.0)
        out[i] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return out


def synth_lines_global(lo: int, hi: int, h: int, w: int, base_seed: int = 1234) -> np.ndarray:
    """Lines lo..hi-1 of an unbounded synthetic set in which line g is a function of its GLOBAL index g alone (its own
    PCG64 stream, seed base_seed + g): any sharding of the set over ranks reads the same lines, which is what lets a
    multi-GPU run be compared with a single-GPU run of the same lines."""
    out = np.empty((max(0, hi - lo), h, w), dtype=np.uint8)
    for g in range(lo, hi):
        out[g - lo] = synth_lines(1, h, w, seed=base_seed + g)[0]
    return out


def synth_score_maps(h: int, w: int, seed: int, n_words: int = 14) -> Tuple[np.ndarray, np.ndarray]:
    """CRAFT-like region / affinity maps: words are rows of Gaussian character blobs (some rotated, some touching the
    border, some tiny, some weak), affinity blobs sit between neighbouring characters."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    text = np.zeros((h, w), np.float32)
    link = np.zeros((h, w), np.float32)

    def blob(cx, cy, sx_, sy_, ang, amp):
        c, s = math.cos(ang), math.sin(ang)
        u = (xx - cx) * c + (yy - cy) * s
        v = -(xx - cx) * s + (yy - cy) * c
        return amp * np.exp(-0.5 * ((u / sx_) ** 2 + (v / sy_) ** 2)).astype(np.float32)

    for _ in range(n_words):
        n_ch = int(rng.integers(1, 9))
        cx, cy = rng.uniform(-4, w + 4), rng.uniform(-2, h + 2)
        ang = rng.choice([0.0, 0.0, rng.uniform(-0.6, 0.6), math.pi / 2])
        size = rng.uniform(1.2, 5.0)
        amp = rng.choice([1.0, 0.95, 0.8, 0.62])
        step = size * rng.uniform(1.8, 2.6)
        prev = None
        for i in range(n_ch):
            px, py = cx + i * step * math.cos(ang), cy + i * step * math.sin(ang)
            text = np.maximum(text, blob(px, py, size, size * rng.uniform(0.8, 1.4), ang, amp))
            if prev is not None:
                link = np.maximum(link, blob((px + prev[0]) / 2, (py + prev[1]) / 2, size * 0.7, size * 0.6, ang,
                                             rng.choice([0.9, 0.6, 0.3])))
            prev = (px, py)
    noise = rng.normal(0, 0.01, (h, w)).astype(np.float32)
    return np.clip(text + noise, 0, 1).astype(np.float32), np.clip(link - noise, 0, 1).astype(np.float32)
