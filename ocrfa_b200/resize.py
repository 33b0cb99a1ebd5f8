"""Host side of the GPU AlignCollate (include/ocrfa_b200.h: ocrfa_align_collate_u8).

Builds, per line, the fixed-point coefficient tables Pillow's 8-bit BICUBIC resize uses
(Pillow src/libImaging/Resample.c: precompute_coeffs + normalize_coeffs_8bpc; called by the reference at
/root/reference/src/ocrfa/utils.py:598) in float64 with Pillow's operation order, packs the ragged uint8 lines and
descriptors, and launches the kernels. The arithmetic on pixels (histogram, percentiles, LUT, two resample passes)
runs on the GPU; the tables depend only on (input size, output size) and are cached.
"""
from __future__ import annotations

import ctypes as C
import math
from functools import lru_cache
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib

PRECISION_BITS = 32 - 8 - 2

LINE_DESC = np.dtype([("src_off", "<i8"), ("tmp_off", "<i8"), ("src_h", "<i4"), ("src_w", "<i4"), ("out_w", "<i4"),
                      ("kh_off", "<i4"), ("kh_size", "<i4"), ("bh_off", "<i4"), ("kv_off", "<i4"), ("kv_size", "<i4"),
                      ("bv_off", "<i4"), ("contrast", "<i4")])
assert LINE_DESC.itemsize == 56


@lru_cache(maxsize=4096)
def bicubic_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """(bounds int32 [out,2] = (first input index, tap count), coef int32 [out,ksize], ksize) for one axis.

    Vectorised over output positions, sequential over taps so that the running sum `ww` rounds exactly like the C
    loop; every expression keeps Pillow's association order.
    """
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    xx = np.arange(out_size, dtype=np.float64)
    center = (xx + 0.5) * scale
    xmin = np.trunc(center - support + 0.5).astype(np.int64)
    xmin = np.maximum(xmin, 0)
    xmax = np.trunc(center + support + 0.5).astype(np.int64)
    xmax = np.minimum(xmax, in_size)
    n = xmax - xmin
    k = np.zeros((out_size, ksize), dtype=np.float64)
    ww = np.zeros(out_size, dtype=np.float64)
    for x in range(ksize):
        live = x < n
        arg = ((x + xmin).astype(np.float64) - center + 0.5) * ss
        a = np.abs(arg)
        w_near = ((-0.5 + 2.0) * a - (-0.5 + 3.0)) * a * a + 1
        w_far = (((a - 5) * a + 8) * a - 4) * -0.5
        w = np.where(a < 1.0, w_near, np.where(a < 2.0, w_far, 0.0))
        w = np.where(live, w, 0.0)
        k[:, x] = w
        ww = np.where(live, ww + w, ww)
    nz = ww != 0.0
    k[nz] = k[nz] / ww[nz, None]
    v = k * float(1 << PRECISION_BITS)
    kk = np.where(k < 0, np.trunc(-0.5 + v), np.trunc(0.5 + v)).astype(np.int32)
    bounds = np.stack([xmin, n], axis=1).astype(np.int32)
    return bounds, kk, ksize


def resized_width(w: int, h: int, img_h: int, img_w: int) -> int:
    """utils.py:592-596."""
    ratio = w / float(h)
    return img_w if math.ceil(img_h * ratio) > img_w else math.ceil(img_h * ratio)


def align_collate_device(images: Sequence[np.ndarray], img_h: int, img_w: int, adjust_contrast: float,
                         device: torch.device):
    """uint8 lines of arbitrary size -> (uint8 [B, img_h, img_w] CUDA tensor, int32 widths [B] CUDA tensor)."""
    lib = _lib.load()
    B = len(images)
    desc = np.zeros(B, dtype=LINE_DESC)
    tables: List[np.ndarray] = []
    t_off = 0
    src_off = 0
    tmp_off = 0
    widths = np.empty(B, dtype=np.int32)

    def add_table(arr: np.ndarray) -> int:
        nonlocal t_off
        off = t_off
        tables.append(np.ascontiguousarray(arr, dtype=np.int32).reshape(-1))
        t_off += arr.size
        return off

    # lines that already live on the GPU (crops.DeviceLine windows of ONE packed buffer) are read in place
    on_device = len(images) > 0 and all(hasattr(img, "buffer") and hasattr(img, "off") for img in images)
    if on_device and any(img.buffer is not images[0].buffer for img in images):
        raise ValueError("device lines of one batch must share one packed buffer")
    cache = {}
    for i, img in enumerate(images):
        if img.ndim != 2 or img.dtype != np.uint8:
            raise ValueError("lines must be uint8 grayscale arrays [h, w]")
        h, w = img.shape
        if on_device:
            src_off = img.off
        rw = resized_width(w, h, img_h, img_w)
        d = desc[i]
        d["src_off"], d["tmp_off"] = src_off, tmp_off
        d["src_h"], d["src_w"], d["out_w"] = h, w, rw
        d["contrast"] = 1 if adjust_contrast > 0 else 0
        if rw != w:
            key = ("h", w, rw)
            if key not in cache:
                b, kk, ks = bicubic_coeffs(w, rw)
                cache[key] = (add_table(kk), ks, add_table(b))
            d["kh_off"], d["kh_size"], d["bh_off"] = cache[key]
        if img_h != h:
            key = ("v", h, img_h)
            if key not in cache:
                b, kk, ks = bicubic_coeffs(h, img_h)
                cache[key] = (add_table(kk), ks, add_table(b))
            d["kv_off"], d["kv_size"], d["bv_off"] = cache[key]
        widths[i] = rw
        src_off += h * w
        tmp_off += h * rw
    if not on_device:
        src = torch.empty(max(1, src_off), dtype=torch.uint8).pin_memory()
        src_np = src.numpy()
        o = 0
        for img in images:
            src_np[o:o + img.size] = img.reshape(-1)
            o += img.size
    tab = np.concatenate(tables) if tables else np.zeros(1, dtype=np.int32)
    with torch.cuda.device(device):
        d_src = images[0].buffer if on_device else src.to(device, non_blocking=True)
        d_desc = torch.from_numpy(desc.view(np.uint8).reshape(-1)).to(device)
        d_tab = torch.from_numpy(tab).to(device)
        d_tmp = torch.empty(max(1, tmp_off), dtype=torch.uint8, device=device)
        d_lut = torch.empty(B * 256, dtype=torch.uint8, device=device)
        out = torch.empty((B, img_h, img_w), dtype=torch.uint8, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        _lib.check(lib.ocrfa_align_collate_u8(C.c_void_p(d_src.data_ptr()), C.c_void_p(d_desc.data_ptr()),
                                              C.c_void_p(d_tab.data_ptr()), B, img_h, img_w, float(adjust_contrast),
                                              C.c_void_p(d_tmp.data_ptr()), C.c_void_p(d_lut.data_ptr()),
                                              C.c_void_p(out.data_ptr()), C.c_void_p(stream)))
        d_w = torch.from_numpy(widths).to(device)
    return out, d_w
