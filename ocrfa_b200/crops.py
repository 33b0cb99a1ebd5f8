"""Host side of the GPU crop + resize that feeds the recognizer (include/ocrfa_b200.h: ocrfa_crop_resize_u8).

Mirrors the reference's `get_image_list` / `compute_ratio_and_resize` / `calculate_ratio`
(/root/reference/src/ocrfa/utils.py:1351-1418) for axis-aligned boxes (`horizontal_list`): same arguments, same
return value `(image_list, max_width)`, crops bit-identical to the reference's
`cv2.resize(..., interpolation=Image.Resampling.LANCZOS)` - the PIL enum's value is 1, i.e. OpenCV INTER_LINEAR.
The pixel arithmetic runs on the GPU; this module restates OpenCV's per-axis coefficient set-up (imgproc
resize.cpp: float32 `fx = (dx + 0.5) * scale - 0.5`, border clamps, `saturate_cast<short>(coef * 2048)`), which
depends only on (source size, destination size) and is cached.

Rotated boxes (`free_list`, `four_point_transform` = cv2.warpPerspective) are not ported: a non-empty `free_list`
raises NotImplementedError rather than silently taking another code path.
"""
from __future__ import annotations

import ctypes as C
import math
from functools import lru_cache
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib

CROP_DESC = np.dtype([("out_off", "<i8"), ("x0", "<i4"), ("y0", "<i4"), ("src_w", "<i4"), ("src_h", "<i4"),
                      ("dst_w", "<i4"), ("dst_h", "<i4"), ("xtab_off", "<i4"), ("ytab_off", "<i4"), ("mode", "<i4"),
                      ("reserved", "<i4")])
assert CROP_DESC.itemsize == 48
MODE_LINEAR, MODE_COPY, MODE_AREA2 = 0, 1, 2
_COEF_SCALE = np.float32(2048.0)  # INTER_RESIZE_COEF_SCALE = 1 << 11


def calculate_ratio(width, height):
    """utils.py:1351-1358."""
    ratio = width / height
    if ratio < 1.0:
        ratio = 1. / ratio
    return ratio


def resized_shape(width: int, height: int, model_height: int) -> Tuple[int, int, float]:
    """(dst_w, dst_h, ratio) of compute_ratio_and_resize (utils.py:1361-1372)."""
    ratio = width / height
    if ratio < 1.0:
        ratio = calculate_ratio(width, height)
        return model_height, int(model_height * ratio), ratio
    return int(model_height * ratio), model_height, ratio


@lru_cache(maxsize=8192)
def linear_table(src: int, dst: int, clamp: bool) -> np.ndarray:
    """int32 [dst, 2]: (source index, a0 | a1 << 16) for one axis of OpenCV's 8-bit INTER_LINEAR resize.

    `clamp` is the horizontal variant (index and weight clamped at the borders when the table is built); the
    vertical variant keeps the raw index (rows are clamped when they are fetched)."""
    scale = 1.0 / (float(dst) / src)  # resize(): scale_x = 1. / inv_scale_x, doubles
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp:
        lo = s < 0
        f[lo] = 0.0
        s[lo] = 0
        hi = s >= src - 1
        f[hi] = 0.0
        s[hi] = src - 1
    a0 = np.rint((np.float32(1.0) - f) * _COEF_SCALE).astype(np.int64)  # saturate_cast<short>: round half to even
    a1 = np.rint(f * _COEF_SCALE).astype(np.int64)
    packed = (a0 & 0xffff) | ((a1 & 0xffff) << 16)
    packed = np.where(packed >= 1 << 31, packed - (1 << 32), packed)
    return np.stack([s, packed], axis=1).astype(np.int32)


def crop_resize_device(page: torch.Tensor, boxes: Sequence[Tuple[int, int, int, int]], model_height: int):
    """boxes: clipped (x_min, x_max, y_min, y_max) with non-empty area. Returns (uint8 CUDA buffer with the resized
    crops packed back to back, list of (offset, dst_h, dst_w))."""
    lib = _lib.load()
    if page.dim() != 2 or page.dtype != torch.uint8 or not page.is_cuda:
        raise ValueError("page must be a 2-D uint8 CUDA tensor")
    n = len(boxes)
    desc = np.zeros(n, dtype=CROP_DESC)
    tables: List[np.ndarray] = []
    t_off = 0
    cache = {}
    layout = []
    off = 0
    for i, (x0, x1, y0, y1) in enumerate(boxes):
        w, h = x1 - x0, y1 - y0
        dw, dh, _ = resized_shape(w, h, model_height)
        d = desc[i]
        d["out_off"] = off
        d["x0"], d["y0"], d["src_w"], d["src_h"], d["dst_w"], d["dst_h"] = x0, y0, w, h, dw, dh
        if (dw, dh) == (w, h):
            d["mode"] = MODE_COPY
        elif w == 2 * dw and h == 2 * dh:
            d["mode"] = MODE_AREA2  # resize(): INTER_LINEAR with an exact 2x shrink in both axes is run as INTER_AREA
        else:
            d["mode"] = MODE_LINEAR
            for key, src, dst, clamp, field in (("x", w, dw, True, "xtab_off"), ("y", h, dh, False, "ytab_off")):
                k = (key, src, dst)
                if k not in cache:
                    t = linear_table(src, dst, clamp)
                    cache[k] = t_off
                    tables.append(t.reshape(-1))
                    t_off += t.size
                d[field] = cache[k]
        layout.append((off, dh, dw))
        off += dw * dh
        off = (off + 3) & ~3  # keep every crop 4-byte aligned (32-bit stores)
    tab = np.concatenate(tables) if tables else np.zeros(2, dtype=np.int32)
    dev = page.device
    with torch.cuda.device(dev):
        d_desc = torch.from_numpy(desc.view(np.uint8).reshape(-1)).to(dev)
        d_tab = torch.from_numpy(tab).to(dev)
        out = torch.empty(max(4, off), dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        if n:
            _lib.check(lib.ocrfa_crop_resize_u8(C.c_void_p(page.data_ptr()), page.shape[0], page.shape[1], page.stride(0),
                                                C.c_void_p(d_desc.data_ptr()), C.c_void_p(d_tab.data_ptr()), n,
                                                C.c_void_p(out.data_ptr()), C.c_void_p(stream)))
    return out, layout


def get_image_list(horizontal_list, free_list, img, model_height=64, sort_output=True, device="cuda"):
    """Drop-in for the reference's get_image_list (utils.py:1375-1418): `img` is the uint8 grayscale page (numpy
    [H, W] or a CUDA tensor); returns ([(box, crop uint8 ndarray)], max_width). Crops are computed on the GPU."""
    if len(free_list) > 0:
        raise NotImplementedError("rotated boxes (free_list / four_point_transform) are not ported to the GPU path")
    if isinstance(img, np.ndarray):
        if img.ndim != 2 or img.dtype != np.uint8:
            raise ValueError("img must be a uint8 grayscale page [H, W]")
        dev = torch.device(device if device != "cuda" else "cuda:0")
        if dev.type != "cuda":
            raise _lib.OcrfaError("ocrfa_b200 has no CPU path: device must be a CUDA device")
        page = torch.from_numpy(np.ascontiguousarray(img)).to(dev)
    else:
        page = img
    maximum_y, maximum_x = page.shape
    max_ratio_hori, max_ratio_free = 1, 1
    max_ratio_free = math.ceil(max_ratio_free)
    kept, boxes = [], []
    for box in horizontal_list:
        x_min = max(0, box[0])
        x_max = min(box[1], maximum_x)
        y_min = max(0, box[2])
        y_max = min(box[3], maximum_y)
        width = x_max - x_min
        height = y_max - y_min
        ratio = calculate_ratio(width, height)
        new_width = int(model_height * ratio)
        if new_width == 0:
            continue
        _, _, ratio = resized_shape(width, height, model_height)
        kept.append([[x_min, y_min], [x_max, y_min], [x_max, y_max], [x_min, y_max]])
        boxes.append((int(x_min), int(x_max), int(y_min), int(y_max)))
        max_ratio_hori = max(ratio, max_ratio_hori)
    image_list = []
    if boxes:
        out, layout = crop_resize_device(page, boxes, model_height)
        host = out.cpu().numpy()
        for coords, (off, dh, dw) in zip(kept, layout):
            image_list.append((coords, host[off:off + dh * dw].reshape(dh, dw).copy()))
    max_ratio_hori = math.ceil(max_ratio_hori)
    max_ratio = max(max_ratio_hori, max_ratio_free)
    max_width = math.ceil(max_ratio) * model_height
    if sort_output:
        image_list = sorted(image_list, key=lambda item: item[0][0][1])  # sort by vertical position
    return image_list, max_width
