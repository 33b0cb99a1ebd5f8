"""Host side of the GPU crop + resize that feeds the recognizer (include/ocrfa_b200.h: ocrfa_crop_resize_u8).

Mirrors the reference's `get_image_list` / `compute_ratio_and_resize` / `calculate_ratio`
(/root/reference/src/ocrfa/utils.py:1218-1231, 1351-1418): same arguments, same
return value `(image_list, max_width)`, crops bit-identical to the reference's
`cv2.resize(..., interpolation=Image.Resampling.LANCZOS)` - the PIL enum's value is 1, i.e. OpenCV INTER_LINEAR.
The pixel arithmetic runs on the GPU; this module restates OpenCV's per-axis coefficient set-up (imgproc
resize.cpp: float32 `fx = (dx + 0.5) * scale - 0.5`, border clamps, `saturate_cast<short>(coef * 2048)`), which
depends only on (source size, destination size) and is cached.

Rotated boxes (`free_list`, `four_point_transform` = cv2.getPerspectiveTransform + cv2.warpPerspective) take the
perspective-warp kernel: the 3x3 matrices are parameters, computed here with OpenCV's operation order (float32
cross terms, float64 LU solve, closed-form 3x3 inverse); the per-pixel fixed-point remap runs on the GPU.
"""
from __future__ import annotations

import ctypes as C
import math
from functools import lru_cache
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib

CROP_DESC = np.dtype([("out_off", "<i8"), ("src_off", "<i8"), ("src_stride", "<i4"), ("src_w", "<i4"), ("src_h", "<i4"),
                      ("dst_w", "<i4"), ("dst_h", "<i4"), ("xtab_off", "<i4"), ("ytab_off", "<i4"), ("mode", "<i4")])
assert CROP_DESC.itemsize == 48
WARP_DESC = np.dtype([("out_off", "<i8"), ("dst_w", "<i4"), ("dst_h", "<i4"), ("block_w", "<i4"), ("reserved", "<i4"),
                      ("m", "<f8", (9,))])
assert WARP_DESC.itemsize == 96
MODE_LINEAR, MODE_COPY, MODE_AREA2 = 0, 1, 2
_COEF_SCALE = np.float32(2048.0)  # INTER_RESIZE_COEF_SCALE = 1 << 11


class DeviceLine:
    """A resized crop that stays on the GPU: a [h, w] uint8 window at byte offset `off` of a packed CUDA buffer
    (the layout ocrfa_align_collate_u8 reads). `get_image_list(..., as_device=True)` returns these instead of numpy
    arrays so that crop -> AlignCollate -> recognition never round-trips through the host."""

    __slots__ = ("buffer", "off", "shape")
    ndim = 2
    dtype = np.dtype(np.uint8)

    def __init__(self, buffer: torch.Tensor, off: int, h: int, w: int):
        self.buffer, self.off, self.shape = buffer, int(off), (int(h), int(w))

    def numpy(self) -> np.ndarray:
        h, w = self.shape
        return self.buffer[self.off:self.off + h * w].cpu().numpy().reshape(h, w)

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)


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


# ------------------------------------------------------------------------------------------------------------------
# rotated boxes: four_point_transform (utils.py:1218-1231). The 3x3 matrices are parameters (9 doubles per box), so
# they are computed here exactly as OpenCV computes them; the per-pixel warp runs on the GPU.
# ------------------------------------------------------------------------------------------------------------------
def _lu_solve8(a: np.ndarray, b: np.ndarray):
    """OpenCV's LUImpl<double> (core/matrix_decomp.cpp) on an 8x8 system: partial pivoting, `d = -1/a[i][i]`,
    row updates `a[j][k] += alpha * a[i][k]`, then back substitution - same operation order, float64."""
    a = a.copy()
    b = b.copy()
    m = 8
    eps = np.finfo(np.float64).eps * 100
    for i in range(m):
        k = i
        for j in range(i + 1, m):
            if abs(a[j, i]) > abs(a[k, i]):
                k = j
        if abs(a[k, i]) < eps:
            return None
        if k != i:
            a[[i, k], i:] = a[[k, i], i:]
            b[[i, k]] = b[[k, i]]
        d = -1.0 / a[i, i]
        for j in range(i + 1, m):
            alpha = a[j, i] * d
            a[j, i + 1:] += alpha * a[i, i + 1:]  # element-wise, the same products and sums as the scalar loop
            b[j] += alpha * b[i]
    for i in range(m - 1, -1, -1):
        sacc = b[i]
        for kk in range(i + 1, m):
            sacc -= a[i, kk] * b[kk]
        b[i] = sacc / a[i, i]
    return b


def get_perspective_transform(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """cv2.getPerspectiveTransform(src, dst) for float32 [4, 2] points (imgproc/imgwarp.cpp: the cross terms are
    float32 products, the 8x8 system is solved in float64 by LU)."""
    src = np.asarray(src, dtype=np.float32)
    dst = np.asarray(dst, dtype=np.float32)
    a = np.zeros((8, 8), np.float64)
    b = np.zeros(8, np.float64)
    for i in range(4):
        a[i, 0] = a[i + 4, 3] = src[i, 0]
        a[i, 1] = a[i + 4, 4] = src[i, 1]
        a[i, 2] = a[i + 4, 5] = 1
        a[i, 6] = np.float32(-src[i, 0] * dst[i, 0])
        a[i, 7] = np.float32(-src[i, 1] * dst[i, 0])
        a[i + 4, 6] = np.float32(-src[i, 0] * dst[i, 1])
        a[i + 4, 7] = np.float32(-src[i, 1] * dst[i, 1])
        b[i] = dst[i, 0]
        b[i + 4] = dst[i, 1]
    x = _lu_solve8(a, b)
    if x is None:
        x = np.zeros(8)  # cv::solve leaves zeros for a singular system
    return np.append(x, 1.0).reshape(3, 3)


def invert3(S: np.ndarray) -> np.ndarray:
    """cv::invert of a 3x3 float64 matrix (closed form, core/lapack.cpp), as warpPerspective applies it."""
    det = (S[0, 0] * (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) - S[0, 1] * (S[1, 0] * S[2, 2] - S[1, 2] * S[2, 0]) +
           S[0, 2] * (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0]))
    if det == 0.0:
        return np.zeros((3, 3))
    d = 1.0 / det
    t = np.empty(9)
    t[0] = (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) * d
    t[1] = (S[0, 2] * S[2, 1] - S[0, 1] * S[2, 2]) * d
    t[2] = (S[0, 1] * S[1, 2] - S[0, 2] * S[1, 1]) * d
    t[3] = (S[1, 2] * S[2, 0] - S[1, 0] * S[2, 2]) * d
    t[4] = (S[0, 0] * S[2, 2] - S[0, 2] * S[2, 0]) * d
    t[5] = (S[0, 2] * S[1, 0] - S[0, 0] * S[1, 2]) * d
    t[6] = (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0]) * d
    t[7] = (S[0, 1] * S[2, 0] - S[0, 0] * S[2, 1]) * d
    t[8] = (S[0, 0] * S[1, 1] - S[0, 1] * S[1, 0]) * d
    return t.reshape(3, 3)


def four_point_geometry(rect: np.ndarray):
    """(mw, mh, inverse matrix) of four_point_transform (utils.py:1218-1231) for rect = float32 [4, 2] (tl, tr, br, bl)."""
    (tl, tr, br, bl) = rect
    width_a = np.sqrt(((br[0] - bl[0]) ** 2) + ((br[1] - bl[1]) ** 2))
    width_bs = np.sqrt(((tr[0] - tl[0]) ** 2) + ((tr[1] - tl[1]) ** 2))
    mw = max(int(width_a), int(width_bs))
    height_a = np.sqrt(((tr[0] - br[0]) ** 2) + ((tr[1] - br[1]) ** 2))
    height_bs = np.sqrt(((tl[0] - bl[0]) ** 2) + ((tl[1] - bl[1]) ** 2))
    mh = max(int(height_a), int(height_bs))
    dst = np.array([[0, 0], [mw - 1, 0], [mw - 1, mh - 1], [0, mh - 1]], dtype="float32")
    return mw, mh, invert3(get_perspective_transform(rect, dst))


def warp_block_width(dst_w: int, dst_h: int) -> int:
    """Column-block width of OpenCV's WarpPerspectiveInvoker (BLOCK_SZ = 32)."""
    bh0 = min(16, dst_h)
    return min(1024 // bh0, dst_w)


def _resize_desc(d, src_off, src_stride, w, h, model_height, cache, tables, t_off):
    dw, dh, _ = resized_shape(w, h, model_height)
    d["src_off"], d["src_stride"], d["src_w"], d["src_h"], d["dst_w"], d["dst_h"] = src_off, src_stride, w, h, dw, dh
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
                cache[k] = t_off[0]
                tables.append(t.reshape(-1))
                t_off[0] += t.size
            d[field] = cache[k]
    return dw, dh


def crop_resize_device(page: torch.Tensor, boxes: Sequence[Tuple[int, int, int, int]], model_height: int,
                       quads: Sequence[Tuple[int, int, np.ndarray]] = ()):
    """boxes: clipped (x_min, x_max, y_min, y_max) with non-empty area; quads: (mw, mh, inverse matrix) of rotated
    boxes. Returns (uint8 CUDA buffer with the resized crops packed back to back - quads first, then boxes - and the
    list of (offset, dst_h, dst_w))."""
    lib = _lib.load()
    if page.dim() != 2 or page.dtype != torch.uint8 or not page.is_cuda or page.stride(1) != 1:
        raise ValueError("page must be a 2-D uint8 CUDA tensor with contiguous rows")
    dev = page.device
    nq, nb = len(quads), len(boxes)
    tables: List[np.ndarray] = []
    t_off = [0]
    cache = {}
    layout = []
    off = 0
    stream = torch.cuda.current_stream(dev).cuda_stream

    def finish(desc, n_off):
        return (n_off + 3) & ~3  # keep every crop 4-byte aligned (32-bit stores)

    with torch.cuda.device(dev):
        warped = None
        qdesc = np.zeros(nq, dtype=CROP_DESC)
        if nq:
            wdesc = np.zeros(nq, dtype=WARP_DESC)
            woff = 0
            for i, (mw, mh, minv) in enumerate(quads):
                wd = wdesc[i]
                wd["out_off"], wd["dst_w"], wd["dst_h"], wd["block_w"] = woff, mw, mh, warp_block_width(mw, mh)
                wd["m"] = np.asarray(minv, dtype=np.float64).reshape(9)
                d = qdesc[i]
                d["out_off"] = off
                dw, dh = _resize_desc(d, woff, mw, mw, mh, model_height, cache, tables, t_off)
                layout.append((off, dh, dw))
                off = finish(d, off + dw * dh)
                woff += mw * mh
            warped = torch.empty(max(4, woff), dtype=torch.uint8, device=dev)
            d_wdesc = torch.from_numpy(wdesc.view(np.uint8).reshape(-1)).to(dev)
            _lib.check(lib.ocrfa_warp_perspective_u8(C.c_void_p(page.data_ptr()), page.shape[0], page.shape[1],
                                                     page.stride(0), C.c_void_p(d_wdesc.data_ptr()), nq,
                                                     C.c_void_p(warped.data_ptr()), C.c_void_p(stream)))
        bdesc = np.zeros(nb, dtype=CROP_DESC)
        for i, (x0, x1, y0, y1) in enumerate(boxes):
            d = bdesc[i]
            d["out_off"] = off
            dw, dh = _resize_desc(d, y0 * page.stride(0) + x0, page.stride(0), x1 - x0, y1 - y0, model_height, cache,
                                  tables, t_off)
            layout.append((off, dh, dw))
            off = finish(d, off + dw * dh)
        tab = np.concatenate(tables) if tables else np.zeros(2, dtype=np.int32)
        d_tab = torch.from_numpy(tab).to(dev)
        out = torch.empty(max(4, off), dtype=torch.uint8, device=dev)
        for src, desc, n in ((warped, qdesc, nq), (page, bdesc, nb)):
            if n:
                d_desc = torch.from_numpy(desc.view(np.uint8).reshape(-1)).to(dev)
                _lib.check(lib.ocrfa_crop_resize_u8(C.c_void_p(src.data_ptr()), C.c_void_p(d_desc.data_ptr()),
                                                    C.c_void_p(d_tab.data_ptr()), n, C.c_void_p(out.data_ptr()),
                                                    C.c_void_p(stream)))
        torch.cuda.current_stream(dev).synchronize()  # the descriptor / table tensors are released on return
    return out, layout


def get_image_list(horizontal_list, free_list, img, model_height=64, sort_output=True, device="cuda", as_device=False):
    """Drop-in for the reference's get_image_list (utils.py:1375-1418): `img` is the uint8 grayscale page (numpy
    [H, W] or a CUDA tensor); returns ([(box, crop uint8 ndarray)], max_width). Rotated boxes (free_list) go through
    the perspective-warp kernel, axis-aligned ones are cropped in place; both are resized on the GPU.
    as_device=True keeps the crops on the GPU (DeviceLine items, accepted by this package's AlignCollate / get_text)."""
    if isinstance(img, np.ndarray):
        if img.ndim != 2 or img.dtype != np.uint8:
            raise ValueError("img must be a uint8 grayscale page [H, W]")
        dev = _lib.resolve_device(device)
        page = torch.from_numpy(np.ascontiguousarray(img)).to(dev)
    else:
        page = img
    maximum_y, maximum_x = page.shape
    max_ratio_hori, max_ratio_free = 1, 1
    kept, quads, boxes = [], [], []
    for box in free_list:
        rect = np.array(box, dtype="float32")
        mw, mh, minv = four_point_geometry(rect)
        if mw <= 0 or mh <= 0:
            # the reference fails inside cv2.warpPerspective / cv2.resize on an empty crop (utils.py:1218-1231, 1361)
            raise ValueError(f"degenerate free-form box {box!r}: warped size {mw} x {mh}")
        ratio = calculate_ratio(mw, mh)
        new_width = int(model_height * ratio)
        if new_width == 0:
            continue
        _, _, ratio = resized_shape(mw, mh, model_height)
        kept.append(box)  # box = [[x1,y1],[x2,y2],[x3,y3],[x4,y4]]
        quads.append((mw, mh, minv))
        max_ratio_free = max(ratio, max_ratio_free)
    max_ratio_free = math.ceil(max_ratio_free)
    for box in horizontal_list:
        x_min = max(0, box[0])
        x_max = min(box[1], maximum_x)
        y_min = max(0, box[2])
        y_max = min(box[3], maximum_y)
        width = x_max - x_min
        height = y_max - y_min
        if width <= 0 or height <= 0:
            # a box that is empty after clipping to the page: the reference raises from cv2.resize (utils.py:1398-1405)
            raise ValueError(f"box {list(box)!r} is empty after clipping to the {maximum_x} x {maximum_y} page")
        ratio = calculate_ratio(width, height)
        new_width = int(model_height * ratio)
        if new_width == 0:
            continue
        _, _, ratio = resized_shape(width, height, model_height)
        kept.append([[x_min, y_min], [x_max, y_min], [x_max, y_max], [x_min, y_max]])
        boxes.append((int(x_min), int(x_max), int(y_min), int(y_max)))
        max_ratio_hori = max(ratio, max_ratio_hori)
    image_list = []
    if kept:
        out, layout = crop_resize_device(page, boxes, model_height, quads)
        if as_device:
            for coords, (off, dh, dw) in zip(kept, layout):
                image_list.append((coords, DeviceLine(out, off, dh, dw)))
        else:
            host = out.cpu().numpy()
            for coords, (off, dh, dw) in zip(kept, layout):
                image_list.append((coords, host[off:off + dh * dw].reshape(dh, dw).copy()))
    max_ratio_hori = math.ceil(max_ratio_hori)
    max_ratio = max(max_ratio_hori, max_ratio_free)
    max_width = math.ceil(max_ratio) * model_height
    if sort_output:
        image_list = sorted(image_list, key=lambda item: item[0][0][1])  # sort by vertical position
    return image_list, max_width
