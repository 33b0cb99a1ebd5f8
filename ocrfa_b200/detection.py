"""Host mirror of the reference's CRAFT text-detection path around the network forward.

Same names, arguments and return values as the reference (/root/reference/src/ocrfa/):
  resize geometry + detector input   utils.py:479-524 (resize_aspect_ratio, normalize_mean_variance), detectors.py:33-43
  get_det_boxes                      utils.py:106-175, 329-340
  adjust_result_coordinates          utils.py:343-349
  test_net / get_textbox             detectors.py:24-73, 95-115
  group_text_box                     utils.py:1235-1348

Where the pixels are touched the work runs on the GPU (include/ocrfa_b200.h): `ocrfa_det_preprocess_u8` builds the
network input from the uint8 page in one kernel, `CraftNet` runs the forward, `ocrfa_det_components` labels the score
maps and reduces every surviving component to its stats and dilated row extents. What stays on the host is
parameter-sized: per component one `cv2.minAreaRect` over at most 2 x (ROI rows) points - OpenCV is the library the
reference itself calls for that step, and a minimum-area rectangle depends on a pixel set only through its convex
hull, so the result is bit-identical to the reference's call on the full mask - and the box grouping (list logic over
a few hundred boxes). There is no CPU path for the pixel work: without the CUDA library these functions raise.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import numpy as np
import torch

from . import _lib
from .crops import MODE_AREA2, MODE_COPY, MODE_LINEAR, linear_table

DET_COMP = np.dtype([("label", "<i4"), ("area", "<i4"), ("left", "<i4"), ("top", "<i4"), ("width", "<i4"),
                     ("height", "<i4"), ("niter", "<i4"), ("sx", "<i4"), ("sy", "<i4"), ("ex", "<i4"), ("ey", "<i4"),
                     ("row_off", "<i4"), ("max_text", "<f4")])
assert DET_COMP.itemsize == 52
_MEAN = (0.485, 0.456, 0.406)
_VARIANCE = (0.229, 0.224, 0.225)


def _cuda_device(device) -> torch.device:
    return _lib.resolve_device(device)


# ------------------------------------------------------------------------------------------------------------------
# detector input
# ------------------------------------------------------------------------------------------------------------------
def resize_geometry(height: int, width: int, square_size, mag_ratio=1.0) -> Tuple[int, int, int, int, float]:
    """(target_h, target_w, canvas_h, canvas_w, ratio) of resize_aspect_ratio (utils.py:497-517)."""
    longest = max(height, width)
    target_size = mag_ratio * longest
    if target_size > square_size:
        target_size = square_size
    ratio = target_size / longest
    target_h, target_w = int(height * ratio), int(width * ratio)
    canvas_h = target_h + (-target_h) % 32
    canvas_w = target_w + (-target_w) % 32
    return target_h, target_w, canvas_h, canvas_w, ratio


def detector_input(image, canvas_size, mag_ratio=1.0, device="cuda") -> Tuple[torch.Tensor, float]:
    """resize_aspect_ratio(INTER_LINEAR) + normalize_mean_variance + HWC->CHW for one uint8 [H, W, 3] image (numpy or
    CUDA tensor): returns (x float32 [1, 3, H32, W32] on the GPU, target ratio). Bit-exact with the reference's x."""
    lib = _lib.load()
    if isinstance(image, np.ndarray):
        if image.ndim != 3 or image.shape[2] != 3 or image.dtype != np.uint8:
            raise ValueError("image must be uint8 [H, W, 3]")
        img = torch.from_numpy(np.ascontiguousarray(image)).to(_cuda_device(device))
    else:
        img = image
        if img.dim() != 3 or img.shape[2] != 3 or img.dtype != torch.uint8 or not img.is_cuda:
            raise ValueError("image must be a uint8 [H, W, 3] CUDA tensor")
        if img.stride(2) != 1 or img.stride(1) != 3:
            img = img.contiguous()
    h, w = int(img.shape[0]), int(img.shape[1])
    th, tw, ch, cw, ratio = resize_geometry(h, w, canvas_size, mag_ratio)
    if th <= 0 or tw <= 0:
        raise ValueError(f"image of {h}x{w} resizes to nothing at canvas_size={canvas_size}")
    dev = img.device
    if (th, tw) == (h, w):
        mode, xt, yt = MODE_COPY, None, None
    elif h == 2 * th and w == 2 * tw:
        mode, xt, yt = MODE_AREA2, None, None
    else:
        mode = MODE_LINEAR
        xt = torch.from_numpy(linear_table(w, tw, True).reshape(-1).copy()).to(dev)
        yt = torch.from_numpy(linear_table(h, th, False).reshape(-1).copy()).to(dev)
    mean255 = (C.c_float * 3)(*[np.float32(m * 255.0) for m in _MEAN])
    var255 = (C.c_float * 3)(*[np.float32(v * 255.0) for v in _VARIANCE])
    with torch.cuda.device(dev):
        x = torch.empty((1, 3, ch, cw), dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev)
        _lib.check(lib.ocrfa_det_preprocess_u8(
            C.c_void_p(img.data_ptr()), h, w, int(img.stride(0)), th, tw, ch, cw,
            C.c_void_p(xt.data_ptr() if xt is not None else 0), C.c_void_p(yt.data_ptr() if yt is not None else 0), mode,
            mean255, var255, C.c_void_p(x.data_ptr()), C.c_void_p(stream.cuda_stream)))
        if xt is not None:
            stream.synchronize()  # the table tensors are released on return
    return x, ratio


# ------------------------------------------------------------------------------------------------------------------
# score maps -> boxes
# ------------------------------------------------------------------------------------------------------------------
def det_components(textmap: torch.Tensor, linkmap: torch.Tensor, text_threshold, link_threshold, low_text,
                   want_labels: bool = False):
    """ocrfa_det_components on two float32 CUDA maps of equal shape [H, W]; they may be strided views of the CRAFT
    output `y[n]` ([H, W, 2]: textmap = y[n, :, :, 0], linkmap = y[n, :, :, 1]), which are read in place.
    Returns (n_labels, comps structured array [DET_COMP], extents int32 [rows, 2], labels int32 [H, W] CUDA or None)."""
    lib = _lib.load()
    for m in (textmap, linkmap):
        if not (isinstance(m, torch.Tensor) and m.is_cuda and m.dtype == torch.float32 and m.dim() == 2):
            raise ValueError("score maps must be 2-D float32 CUDA tensors")
    if textmap.shape != linkmap.shape or textmap.device != linkmap.device:
        raise ValueError("textmap and linkmap differ in shape or device")
    H, W = int(textmap.shape[0]), int(textmap.shape[1])
    ps = int(textmap.stride(1))
    if not (ps >= 1 and textmap.stride(0) == W * ps and linkmap.stride(1) == ps and linkmap.stride(0) == W * ps):
        textmap, linkmap, ps = textmap.contiguous(), linkmap.contiguous(), 1
    dev = textmap.device
    n = H * W
    max_comps = n // 10 + 1
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        labels = torch.empty((H, W), dtype=torch.int32, device=dev) if want_labels else None
        comps = torch.empty(max_comps * DET_COMP.itemsize, dtype=torch.uint8, device=dev)
        counts = torch.empty(4, dtype=torch.int32, device=dev)
        for max_rows in (n + 64, 3 * n + 64):  # rows <= 2.4 x pixels always; the first size fits real score maps
            ws_bytes = int(lib.ocrfa_det_workspace_bytes(H, W, max_rows))
            ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
            off = (-ws.data_ptr()) % 256
            ext = torch.empty((max_rows, 2), dtype=torch.int32, device=dev)
            _lib.check(lib.ocrfa_det_components(
                C.c_void_p(textmap.data_ptr()), C.c_void_p(linkmap.data_ptr()), ps, H, W,
                C.c_float(np.float32(low_text)), C.c_float(np.float32(link_threshold)), C.c_float(np.float32(text_threshold)),
                C.c_void_p(labels.data_ptr() if labels is not None else 0), C.c_void_p(comps.data_ptr()), max_comps,
                C.c_void_p(ext.data_ptr()), max_rows, C.c_void_p(counts.data_ptr()), C.c_void_p(ws.data_ptr() + off),
                ws_bytes, C.c_void_p(stream.cuda_stream)))
            n_labels, n_comps, n_rows, overflow = (int(v) for v in counts.cpu().tolist())
            if not overflow:
                break
        else:
            raise _lib.OcrfaError("ocrfa_det_components: extent buffer overflow")
        comps_h = comps[:n_comps * DET_COMP.itemsize].cpu().numpy().view(DET_COMP).copy()
        ext_h = ext[:n_rows].cpu().numpy().copy()
    return n_labels, comps_h, ext_h, labels


def boxes_from_components(comps: np.ndarray, extents: np.ndarray) -> Tuple[List[np.ndarray], List[int]]:
    """The tail of get_det_boxes_core (utils.py:155-175) from dilated row extents: minAreaRect -> boxPoints, the
    diamond-shape fix-up, clockwise start at the top-left corner. Returns (det, mapper)."""
    import cv2

    det, mapper = [], []
    if len(comps) == 0:
        return det, mapper
    # all components at once: the point list of a mask is, row by row, (min x, y) then (max x, y) - the order np.where
    # gives on the full mask - with empty rows dropped and single-pixel rows contributing one point
    extents = np.asarray(extents).reshape(-1, 2)
    rows = (comps["ey"] - comps["sy"]).astype(np.int64)
    row_off = comps["row_off"].astype(np.int64)
    comp_of_row = np.repeat(np.arange(len(comps)), rows)
    first_row = np.repeat(row_off, rows)
    idx = first_row + (np.arange(rows.sum()) - np.repeat(np.cumsum(rows) - rows, rows))
    lo, hi = extents[idx, 0].astype(np.int64), extents[idx, 1].astype(np.int64)
    yy = (idx - first_row) + np.repeat(comps["sy"].astype(np.int64), rows)
    live = hi >= 0
    pts_all = np.stack([np.stack([lo, yy], axis=1), np.stack([hi, yy], axis=1)], axis=1)  # [rows, 2 points, (x, y)]
    keep = np.stack([live, live & (hi != lo)], axis=1)
    per_comp = np.bincount(comp_of_row, weights=keep.sum(axis=1), minlength=len(comps)).astype(np.int64)
    flat = pts_all[keep]
    ends = np.cumsum(per_comp)
    big = np.iinfo(np.int64).max
    x_lo = np.minimum.reduceat(np.where(live, lo, big), np.cumsum(rows) - rows)
    x_hi = np.maximum.reduceat(np.where(live, hi, -1), np.cumsum(rows) - rows)
    y_lo = np.minimum.reduceat(np.where(live, yy, big), np.cumsum(rows) - rows)
    y_hi = np.maximum.reduceat(np.where(live, yy, -1), np.cumsum(rows) - rows)
    for n, c in enumerate(comps):
        pts = flat[ends[n] - per_comp[n]:ends[n]]
        box = cv2.boxPoints(cv2.minAreaRect(pts))
        side_a, side_b = np.linalg.norm(box[0] - box[1]), np.linalg.norm(box[1] - box[2])
        if abs(1 - max(side_a, side_b) / (min(side_a, side_b) + 1e-5)) <= 0.1:
            l, r, t, b = x_lo[n], x_hi[n], y_lo[n], y_hi[n]
            box = np.array([[l, t], [r, t], [r, b], [l, b]], dtype=np.float32)
        first = box.sum(axis=1).argmin()
        det.append(np.array(np.roll(box, 4 - first, 0)))
        mapper.append(int(c["label"]))
    return det, mapper


def get_det_boxes(textmap, linkmap, text_threshold, link_threshold, low_text, poly=False, estimate_num_chars=False,
                  device="cuda"):
    """utils.py:329-340: (boxes, polys, mapper). textmap / linkmap: float32 [H, W], numpy or CUDA tensors. `poly=True`
    (get_poly_core) and `estimate_num_chars` are never enabled by OCR.readtext / transform_ua's defaults and are not
    offered here."""
    if poly or estimate_num_chars:
        raise NotImplementedError("get_det_boxes: poly / estimate_num_chars are not part of the GPU detection path")
    maps = []
    for m in (textmap, linkmap):
        if isinstance(m, np.ndarray):
            m = torch.from_numpy(np.ascontiguousarray(m, dtype=np.float32)).to(_cuda_device(device))
        maps.append(m)
    _, comps, extents, _ = det_components(maps[0], maps[1], text_threshold, link_threshold, low_text)
    boxes, mapper = boxes_from_components(comps, extents)
    return boxes, [None] * len(boxes), mapper


def adjust_result_coordinates(polys, ratio_w, ratio_h, ratio_net=2):
    """utils.py:343-349: heat-map coordinates -> page coordinates (the product is formed in float64 and stored back
    into the float32 boxes, as NumPy's in-place multiply does)."""
    if len(polys) > 0:
        polys = np.array(polys)
        scale = (ratio_w * ratio_net, ratio_h * ratio_net)
        for k in range(len(polys)):
            if polys[k] is not None:
                polys[k] *= scale
    return polys


def test_net(canvas_size, mag_ratio, net, image, text_threshold, link_threshold, low_text, poly, device,
             estimate_num_chars=False):
    """detectors.py:24-73: (boxes_list, polys_list) for one image [H, W, 3] or a batch [N, H, W, 3]. `net` is a
    ocrfa_b200.CraftNet (or any callable x -> (y [N, H/2, W/2, 2] CUDA float32, feature))."""
    if poly or estimate_num_chars:
        raise NotImplementedError("test_net: poly / estimate_num_chars are not part of the GPU detection path")
    dev = _cuda_device(device)
    if (isinstance(image, np.ndarray) and image.ndim == 4) or (isinstance(image, torch.Tensor) and image.dim() == 4):
        images = list(image)
    else:
        images = [image]
    inputs = [detector_input(img, canvas_size, mag_ratio, dev) for img in images]
    target_ratio = inputs[-1][1]  # the reference keeps the last image's ratio for the whole batch (detectors.py:35-38)
    ratio_h = ratio_w = 1 / target_ratio
    x = inputs[0][0] if len(inputs) == 1 else torch.cat([t for t, _ in inputs], dim=0)
    y, _ = net(x)
    boxes_list, polys_list = [], []
    for out in y:
        boxes, polys, _ = get_det_boxes(out[:, :, 0], out[:, :, 1], text_threshold, link_threshold, low_text)
        boxes = adjust_result_coordinates(boxes, ratio_w, ratio_h)
        polys = list(boxes)  # poly=False: every polygon falls back to its box (detectors.py:66-68)
        boxes_list.append(boxes)
        polys_list.append(polys)
    return boxes_list, polys_list


def get_textbox(detector, image, canvas_size, mag_ratio, text_threshold, link_threshold, low_text, poly, device,
                optimal_num_chars=None, **kwargs):
    """detectors.py:95-115: per image a list of int32 [8] polygons (x1, y1, ..., x4, y4)."""
    if optimal_num_chars is not None:
        raise NotImplementedError("get_textbox: optimal_num_chars is not part of the GPU detection path")
    _, polys_list = test_net(canvas_size, mag_ratio, detector, image, text_threshold, link_threshold, low_text, poly,
                             device)
    return [[np.array(box).astype(np.int32).reshape((-1)) for box in polys] for polys in polys_list]


# ------------------------------------------------------------------------------------------------------------------
# box grouping (host list logic over the detector's polygons)
# ------------------------------------------------------------------------------------------------------------------
def _merge_margin(add_margin, x_lo, x_hi, y_lo, y_hi, short_side):
    margin = int(add_margin * short_side)
    return [x_lo - margin, x_hi + margin, y_lo - margin, y_hi + margin]


def group_text_box(polys, slope_ths=0.1, ycenter_ths=0.5, height_ths=0.5, width_ths=1.0, add_margin=0.05,
                   sort_output=True):
    """utils.py:1235-1348: splits the polygons into near-horizontal boxes - merged along text lines into
    [x_min, x_max, y_min, y_max] - and slanted ones, returned as margin-expanded quads. Returns (merged_list, free_list)."""
    level, free_list = [], []
    for poly in polys:
        top_slope = (poly[3] - poly[1]) / np.maximum(10, (poly[2] - poly[0]))
        bottom_slope = (poly[5] - poly[7]) / np.maximum(10, (poly[4] - poly[6]))
        if max(abs(top_slope), abs(bottom_slope)) < slope_ths:
            xs, ys = [poly[0], poly[2], poly[4], poly[6]], [poly[1], poly[3], poly[5], poly[7]]
            x_lo, x_hi, y_lo, y_hi = min(xs), max(xs), min(ys), max(ys)
            level.append([x_lo, x_hi, y_lo, y_hi, 0.5 * (y_lo + y_hi), y_hi - y_lo])
            continue
        height = np.linalg.norm([poly[6] - poly[0], poly[7] - poly[1]])
        width = np.linalg.norm([poly[2] - poly[0], poly[3] - poly[1]])
        margin = int(1.44 * add_margin * min(width, height))
        theta13 = abs(np.arctan((poly[1] - poly[5]) / np.maximum(10, (poly[0] - poly[4]))))
        theta24 = abs(np.arctan((poly[3] - poly[7]) / np.maximum(10, (poly[2] - poly[6]))))
        c13, s13 = np.cos(theta13) * margin, np.sin(theta13) * margin
        c24, s24 = np.cos(theta24) * margin, np.sin(theta24) * margin
        free_list.append([[poly[0] - c13, poly[1] - s13], [poly[2] + c24, poly[3] - s24],
                          [poly[4] + c13, poly[5] + s13], [poly[6] - c24, poly[7] + s24]])
    if sort_output:
        level = sorted(level, key=lambda item: item[4])

    # 1. text lines: consecutive boxes whose centre stays within ycenter_ths x (mean height) of the line's mean centre
    # (the reference takes np.mean of the growing lists; heights are integers and centres multiples of 0.5, so the sums
    # are exact in float64 and sum / count is the same correctly rounded quotient)
    lines, line, height_sum, centre_sum = [], [], 0.0, 0.0
    for box in level:
        if line and not abs(centre_sum / len(line) - box[4]) < ycenter_ths * (height_sum / len(line)):
            lines.append(line)
            line, height_sum, centre_sum = [], 0.0, 0.0
        line.append(box)
        height_sum += float(box[5])
        centre_sum += float(box[4])
    lines.append(line)

    # 2. inside a line, left to right: chain boxes of comparable height that sit closer than width_ths x (box height)
    merged_list = []
    for line in lines:
        if len(line) == 1:
            box = line[0]
            merged_list.append(_merge_margin(add_margin, box[0], box[1], box[2], box[3], min(box[1] - box[0], box[5])))
            continue
        chains, chain, height_sum, x_right = [], [], 0.0, None
        for box in sorted(line, key=lambda item: item[0]):
            if chain:
                mean_height = height_sum / len(chain)
                if not ((abs(mean_height - box[5]) < height_ths * mean_height)
                        and ((box[0] - x_right) < width_ths * (box[3] - box[2]))):
                    chains.append(chain)
                    chain, height_sum = [], 0.0
            chain.append(box)
            height_sum += float(box[5])
            x_right = box[1]
        if chain:
            chains.append(chain)
        for chain in chains:
            x_lo = min(chain, key=lambda b: b[0])[0]
            x_hi = max(chain, key=lambda b: b[1])[1]
            y_lo = min(chain, key=lambda b: b[2])[2]
            y_hi = max(chain, key=lambda b: b[3])[3]
            merged_list.append(_merge_margin(add_margin, x_lo, x_hi, y_lo, y_hi, min(x_hi - x_lo, y_hi - y_lo)))
    return merged_list, free_list
