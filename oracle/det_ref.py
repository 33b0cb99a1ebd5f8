"""CPU restatement of the reference's detector pre- and post-processing (TEST INFRASTRUCTURE - only tests/, smoke()
and the baseline legs may import this; nothing under ocrfa_b200/ does).

Follows, with the same third-party calls the reference makes (OpenCV, NumPy - both present in this image):
  resize_aspect_ratio          /root/reference/src/ocrfa/utils.py:497-524
  normalize_mean_variance      /root/reference/src/ocrfa/utils.py:479-484
  test_net (input assembly)    /root/reference/src/ocrfa/detectors.py:24-43
  get_det_boxes_core           /root/reference/src/ocrfa/utils.py:106-175
  adjust_result_coordinates    /root/reference/src/ocrfa/utils.py:343-349
  get_textbox                  /root/reference/src/ocrfa/detectors.py:95-115
Pinned by tests/golden/detection.npz, which oracle/gen_golden_det.py produced by running the reference's own
functions. `components_ref` additionally restates what the CUDA path hands to the host (component stats + dilated row
extents), computed the slow way (full masks + cv2.dilate), so the kernels can be compared field by field.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import cv2
import numpy as np

MEAN = (0.485, 0.456, 0.406)
VARIANCE = (0.229, 0.224, 0.225)


def resize_aspect_ratio_ref(img: np.ndarray, square_size, mag_ratio=1.0):
    """utils.py:497-524 with interpolation=cv2.INTER_LINEAR (detectors.py:35-37)."""
    h, w, ch = img.shape
    target = mag_ratio * max(h, w)
    if target > square_size:
        target = square_size
    ratio = target / max(h, w)
    th, tw = int(h * ratio), int(w * ratio)
    proc = cv2.resize(img, (tw, th), interpolation=cv2.INTER_LINEAR)
    h32 = th if th % 32 == 0 else th + (32 - th % 32)
    w32 = tw if tw % 32 == 0 else tw + (32 - tw % 32)
    canvas = np.zeros((h32, w32, ch), dtype=np.float32)
    canvas[:th, :tw, :] = proc
    return canvas, ratio, (int(w32 / 2), int(h32 / 2))


def normalize_mean_variance_ref(img: np.ndarray) -> np.ndarray:
    """utils.py:479-484: float32 arithmetic, the constants are float32(mean * 255.0)."""
    out = img.copy().astype(np.float32)
    out -= np.array([m * 255.0 for m in MEAN], dtype=np.float32)
    out /= np.array([v * 255.0 for v in VARIANCE], dtype=np.float32)
    return out


def detector_input_ref(img: np.ndarray, canvas_size, mag_ratio=1.0):
    """detectors.py:33-43 for one image: x float32 [3, H32, W32] and the target ratio."""
    canvas, ratio, _ = resize_aspect_ratio_ref(img, canvas_size, mag_ratio)
    return np.ascontiguousarray(np.transpose(normalize_mean_variance_ref(canvas), (2, 0, 1))), ratio


def _thresholds(textmap, linkmap, low_text, link_threshold):
    _, text_score = cv2.threshold(textmap, low_text, 1, 0)
    _, link_score = cv2.threshold(linkmap, link_threshold, 1, 0)
    comb = np.clip(text_score + link_score, 0, 1).astype(np.uint8)
    return text_score, link_score, comb


def get_det_boxes_core_ref(textmap, linkmap, text_threshold, link_threshold, low_text):
    """utils.py:106-175 with estimate_num_chars=False. Returns (det, labels, mapper)."""
    textmap = np.array(textmap, dtype=np.float32, copy=True)
    linkmap = np.array(linkmap, dtype=np.float32, copy=True)
    img_h, img_w = textmap.shape
    text_score, link_score, comb = _thresholds(textmap, linkmap, low_text, link_threshold)
    n_labels, labels, stats, _ = cv2.connectedComponentsWithStats(comb, connectivity=4)
    det, mapper = [], []
    for k in range(1, n_labels):
        size = stats[k, cv2.CC_STAT_AREA]
        if size < 10:
            continue
        member = labels == k
        if np.max(textmap[member]) < text_threshold:
            continue
        seg = np.zeros(textmap.shape, dtype=np.uint8)
        seg[member] = 255
        mapper.append(k)
        seg[np.logical_and(link_score == 1, text_score == 0)] = 0
        x, y = stats[k, cv2.CC_STAT_LEFT], stats[k, cv2.CC_STAT_TOP]
        w, h = stats[k, cv2.CC_STAT_WIDTH], stats[k, cv2.CC_STAT_HEIGHT]
        niter = int(math.sqrt(size * min(w, h) / (w * h)) * 2)
        sx, ex, sy, ey = x - niter, x + w + niter + 1, y - niter, y + h + niter + 1
        sx, sy = max(sx, 0), max(sy, 0)
        if ex >= img_w:
            ex = img_w
        if ey >= img_h:
            ey = img_h
        kernel = cv2.getStructuringElement(cv2.MORPH_RECT, (1 + niter, 1 + niter))
        seg[sy:ey, sx:ex] = cv2.dilate(seg[sy:ey, sx:ex], kernel)
        pts = np.roll(np.array(np.where(seg != 0)), 1, axis=0).transpose().reshape(-1, 2)
        box = cv2.boxPoints(cv2.minAreaRect(pts))
        w_, h_ = np.linalg.norm(box[0] - box[1]), np.linalg.norm(box[1] - box[2])
        if abs(1 - max(w_, h_) / (min(w_, h_) + 1e-5)) <= 0.1:
            l, r = min(pts[:, 0]), max(pts[:, 0])
            t, b = min(pts[:, 1]), max(pts[:, 1])
            box = np.array([[l, t], [r, t], [r, b], [l, b]], dtype=np.float32)
        start = box.sum(axis=1).argmin()
        det.append(np.array(np.roll(box, 4 - start, 0)))
    return det, labels, mapper


def adjust_result_coordinates_ref(polys, ratio_w, ratio_h, ratio_net=2):
    """utils.py:343-349."""
    if len(polys) > 0:
        polys = np.array(polys)
        for k in range(len(polys)):
            if polys[k] is not None:
                polys[k] *= (ratio_w * ratio_net, ratio_h * ratio_net)
    return polys


def textbox_from_maps_ref(score_text, score_link, target_ratio, text_threshold, link_threshold, low_text) -> List[np.ndarray]:
    """test_net's post-processing + get_textbox's flattening for one image with poly=False (detectors.py:49-70, 107-113):
    a list of int32 [8] arrays."""
    boxes, _, _ = get_det_boxes_core_ref(score_text, score_link, text_threshold, link_threshold, low_text)
    ratio = 1 / target_ratio
    boxes = adjust_result_coordinates_ref(boxes, ratio, ratio)
    return [np.array(b).astype(np.int32).reshape(-1) for b in boxes]


def components_ref(textmap, linkmap, text_threshold, link_threshold, low_text):
    """What ocrfa_det_components returns, computed with full masks: (n_labels, labels, comps, extents) where comps is a
    list of dicts (label, area, left, top, width, height, niter, sx, sy, ex, ey, row_off, max_text) in label order and
    extents an int32 [rows, 2] array of the dilated mask's (min x, max x) per ROI row (INT_MAX, -1 for empty rows)."""
    textmap = np.ascontiguousarray(textmap, dtype=np.float32)
    linkmap = np.ascontiguousarray(linkmap, dtype=np.float32)
    img_h, img_w = textmap.shape
    text_score, link_score, comb = _thresholds(textmap, linkmap, low_text, link_threshold)
    n_labels, labels, stats, _ = cv2.connectedComponentsWithStats(comb, connectivity=4)
    comps, ext = [], []
    for k in range(1, n_labels):
        size = int(stats[k, cv2.CC_STAT_AREA])
        if size < 10:
            continue
        member = labels == k
        mt = np.max(textmap[member])
        if mt < np.float32(text_threshold):
            continue
        seg = np.zeros(textmap.shape, dtype=np.uint8)
        seg[member] = 255
        seg[np.logical_and(link_score == 1, text_score == 0)] = 0
        x, y, w, h = (int(stats[k, c]) for c in (cv2.CC_STAT_LEFT, cv2.CC_STAT_TOP, cv2.CC_STAT_WIDTH, cv2.CC_STAT_HEIGHT))
        niter = int(math.sqrt(size * min(w, h) / (w * h)) * 2)
        sx, sy = max(x - niter, 0), max(y - niter, 0)
        ex, ey = x + w + niter + 1, y + h + niter + 1
        if ex >= img_w:
            ex = img_w
        if ey >= img_h:
            ey = img_h
        roi = cv2.dilate(seg[sy:ey, sx:ex], cv2.getStructuringElement(cv2.MORPH_RECT, (1 + niter, 1 + niter)))
        comps.append(dict(label=k, area=size, left=x, top=y, width=w, height=h, niter=niter, sx=sx, sy=sy, ex=ex, ey=ey,
                          row_off=len(ext), max_text=np.float32(mt)))
        for r in range(ey - sy):
            xs = np.nonzero(roi[r])[0]
            ext.append((sx + int(xs[0]), sx + int(xs[-1])) if len(xs) else (np.iinfo(np.int32).max, -1))
    return n_labels, labels.astype(np.int32), comps, np.array(ext, dtype=np.int32).reshape(-1, 2)


from ocrfa_b200.synth import synth_score_maps  # noqa: E402,F401  (plain data generation, shared with tools/)
