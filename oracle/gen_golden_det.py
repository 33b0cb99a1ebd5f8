"""Generates tests/golden/detection.npz by running the UNMODIFIED reference's detector pre-/post-processing:
resize_aspect_ratio + normalize_mean_variance (utils.py:479-524), get_det_boxes (utils.py:329-340), get_textbox /
test_net (detectors.py:24-115, fed by a stand-in `net` that returns prepared score maps, so the real code path
including adjust_result_coordinates runs), group_text_box (utils.py:1235-1348) with OCR.transform_ua's arguments and
min_size filter (ocr.py:283-355). TEST INFRASTRUCTURE, dev container only (needs /root/reference):

    python oracle/gen_golden_det.py
"""
from __future__ import annotations

import importlib
import os
import sys

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle._refimport import import_reference  # noqa: E402
from oracle.det_ref import synth_score_maps  # noqa: E402


def synth_rgb(h, w, seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(150, 256, (h, w, 3)).astype(np.float64)
    for _ in range(40):
        y, x = int(rng.integers(0, h - 3)), int(rng.integers(0, w - 3))
        img[y:y + int(rng.integers(2, 12)), x:x + int(rng.integers(2, 30))] = rng.integers(0, 100, 3)
    return np.clip(np.rint(img + rng.normal(0, 5, img.shape)), 0, 255).astype(np.uint8)


def pack_boxes(prefix, boxes, out):
    out[prefix + "_n"] = np.int64(len(boxes))
    for i, b in enumerate(boxes):
        out[f"{prefix}_{i}"] = np.asarray(b)


def main():
    utils, _ = import_reference()
    det = importlib.import_module("src.ocrfa.detectors")
    out = {}
    # 1. detector input: (image shape, canvas_size, mag_ratio) covering copy, shrink, exact 2x shrink, magnify
    pre_cases = [((70, 100), 2560, 1.0), ((70, 100), 64, 1.0), ((128, 200), 100, 1.0), ((50, 90), 2560, 1.5),
                 ((97, 33), 80, 1.0)]
    out["pre_cases"] = np.array([[h, w, c, int(m * 100)] for (h, w), c, m in pre_cases], dtype=np.int64)
    for i, ((h, w), canvas, mag) in enumerate(pre_cases):
        img = synth_rgb(h, w, 100 + i)
        resized, ratio, heat = utils.resize_aspect_ratio(img, canvas, interpolation=cv2.INTER_LINEAR, mag_ratio=mag)
        x = np.transpose(utils.normalize_mean_variance(resized), (2, 0, 1))
        out[f"pre{i}_img"] = img
        out[f"pre{i}_x"] = np.ascontiguousarray(x, dtype=np.float32)
        out[f"pre{i}_ratio"] = np.float64(ratio)
        out[f"pre{i}_heat"] = np.array(heat, dtype=np.int64)
    # 2. get_det_boxes on synthetic score maps, default thresholds of OCR.readtext and a second set
    map_cases = [(96, 160, 3, 14), (192, 320, 5, 40), (64, 64, 9, 6), (48, 208, 11, 10)]
    thr_sets = [(0.7, 0.4, 0.4), (0.5, 0.3, 0.25)]  # (text_threshold, link_threshold, low_text)
    out["map_cases"] = np.array(map_cases, dtype=np.int64)
    out["thr_sets"] = np.array(thr_sets, dtype=np.float64)
    for i, (h, w, seed, n_words) in enumerate(map_cases):
        text, link = synth_score_maps(h, w, seed, n_words)
        out[f"map{i}_text"], out[f"map{i}_link"] = text, link
        for j, (tt, lt, low) in enumerate(thr_sets):
            boxes, polys, mapper = utils.get_det_boxes(text, link, tt, lt, low, False)
            pack_boxes(f"map{i}_thr{j}_box", boxes, out)
            out[f"map{i}_thr{j}_mapper"] = np.array(mapper, dtype=np.int64)
            # get_textbox through the real test_net: the page is 2x the map (CRAFT's stride), canvas large enough -> ratio 1
            page = synth_rgb(2 * h, 2 * w, 7)
            y = torch.from_numpy(np.stack([text, link], axis=-1))[None]

            def fake_net(x, y=y):
                assert x.shape[2] == y.shape[1] * 2 and x.shape[3] == y.shape[2] * 2
                return y, None

            res = det.get_textbox(fake_net, page, 2560, 1.0, tt, lt, low, False, "cpu")
            assert len(res) == 1
            pack_boxes(f"map{i}_thr{j}_poly", res[0], out)
            # the same with a canvas that forces a down-scale (ratio < 1): coordinates are scaled back by 1/ratio
            if i == 0:
                big = synth_rgb(3 * h, 3 * w, 8)  # resized to 2h x 2w by canvas_size = 2w
                res2 = det.get_textbox(fake_net, big, 2 * w, 1.0, tt, lt, low, False, "cpu")
                pack_boxes(f"map{i}_thr{j}_poly_scaled", res2[0], out)
            # OCR.transform_ua: group_text_box + min_size filters (ocr.py:338-353)
            hl, fl = utils.group_text_box(res[0], 0.1, 0.5, 0.5, 0.5, 0.1, True)
            hl = [b for b in hl if max(b[1] - b[0], b[3] - b[2]) > 20]
            fl = [b for b in fl if max(utils.diff([c[0] for c in b]), utils.diff([c[1] for c in b])) > 20]
            out[f"map{i}_thr{j}_hlist"] = np.array(hl, dtype=np.float64).reshape(-1, 4)
            out[f"map{i}_thr{j}_flist"] = np.array(fl, dtype=np.float64).reshape(-1, 4, 2)
    path = os.path.join(ROOT, "tests", "golden", "detection.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;",
          {k: int(out[k]) for k in out if k.endswith("_box_n") or k.endswith("_poly_n")})


if __name__ == "__main__":
    main()
