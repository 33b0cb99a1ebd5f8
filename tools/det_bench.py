#!/usr/bin/env python
"""Timing of the detection glue and of OCR.readtext on one B200, beside the reference's CPU post-processing.

    python tools/det_bench.py [--reps 20]

Prints one JSON line per measurement:
  det_components   ocrfa_det_components alone (CUDA events) on CRAFT-sized score maps, and the whole get_det_boxes call
                   (kernels + D2H of the extents + cv2.minAreaRect per box) against the oracle's restatement of the
                   reference's get_det_boxes_core (NumPy/OpenCV on the host, the reference's own method)
  detector_input   resize + normalise + CHW kernel on a page
  readtext         OCR.readtext on a synthetic 736 x 1280 page with synthetic weights: CRAFT forward + labeling + crops +
                   recognition, batch_size 64
Only the CPU-baseline leg imports oracle/ (as the thing compared against, never as the product path).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import ocrfa_b200 as ob  # noqa: E402
from ocrfa_b200 import _lib  # noqa: E402
from ocrfa_b200.detection import DET_COMP, det_components, detector_input, get_det_boxes  # noqa: E402


def cuda_ms(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def wall_ms(fn, reps):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) * 1e3 / reps


def cpu_readtext_legs(reader, rgb, tt, lt, low):
    """The reference's CPU path for the same page, leg by leg, with the oracle's restatements (torch CPU fp32 with all
    host threads for the two networks, OpenCV / NumPy for the rest): one pass, the timed thing is the oracle."""
    import cv2
    from oracle import craft_ref, det_ref as D, recognizer_ref as R
    from ocrfa_b200.detection import group_text_box
    from ocrfa_b200.synth import synth_state_dict

    torch.set_num_threads(os.cpu_count() or 1)
    legs = {}

    def leg(name, fn):
        t0 = time.perf_counter()
        out = fn()
        legs[name + "_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
        return out

    grey = leg("grey", lambda: cv2.cvtColor(rgb, cv2.COLOR_BGR2GRAY))
    x, ratio = leg("detector_input", lambda: D.detector_input_ref(rgb, 2560, 1.0))
    sd_craft = ob.synth_craft_state_dict(0)
    y, _ = leg("craft_forward", lambda: craft_ref.craft_forward_ref(sd_craft, x[None]))
    y = y[0].numpy()
    polys = leg("boxes", lambda: D.textbox_from_maps_ref(np.ascontiguousarray(y[:, :, 0]), np.ascontiguousarray(y[:, :, 1]),
                                                        ratio, tt, lt, low))
    h_list, f_list = leg("grouping", lambda: group_text_box(polys, 0.1, 0.5, 0.5, 0.5, 0.1, True))
    h_list = [b for b in h_list if max(b[1] - b[0], b[3] - b[2]) > 10]
    f_list = [b for b in f_list if max(max(c[0] for c in b) - min(c[0] for c in b),
                                       max(c[1] for c in b) - min(c[1] for c in b)) > 10]
    def crops_cv2():  # get_image_list with the library calls the reference makes (utils.py:1218-1231, 1361-1418)
        out, max_ratio = [], 1.0
        for quad in f_list:
            rect = np.array(quad, dtype="float32")
            (tl, tr, br, bl) = rect
            mw = max(int(np.linalg.norm(br - bl)), int(np.linalg.norm(tr - tl)))
            mh = max(int(np.linalg.norm(tr - br)), int(np.linalg.norm(tl - bl)))
            if mw <= 0 or mh <= 0:
                continue
            dst = np.array([[0, 0], [mw - 1, 0], [mw - 1, mh - 1], [0, mh - 1]], dtype="float32")
            out.append((quad[0][1], cv2.warpPerspective(grey, cv2.getPerspectiveTransform(rect, dst), (mw, mh))))
        for b in h_list:
            x0, x1, y0, y1 = max(0, int(b[0])), min(int(b[1]), grey.shape[1]), max(0, int(b[2])), min(int(b[3]), grey.shape[0])
            if x1 > x0 and y1 > y0:
                out.append((y0, grey[y0:y1, x0:x1]))
        resized = []
        for key, crop in out:
            h, w = crop.shape
            ratio = w / h
            if ratio < 1.0:
                ratio = 1.0 / ratio
                size = (64, int(64 * ratio))
            else:
                size = (int(64 * ratio), 64)
            if min(size) == 0:
                continue
            max_ratio = max(max_ratio, ratio)
            resized.append((key, cv2.resize(crop, size, interpolation=1)))
        resized.sort(key=lambda kc: kc[0])
        return [(None, c) for _, c in resized], int(np.ceil(max_ratio)) * 64

    image_list, max_width = leg("crops", crops_cv2)

    def collate_pil(lines, img_h, img_w, contrast):  # AlignCollate / NormalizePAD with Pillow, as the reference (utils.py:533-602)
        import math
        from PIL import Image

        out = np.zeros((len(lines), 1, img_h, img_w), np.float32)
        for i, im in enumerate(lines):
            if contrast > 0:
                im = R.adjust_contrast_grey_ref(im, contrast)
            h, w = im.shape
            rw = min(img_w, math.ceil(img_h * (w / float(h))))
            a = (np.asarray(Image.fromarray(im, "L").resize((rw, img_h), Image.BICUBIC), np.float32) / 255.0 - 0.5) / 0.5
            out[i, 0, :, :rw] = a
            out[i, 0, :, rw:] = a[:, rw - 1:rw]
        return out

    sd = synth_state_dict(0)
    chars = ["[blank]"] + list(reader.character)
    ign = R.ignore_idx_ref(reader.character, "".join(set(reader.character) - set(reader.lang_char)))
    crops = [c for _, c in image_list]

    def recognize():
        n = 0
        for contrast in (0.0, 0.5):  # with untrained weights every line takes the low-confidence second pass
            for i in range(0, len(crops), 64):
                xb = collate_pil(crops[i:i + 64], 64, int(max_width), contrast)
                n += len(R.recognizer_predict_ref(sd, xb, chars, ign, "greedy", 5))
        return n

    leg("recognition", recognize)
    legs["lines"] = len(crops)
    legs["threads"] = os.cpu_count() or 1
    legs["total_ms"] = round(sum(v for k, v in legs.items() if k.endswith("_ms")), 1)
    return legs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU leg of the readtext line (about a minute)")
    args = ap.parse_args()
    from oracle import det_ref as D  # CPU baseline + synthetic score maps

    lib = _lib.load()
    for (h, w, words) in ((368, 640, 160), (1280, 1280, 700)):
        text, link = D.synth_score_maps(h, w, 31, words)
        y = torch.from_numpy(np.stack([text, link], axis=-1)).cuda()
        t, l = y[:, :, 0], y[:, :, 1]
        n = h * w
        max_rows, max_comps = n + 64, n // 10 + 1
        ws_bytes = int(lib.ocrfa_det_workspace_bytes(h, w, max_rows))
        ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device="cuda")
        off = (-ws.data_ptr()) % 256
        comps = torch.empty(max_comps * DET_COMP.itemsize, dtype=torch.uint8, device="cuda")
        ext = torch.empty((max_rows, 2), dtype=torch.int32, device="cuda")
        counts = torch.empty(4, dtype=torch.int32, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream

        def kernels():
            _lib.check(lib.ocrfa_det_components(C.c_void_p(t.data_ptr()), C.c_void_p(l.data_ptr()), 2, h, w, C.c_float(0.4),
                                                C.c_float(0.4), C.c_float(0.7), None, C.c_void_p(comps.data_ptr()), max_comps,
                                                C.c_void_p(ext.data_ptr()), max_rows, C.c_void_p(counts.data_ptr()),
                                                C.c_void_p(ws.data_ptr() + off), ws_bytes, C.c_void_p(stream)))

        k_ms = cuda_ms(kernels, args.reps)
        call_ms = wall_ms(lambda: get_det_boxes(t, l, 0.7, 0.4, 0.4), args.reps)
        n_boxes = len(get_det_boxes(t, l, 0.7, 0.4, 0.4)[0])
        t0 = time.perf_counter()
        cpu_reps = 3
        for _ in range(cpu_reps):
            ref = D.get_det_boxes_core_ref(text, link, 0.7, 0.4, 0.4)[0]
        cpu_ms = (time.perf_counter() - t0) * 1e3 / cpu_reps
        assert len(ref) == n_boxes
        # algorithmic bytes: two float maps read once (8 B/pixel); everything else is scratch
        print(json.dumps({"what": "det_components", "map": [h, w], "boxes": n_boxes, "kernels_ms": round(k_ms, 4),
                          "kernels_algorithmic_GBps": round(n * 8 / k_ms / 1e6, 1),
                          "get_det_boxes_call_ms": round(call_ms, 3), "reference_cpu_ms": round(cpu_ms, 2),
                          "speedup_call_vs_cpu": round(cpu_ms / call_ms, 1)}), flush=True)

    rng = np.random.default_rng(0)
    page = torch.from_numpy(rng.integers(0, 256, (2200, 3400, 3), dtype=np.uint8)).cuda()
    for canvas in (2560, 3400):
        ms = cuda_ms(lambda: detector_input(page, canvas, 1.0), args.reps)
        th, tw = detector_input(page, canvas, 1.0)[0].shape[2:]
        img = page.cpu().numpy()
        t0 = time.perf_counter()
        D.detector_input_ref(img, canvas, 1.0)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        print(json.dumps({"what": "detector_input", "page": [2200, 3400], "canvas": canvas, "out": [int(th), int(tw)],
                          "ms": round(ms, 4), "out_GBps": round(th * tw * 12 / ms / 1e6, 1),
                          "reference_cpu_ms": round(cpu_ms, 1)}), flush=True)

    # readtext end to end with synthetic weights
    from oracle.synth import synth_state_dict

    with tempfile.TemporaryDirectory() as d:
        torch.save({k: torch.from_numpy(v) for k, v in synth_state_dict(0).items()}, os.path.join(d, "arabic.pth"))
        torch.save({k: torch.from_numpy(v) for k, v in ob.synth_craft_state_dict(0).items()},
                   os.path.join(d, "craft_mlt_25k.pth"))
        reader = ob.OCR(model_storage_directory=d, verify_md5=False)
    rgb = torch.from_numpy(rng.integers(0, 256, (736, 1280, 3), dtype=np.uint8)).cuda()
    x, _ = detector_input(rgb, 2560, 1.0)
    yy, _ = reader.detector(x)
    tmap, lmap = yy[0, :, :, 0], yy[0, :, :, 1]
    low, tt = float(torch.quantile(tmap.flatten()[:1 << 20], 0.80)), float(torch.quantile(tmap.flatten()[:1 << 20], 0.93))
    lt = float(torch.quantile(lmap.flatten()[:1 << 20], 0.97))
    kw = dict(batch_size=64, text_threshold=tt, low_text=low, link_threshold=lt, min_size=10)
    res = reader.readtext(rgb, **kw)
    ms = wall_ms(lambda: reader.readtext(rgb, **kw), 5)
    det_ms = wall_ms(lambda: reader.transform_ua(rgb, reformat=False, text_threshold=tt, low_text=low, link_threshold=lt,
                                                 min_size=10), 5)
    line = {"what": "readtext", "page": [736, 1280], "lines": len(res), "batch_size": 64,
            "readtext_wall_ms": round(ms, 2), "detect_wall_ms": round(det_ms, 2),
            "note": "synthetic weights; thresholds taken from the score-map quantiles so that boxes exist"}
    if not args.no_cpu:
        line["reference_cpu"] = cpu_readtext_legs(reader, rgb.cpu().numpy(), tt, lt, low)
        line["speedup_vs_cpu"] = round(line["reference_cpu"]["total_ms"] / ms, 1)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
