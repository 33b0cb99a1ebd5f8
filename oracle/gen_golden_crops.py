"""Generates tests/golden/crops.npz by running the UNMODIFIED reference's get_image_list
(/root/reference/src/ocrfa/utils.py:1375-1418: cv2.resize with interpolation value 1 for every box, and
four_point_transform = cv2.warpPerspective for the rotated ones) on a synthetic page. TEST INFRASTRUCTURE, dev container only (needs /root/reference):

    python oracle/gen_golden_crops.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle._refimport import import_reference  # noqa: E402


def synth_page(h: int, w: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    page = rng.integers(180, 256, (h, w)).astype(np.float64)
    for _ in range(250):  # dark strokes of random size
        y, x = int(rng.integers(0, h - 4)), int(rng.integers(0, w - 4))
        hh, ww = int(rng.integers(2, 30)), int(rng.integers(2, 40))
        page[y:y + hh, x:x + ww] = rng.integers(0, 90)
    page += rng.normal(0, 6, page.shape)
    return np.clip(np.rint(page), 0, 255).astype(np.uint8)


def main():
    utils, _ = import_reference()
    page = synth_page(300, 520, 7)
    rng = np.random.default_rng(11)
    boxes = []
    for _ in range(20):  # [x_min, x_max, y_min, y_max], some reaching outside the page, some tall (vertical text)
        w = int(rng.integers(8, 300))
        h = int(rng.integers(6, 100))
        x0 = int(rng.integers(-20, 520 - 8))
        y0 = int(rng.integers(-10, 300 - 6))
        boxes.append([x0, x0 + w, y0, y0 + h])
    boxes += [[100, 228, 50, 114], [10, 138, 200, 264],      # exactly 2x the model size: OpenCV's INTER_AREA switch
              [300, 364, 100, 132], [5, 69, 5, 37],           # equal to the model size: copy
              [400, 410, 20, 200], [500, 519, 0, 299],        # tall boxes (ratio < 1 branch)
              [0, 520, 0, 32], [0, 520, 268, 300]]            # full-width lines
    # rotated quads (tl, tr, br, bl), some sticking out of the page: the four_point_transform / warpPerspective path
    import math
    free = []
    for _ in range(10):
        cx, cy = rng.uniform(20, 500), rng.uniform(10, 290)
        bw, bh, ang = rng.uniform(20, 260), rng.uniform(8, 60), rng.uniform(-0.9, 0.9)
        c, s_ = math.cos(ang), math.sin(ang)
        pts = np.array([[-bw / 2, -bh / 2], [bw / 2, -bh / 2], [bw / 2, bh / 2], [-bw / 2, bh / 2]])
        quad = pts @ np.array([[c, s_], [-s_, c]]) + [cx, cy] + rng.uniform(-2, 2, (4, 2))
        free.append(np.rint(quad).astype(int).tolist())  # the reference's boxes are integer pixel coordinates
    out = {"page": page, "boxes": np.array(boxes, dtype=np.int64), "free": np.array(free, dtype=np.int64)}
    for mh in (32, 64):
        image_list, max_width = utils.get_image_list(boxes, free, page, model_height=mh, sort_output=True)
        out[f"mh{mh}_max_width"] = np.int64(max_width)
        out[f"mh{mh}_n"] = np.int64(len(image_list))
        for i, (coords, crop) in enumerate(image_list):
            out[f"mh{mh}_box{i}"] = np.array(coords, dtype=np.int64)
            out[f"mh{mh}_crop{i}"] = np.ascontiguousarray(crop, dtype=np.uint8)
    path = os.path.join(ROOT, "tests", "golden", "crops.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v.shape if hasattr(v, "shape") else v) for k, v in list(out.items())[:6]}, "...")


if __name__ == "__main__":
    main()
