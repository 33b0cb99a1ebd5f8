"""Generates tests/golden/craft_boxes.npz: the UNMODIFIED reference detector end to end - test_net / get_textbox
(/root/reference/src/ocrfa/detectors.py:24-115) with the REAL network forward (utils.py:48-98 class OCRFA) loaded with
the synthetic weights of ocrfa_b200.craft.synth_craft_state_dict(0), on synthetic pages - i.e. the reference's boxes for
the reference's own score maps. A random-weight CRAFT does not produce text scores in [0, 1]: the three thresholds are
taken from quantiles of its own maps (stored in the fixture) and put in the middle of a gap between map values that is at least 4e-5 of the map's range wide, so that an
fp32-grade evaluation of the network cannot move a pixel across a threshold. TEST INFRASTRUCTURE, dev container only
(needs /root/reference and torchvision):  python oracle/gen_golden_craft_boxes.py
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle._refimport import import_reference  # noqa: E402
from oracle.gen_golden_det import synth_rgb  # noqa: E402
from ocrfa_b200.craft import synth_craft_state_dict  # noqa: E402


def clear_threshold(values: np.ndarray, q: float, rel_gap: float = 2e-5) -> float:
    """A threshold near the q-quantile of `values` that no value comes closer to than rel_gap of the value range: the
    midpoint of the first gap of at least twice that width at or above the quantile."""
    v = np.sort(values.ravel().astype(np.float64))
    gap = rel_gap * float(v[-1] - v[0])
    i = int(q * (len(v) - 1))
    while i + 1 < len(v) and v[i + 1] - v[i] < 2.0 * gap:
        i += 1
    if i + 1 >= len(v):
        raise RuntimeError("no clear threshold found")
    return float(0.5 * (v[i] + v[i + 1]))


def main():
    utils, _ = import_reference()
    det = importlib.import_module("src.ocrfa.detectors")
    torch.manual_seed(0)
    net = utils.OCRFA(pretrained=False)
    sd = {k: torch.from_numpy(v) for k, v in synth_craft_state_dict(0).items()}
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not [m for m in missing if "num_batches_tracked" not in m] and not unexpected
    net.eval()
    out = {}
    cases = [(192, 288, 2560, 1.0, 21), (150, 400, 320, 1.0, 22)]  # (h, w, canvas_size, mag_ratio, seed): copy and shrink
    out["cases"] = np.array(cases, dtype=np.float64)
    for i, (h, w, canvas, mag, seed) in enumerate(cases):
        page = synth_rgb(h, w, seed)
        resized, ratio, _ = utils.resize_aspect_ratio(page, canvas, interpolation=1, mag_ratio=mag)
        x = torch.from_numpy(np.transpose(utils.normalize_mean_variance(resized), (2, 0, 1))[None])
        with torch.no_grad():
            y, _ = net(x)
        text, link = y[0, :, :, 0].numpy(), y[0, :, :, 1].numpy()
        low = clear_threshold(text, 0.80)
        tt = clear_threshold(text, 0.93)
        lt = clear_threshold(link, 0.90)
        res = det.get_textbox(net, page, canvas, mag, tt, lt, low, False, "cpu")
        boxes = res[0]
        print(f"case {i}: page {h}x{w} -> maps {text.shape}, thresholds text {tt:.6f} link {lt:.6f} low {low:.6f}, {len(boxes)} boxes")
        assert len(boxes) >= 3
        out[f"c{i}_page"] = page
        out[f"c{i}_thr"] = np.array([tt, lt, low], dtype=np.float64)
        out[f"c{i}_text"], out[f"c{i}_link"] = text, link
        out[f"c{i}_boxes"] = np.array(boxes, dtype=np.int32).reshape(-1, 8)
    path = os.path.join(ROOT, "tests", "golden", "craft_boxes.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
