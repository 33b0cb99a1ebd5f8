"""Generates tests/golden/quantized_int8.npz: the reference recognizer's strings on 64 synthetic lines of 32x128 with
`quantize=False` (the fp32 oracle setting) and with `quantize=True` - the literal `OCR(gpu=False)` default, which runs
torch.quantization.quantize_dynamic(model, dtype=torch.qint8) over the LSTM / Linear layers (utils.py:686-690). The
int8 column is RECORDED, not a parity target: dynamic int8 arithmetic is a property of the reference's CPU deployment,
it cannot match fp32 bit for bit. TEST INFRASTRUCTURE, dev container only (needs /root/reference):

    python oracle/gen_golden_quant.py
"""
from __future__ import annotations

import os
import sys
import tempfile
import time
from collections import OrderedDict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle._refimport import import_reference  # noqa: E402
from oracle.synth import synth_lines, synth_state_dict  # noqa: E402


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    utils, _ = import_reference()
    spec = utils.recognition_models["gen1"]["arabic_g1"]
    character = spec["characters"]
    sd = synth_state_dict(0)
    tsd = OrderedDict(("module." + k, torch.from_numpy(v)) for k, v in sd.items())
    params = {"input_channel": 1, "output_channel": 512, "hidden_size": 512}
    imgs = synth_lines(64, 32, 128, seed=4242)
    image_list = [([[0, 0], [128, 0], [128, 32], [0, 32]], im) for im in imgs]
    import ocrfa_b200.charset as cs

    ignore_char = cs.default_ignore_char()
    out = {"seed": np.int64(4242)}
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "synthetic.pth")
        torch.save(tsd, path)
        for key, q in (("fp32", False), ("int8", True)):
            model, converter = utils.get_recognizer("generation1", params, character, {}, {}, path, device="cpu", quantize=q)
            kinds = sorted({type(m).__name__ for m in model.modules()})
            t0 = time.perf_counter()
            res = utils.get_text(character, 32, 128, model, converter, image_list, ignore_char, "greedy", 5, 32, 0.0,
                                 0.5, 0.003, 0, "cpu")
            dt = time.perf_counter() - t0
            out[key + "_text"] = np.array([r[1] for r in res] + [""])
            out[key + "_conf"] = np.array([float(r[2]) for r in res], dtype=np.float64)
            out[key + "_lines_per_s"] = np.float64(len(res) / dt)
            print(key, "quantized modules:" if q else "modules:", [k for k in kinds if "LSTM" in k or "Linear" in k],
                  f"{len(res) / dt:.1f} lines/s")
    same = sum(a == b for a, b in zip(out["fp32_text"][:-1], out["int8_text"][:-1]))
    print(f"int8 vs fp32 strings identical: {same}/64")
    path = os.path.join(ROOT, "tests", "golden", "quantized_int8.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
