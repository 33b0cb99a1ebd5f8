"""Generates tests/golden/*.npz by running the UNMODIFIED reference (dev container only; TEST INFRASTRUCTURE).

    python oracle/gen_golden.py            # needs /root/reference; ~2 min on 8 cores

What is recorded (all produced by reference code, cited):
  * charset / ignore_idx of the default Farsi setup    (ocr.py:273-297, 384-389; utils.py:408-416, 716-722)
  * AlignCollate outputs                               (utils.py:567-602)
  * modules.Model logits on synthetic weights          (modules/blocks.py:422-447)
  * recognizer_predict strings + confidences for greedy / beamsearch / wordbeamsearch (utils.py:605-662)
  * get_text two-pass results                          (utils.py:698-770)
  * CTCLabelConverter.decode_* on hand-built probability matrices: ties, all-blank, flat rows, T=1,
    zero-probability underflow                          (utils.py:1099-1173)
The synthetic weights themselves are NOT stored (215 MB); oracle/synth.py regenerates them from the seed.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from collections import OrderedDict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle._refimport import import_reference  # noqa: E402
from oracle.synth import synth_lines, synth_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
WEIGHT_SEED = 0


def build_reference_recognizer(utils, character):
    sd = synth_state_dict(WEIGHT_SEED)
    tsd = OrderedDict(("module." + k, torch.from_numpy(v)) for k, v in sd.items())
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "synthetic.pth")
        torch.save(tsd, path)
        params = {"input_channel": 1, "output_channel": 512, "hidden_size": 512}
        model, converter = utils.get_recognizer("generation1", params, character, {}, {}, path, device="cpu",
                                                quantize=False)
    model.eval()
    return model, converter


def run_predict(utils, model, converter, x, ignore_idx, decoder, beam):
    res = utils.recognizer_predict(model, converter, [x], int(x.shape[3] / 10), ignore_idx, {}, decoder, beam, "cpu")
    return [r[0] for r in res], np.array([float(r[1]) for r in res], dtype=np.float64)


def main():
    import importlib

    torch.set_num_threads(max(1, os.cpu_count() or 1))
    utils, _modules = import_reference()
    ocr = importlib.import_module("src.ocrfa.ocr")
    os.makedirs(OUT, exist_ok=True)

    spec = utils.recognition_models["gen1"]["arabic_g1"]
    character = spec["characters"]
    dummy = types.SimpleNamespace()
    ocr.OCR.setLanguageList(dummy, ["fa"], spec)
    ignore_char = "".join(set(character) - set(dummy.lang_char))
    ignore_idx = []
    for ch in ignore_char:
        ignore_idx.append(character.index(ch) + 1)
    ignore_idx = sorted(ignore_idx)
    model, converter = build_reference_recognizer(utils, character)
    np.savez_compressed(os.path.join(OUT, "charset.npz"),
                        character=np.array(character), symbols=np.array(spec["symbols"]),
                        lang_char=np.array("".join(sorted(dummy.lang_char))),
                        ignore_char=np.array("".join(sorted(ignore_char))),
                        ignore_idx=np.array(ignore_idx, dtype=np.int64),
                        converter_character=np.array(converter.character),
                        converter_ignore_idx=np.array(converter.ignore_idx, dtype=np.int64),
                        num_class=np.array(len(converter.character)))

    # ---------------- forward + decode fixtures -------------------------------------------------------------
    configs = [("c1_32x128", 3, 32, 128, 11), ("c2_32x512", 3, 32, 512, 12), ("c3_64x256", 2, 64, 256, 13),
               ("c3_64x2048", 1, 64, 2048, 14)]
    for name, n, h, w, seed in configs:
        imgs = synth_lines(n, h, w, seed)
        collate = utils.AlignCollate(imgH=h, imgW=w, keep_ratio_with_pad=True)
        from PIL import Image
        x = collate([Image.fromarray(im, "L") for im in imgs])
        with torch.no_grad():
            logits = model(x, torch.zeros(n, int(w / 10) + 1, dtype=torch.long))
        probs = torch.nn.functional.softmax(logits, dim=2).cpu().numpy()
        probs[:, :, ignore_idx] = 0.0
        probs = probs / np.expand_dims(probs.sum(axis=2), axis=-1)
        out = dict(images=imgs, x=x.numpy(), logits=logits.numpy(), probs=probs.astype(np.float32))
        for dec, beam in (("greedy", 5), ("beamsearch", 5), ("beamsearch", 8), ("wordbeamsearch", 8)):
            if name == "c3_64x2048" and dec != "greedy" and beam != 8:
                continue
            s, c = run_predict(utils, model, converter, x, ignore_idx, dec, beam)
            out[f"str_{dec}_{beam}"] = np.array(s)
            out[f"conf_{dec}_{beam}"] = c
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, "greedy:", [s[:24] for s in out["str_greedy_5"]], "conf", out["conf_greedy_5"])

    # ---------------- get_text two-pass + ragged widths (section 8f rank 1) ------------------------------------
    rag = [synth_lines(1, 32, 100, 21)[0], synth_lines(1, 40, 333, 22)[0], synth_lines(1, 64, 512, 23)[0],
           synth_lines(1, 32, 128, 24)[0]]
    img_h, img_w = 32, 320
    collate = utils.AlignCollate(imgH=img_h, imgW=img_w, keep_ratio_with_pad=True)
    collate_c = utils.AlignCollate(imgH=img_h, imgW=img_w, keep_ratio_with_pad=True, adjust_contrast=0.5)
    from PIL import Image
    xr = collate([Image.fromarray(im, "L") for im in rag]).numpy()
    xc = collate_c([Image.fromarray(im, "L") for im in rag]).numpy()
    image_list = [([[0, 0], [1, 0], [1, 1], [0, 1]], im) for im in rag]
    res = utils.get_text(character, img_h, img_w, model, converter, image_list, ignore_char, "greedy", 5, 4, 0.1, 0.5,
                         0.003, 0, "cpu")
    rag_out = {f"img{i}": im for i, im in enumerate(rag)}
    rag_out.update(x=xr, x_contrast=xc, img_h=np.array(img_h), img_w=np.array(img_w),
                   text=np.array([r[1] for r in res]), conf=np.array([float(r[2]) for r in res]))
    np.savez_compressed(os.path.join(OUT, "ragged_get_text.npz"), **rag_out)
    print("ragged get_text:", [r[1][:20] for r in res], [float(r[2]) for r in res])

    # ---------------- decode-only fixtures on hand-built probability matrices -----------------------------------
    rng = np.random.Generator(np.random.PCG64(77))
    C = len(converter.character)
    mats = []

    def softmax(z):
        z = z - z.max(axis=-1, keepdims=True)
        e = np.exp(z)
        return (e / e.sum(axis=-1, keepdims=True)).astype(np.float32)

    mats.append(softmax(rng.standard_normal((40, C)).astype(np.float32) * 6))            # peaky random
    mats.append(softmax(rng.standard_normal((40, C)).astype(np.float32) * 1.5))          # soft random
    flat = np.full((40, C), 1.0 / C, dtype=np.float32)
    mats.append(flat)                                                                     # perfectly flat: all ties
    blank = np.zeros((40, C), dtype=np.float32); blank[:, 0] = 1.0
    mats.append(blank)                                                                    # all blank
    rep = np.zeros((40, C), dtype=np.float32)
    seq = [5, 5, 0, 5, 7, 7, 7, 0, 0, 9, 9, 37, 37, 0, 37] + [0] * 25
    for t, k in enumerate(seq):
        rep[t, k] = 0.9; rep[t, (k + 1) % C] = 0.1
    mats.append(rep)                                                                      # repeats / blank between same char
    tie = np.zeros((40, C), dtype=np.float32)
    for t in range(40):
        a, b = int(rng.integers(0, C)), int(rng.integers(0, C))
        tie[t, a] += 0.5; tie[t, b] += 0.5
    mats.append(tie)                                                                      # exact 2-way ties
    tiny = softmax(rng.standard_normal((40, C)).astype(np.float32) * 60)                  # underflow to exact zeros
    mats.append(tiny)
    sp = softmax(rng.standard_normal((40, C)).astype(np.float32) * 5)
    space = converter.dict[" "]
    for t in (6, 7, 19, 30, 39):
        sp[t] = 0.0; sp[t, space] = 1.0
    mats.append(sp)                                                                       # word boundaries
    mat = np.stack(mats, 0)
    dec = dict(mat=mat)
    idx = mat.argmax(axis=2).reshape(-1)
    dec["greedy"] = np.array(converter.decode_greedy(idx, torch.IntTensor([mat.shape[1]] * mat.shape[0])))
    for bw in (1, 2, 5, 8, 25):
        dec[f"beam_{bw}"] = np.array(converter.decode_beamsearch(mat, beamWidth=bw))
    for bw in (5, 8):
        dec[f"wordbeam_{bw}"] = np.array(converter.decode_wordbeamsearch(mat, beamWidth=bw))
    one = mat[:, :1, :].copy()
    dec["mat_t1"] = one
    dec["greedy_t1"] = np.array(converter.decode_greedy(one.argmax(axis=2).reshape(-1), torch.IntTensor([1] * one.shape[0])))
    dec["beam_t1_8"] = np.array(converter.decode_beamsearch(one, beamWidth=8))
    np.savez_compressed(os.path.join(OUT, "decode_cases.npz"), **dec)
    for k in ("greedy", "beam_5", "beam_8", "wordbeam_8"):
        print(k, [s[:16] for s in dec[k]])
    print("wrote fixtures to", OUT)


if __name__ == "__main__":
    main()
