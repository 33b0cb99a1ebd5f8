"""Generates tests/golden/reader.npz by driving the UNMODIFIED reference's OCR class (ocr.py: recognize :357-502,
readtext :504-582, transform_ua :283-355) and result helpers (utils.py: get_paragraph :1442, merge_to_free :1176,
set_result_with_confidence :1615). The OCR object is assembled without its downloading constructor: synthetic
recognizer weights (oracle/synth.py, seed 0, quantize=False), and for `readtext` a stand-in detector that returns the
score maps of tests/golden/detection.npz so that every step after the CRAFT forward is the reference's own code.
TEST INFRASTRUCTURE, dev container only (needs /root/reference):

    python oracle/gen_golden_reader.py
"""
from __future__ import annotations

import importlib
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle._refimport import import_reference  # noqa: E402
from oracle.gen_golden import build_reference_recognizer  # noqa: E402
from oracle.gen_golden_crops import synth_page  # noqa: E402
from oracle.gen_golden_det import synth_rgb  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pack_result(out, key, result):
    """[(box, text, conf)] -> arrays (boxes may be ragged lists of 4 points)."""
    out[key + "_n"] = np.int64(len(result))
    out[key + "_text"] = np.array([r[1] for r in result] + [""])  # + "" keeps a unicode dtype when empty
    out[key + "_conf"] = np.array([float(r[2]) for r in result], dtype=np.float64)
    out[key + "_box"] = np.array([np.array(r[0], dtype=np.float64) for r in result], dtype=np.float64).reshape(-1, 4, 2)


def rotation_cases(reader, page, h_list, f_list, out):
    """E. rotation_info test-time augmentation (ocr.py:446-473, utils.py:1599-1631): scipy.ndimage.rotate by right angles."""
    reader.device = torch.device("cpu")
    res = reader.recognize(page, h_list[:4], f_list[:1], rotation_info=[90, 180, 270], batch_size=4, reformat=False)
    pack_result(out, "rec_rot", res)
    res = reader.recognize(page, h_list[4:6], [], rotation_info=[180], batch_size=1, reformat=False, img_height=32)
    pack_result(out, "rec_rot180", res)
    reader.device = "cpu"


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    utils, _ = import_reference()
    ocr_mod = importlib.import_module("src.ocrfa.ocr")
    det_mod = importlib.import_module("src.ocrfa.detectors")
    spec = utils.recognition_models["gen1"]["arabic_g1"]
    reader = object.__new__(ocr_mod.OCR)
    reader.device = "cpu"
    reader.model_lang = "arabic"
    reader.character = spec["characters"]
    reader.setLanguageList(["fa"], spec)
    reader.recognizer, reader.converter = build_reference_recognizer(utils, reader.character)
    reader.get_textbox = det_mod.get_textbox
    out = {}

    crops = np.load(os.path.join(GOLDEN, "crops.npz"))
    page = crops["page"]  # 300 x 520 grey
    h_list = [b for b in crops["boxes"].tolist()][:10]
    f_list = crops["free"].tolist()[:3]
    path = os.path.join(GOLDEN, "reader.npz")
    if "--add-rotation" in sys.argv:  # extend the existing fixture without re-running the other cases
        out = dict(np.load(path))
        rotation_cases(reader, page, h_list, f_list, out)
        np.savez_compressed(path, **out)
        print("added rotation cases:", out["rec_rot_text"], out["rec_rot_conf"], out["rec_rot180_text"])
        return
    rotation_cases(reader, page, h_list, f_list, out)
    out["rec_h_list"] = np.array(h_list, dtype=np.int64)
    out["rec_f_list"] = np.array(f_list, dtype=np.int64)
    # A. per-box path (batch_size=1): every line padded to its own width
    res_a = reader.recognize(page, h_list, f_list, batch_size=1, reformat=False)
    pack_result(out, "rec_a", res_a)
    # B. batched path: the reference takes it only off the CPU device string, so hand it a torch.device
    reader.device = torch.device("cpu")
    res_b = reader.recognize(page, h_list, f_list, batch_size=4, reformat=False)
    pack_result(out, "rec_b", res_b)
    res_beam = reader.recognize(page, h_list[:4], [], decoder="beamsearch", beamWidth=5, batch_size=4, reformat=False)
    pack_result(out, "rec_beam", res_beam)
    res_word = reader.recognize(page, h_list[:4], [], decoder="wordbeamsearch", beamWidth=5, batch_size=4, reformat=False)
    pack_result(out, "rec_word", res_word)
    res_allow = reader.recognize(page, h_list[:4], [], batch_size=4, reformat=False, allowlist="0123456789ابپتث")
    pack_result(out, "rec_allow", res_allow)
    res_block = reader.recognize(page, h_list[:4], [], batch_size=4, reformat=False, blocklist="0123456789دزش")
    pack_result(out, "rec_block", res_block)
    reader.device = "cpu"
    # C. whole page as one line (no boxes), img_height 32
    res_c = reader.recognize(page[40:90, 30:400], None, None, batch_size=1, reformat=False, img_height=32)
    pack_result(out, "rec_c", res_c)
    # output formats on the batched result
    reader.device = torch.device("cpu")
    out["fmt_detail0"] = np.array(reader.recognize(page, h_list, f_list, batch_size=4, reformat=False, detail=0) + [""])
    out["fmt_json"] = np.array(reader.recognize(page, h_list, f_list, batch_size=4, reformat=False, output_format="json") + [""])
    d = reader.recognize(page, h_list, f_list, batch_size=4, reformat=False, output_format="dict")
    out["fmt_dict"] = np.array(json.dumps([{"boxes": np.array(x["boxes"], dtype=np.float64).tolist(), "text": x["text"],
                                            "confident": float(x["confident"])} for x in d], ensure_ascii=False))
    try:  # the reference's merge_to_free indexes past its row list when a slanted box lies below every row
        fm = reader.recognize(page, h_list, f_list, batch_size=4, reformat=False, output_format="free_merge")
        pack_result(out, "fmt_free_merge", fm)
        out["fmt_free_merge_error"] = np.array("")
    except Exception as e:  # noqa: BLE001 - the error type is part of the recorded behaviour
        out["fmt_free_merge_error"] = np.array(type(e).__name__)
    # merge_to_free on an arrangement it handles: slanted boxes that start above / left of the later rows
    level = [r for r in res_b[:len(h_list)]]
    level = sorted(level, key=lambda r: (r[0][0][1] // 60, r[0][0][0]))
    free_items = [([[5, -3], [40, 0], [38, 20], [3, 17]], "free0", 0.5), ([[260, 70], [300, 75], [298, 95], [258, 90]], "free1", 0.25)]
    merged = utils.merge_to_free(level + free_items, [f[0] for f in free_items])
    out["helper_merge_in_order"] = np.array([r[1] for r in level] + [""])
    out["helper_merge_text"] = np.array([r[1] for r in merged] + [""])
    out["helper_merge_level_box"] = np.array([r[0] for r in level], dtype=np.float64).reshape(-1, 4, 2)
    out["helper_merge_level_conf"] = np.array([float(r[2]) for r in level], dtype=np.float64)
    for mode_key, kw in (("para", dict(paragraph=True)), ("para_tight", dict(paragraph=True, x_ths=0.3, y_ths=0.2))):
        para = reader.recognize(page, h_list, f_list, batch_size=4, reformat=False, **kw)
        out[mode_key + "_text"] = np.array([p[1] for p in para] + [""])
        out[mode_key + "_box"] = np.array([p[0] for p in para], dtype=np.float64).reshape(-1, 4, 2)
    # result helpers on their own (pure list logic)
    raw = [list(r) for r in res_b]
    for mode in ("ltr", "rtl"):
        para = utils.get_paragraph(raw, x_ths=1.0, y_ths=0.5, mode=mode)
        out[f"helper_para_{mode}_text"] = np.array([p[1] for p in para] + [""])
        out[f"helper_para_{mode}_box"] = np.array([p[0] for p in para], dtype=np.float64).reshape(-1, 4, 2)
    aug = [res_b, [(b, t + "x", c * 0.5 if i % 2 else c * 1.5) for i, (b, t, c) in enumerate(res_b)]]
    pack_result(out, "helper_best", utils.set_result_with_confidence(aug))

    # D. readtext end to end behind a stand-in detector (fixture maps 0 and 1)
    det = np.load(os.path.join(GOLDEN, "detection.npz"))
    for i in (0, 1):
        text, link = det[f"map{i}_text"], det[f"map{i}_link"]
        y = torch.from_numpy(np.stack([text, link], axis=-1))[None]
        reader.detector = lambda x, y=y: (y, None)
        rgb = synth_rgb(2 * text.shape[0], 2 * text.shape[1], 50 + i)
        out[f"read{i}_rgb"] = rgb
        for bs, dev in ((1, "cpu"), (8, torch.device("cpu"))):
            reader.device = dev
            res = reader.readtext(rgb, batch_size=bs)
            pack_result(out, f"read{i}_bs{bs}", res)
        hl, fl = reader.transform_ua(rgb)
        out[f"read{i}_hlist"] = np.array(hl[0], dtype=np.float64).reshape(-1, 4)
        out[f"read{i}_flist"] = np.array(fl[0], dtype=np.float64).reshape(-1, 4, 2)
    path = os.path.join(GOLDEN, "reader.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), {k: int(v) for k, v in out.items() if k.endswith("_n")})
    print(out["rec_a_text"][:5], out["rec_a_conf"][:5])


if __name__ == "__main__":
    main()
