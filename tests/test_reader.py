"""The `OCR` facade (reference ocr.py:42-582) and its result helpers (utils.py:1176-1215, 1442-1509, 1615-1631).
CPU: the list logic against tests/golden/reader.npz (written by oracle/gen_golden_reader.py from the reference's own
OCR class). GPU: OCR.recognize / OCR.readtext in the fp32 parity mode against the reference's results."""
import json
import os

import numpy as np
import pytest

from conftest import load_golden


def _result(g, key):
    n = int(g[key + "_n"])
    return [(g[key + "_box"][i].tolist(), str(g[key + "_text"][i]), float(g[key + "_conf"][i])) for i in range(n)]


def _texts(arr):
    return [str(s) for s in arr[:-1]]


# ------------------------------------------------------------------------------------------------------------------
# CPU: result helpers
# ------------------------------------------------------------------------------------------------------------------
def test_get_paragraph_matches_reference():
    from ocrfa_b200.reader import get_paragraph

    g = load_golden("reader")
    raw = [list(r) for r in _result(g, "rec_b")]
    for mode in ("ltr", "rtl"):
        para = get_paragraph(raw, x_ths=1.0, y_ths=0.5, mode=mode)
        assert [p[1] for p in para] == _texts(g[f"helper_para_{mode}_text"])
        assert np.array_equal(np.array([p[0] for p in para], dtype=np.float64).reshape(-1, 4, 2), g[f"helper_para_{mode}_box"])


def test_set_result_with_confidence_matches_reference():
    from ocrfa_b200.reader import set_result_with_confidence

    g = load_golden("reader")
    res_b = _result(g, "rec_b")
    aug = [res_b, [(b, t + "x", c * 0.5 if i % 2 else c * 1.5) for i, (b, t, c) in enumerate(res_b)]]
    got = set_result_with_confidence(aug)
    want = _result(g, "helper_best")
    assert [r[1] for r in got] == [r[1] for r in want]
    assert [r[2] for r in got] == [r[2] for r in want]


def test_merge_to_free_matches_reference():
    from ocrfa_b200.reader import merge_to_free

    g = load_golden("reader")
    level = [(b.tolist(), t, float(c)) for b, t, c in
             zip(g["helper_merge_level_box"], _texts(g["helper_merge_in_order"]), g["helper_merge_level_conf"])]
    free_items = [([[5, -3], [40, 0], [38, 20], [3, 17]], "free0", 0.5),
                  ([[260, 70], [300, 75], [298, 95], [258, 90]], "free1", 0.25)]
    merged = merge_to_free(level + free_items, [f[0] for f in free_items])
    assert [r[1] for r in merged] == _texts(g["helper_merge_text"])
    assert merge_to_free(level, []) == level
    # the reference indexes past its row list when a slanted box lies below every row; the mirror fails the same way
    if str(g["fmt_free_merge_error"]) == "IndexError":
        res = _result(g, "rec_b")
        with pytest.raises(IndexError):
            merge_to_free(res, g["rec_f_list"].tolist())


def test_reader_refuses_cpu():
    import torch
    import ocrfa_b200 as ob

    with pytest.raises(ob.OcrfaError):
        ob.OCR(gpu=False)
    if not torch.cuda.is_available():
        with pytest.raises(ob.OcrfaError):
            ob.OCR(gpu=True)


# ------------------------------------------------------------------------------------------------------------------
# GPU: the facade end to end (fp32 parity mode -> identical strings)
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def reader(tmp_path_factory, synth_sd):
    import torch
    import ocrfa_b200 as ob

    d = tmp_path_factory.mktemp("models")
    torch.save({"module." + k: torch.from_numpy(v) for k, v in synth_sd.items()}, os.path.join(d, "arabic.pth"))
    torch.save({k: torch.from_numpy(v) for k, v in ob.synth_craft_state_dict(0).items()},
               os.path.join(d, "craft_mlt_25k.pth"))
    with pytest.raises(FileNotFoundError):  # the released files' MD5 sums are checked unless told otherwise
        ob.OCR(model_storage_directory=str(d))
    with pytest.raises(FileNotFoundError):
        ob.OCR(model_storage_directory=str(d / "nothing_here"), verify_md5=False)
    # display_order=False: the fixture was generated with python-bidi stubbed to the identity (oracle/_refimport.py)
    return ob.OCR(model_storage_directory=str(d), verify_md5=False, precision="fp32", detector_precision="f16",
                  display_order=False)


def _assert_results(got, want, conf_rtol=5e-3):
    assert len(got) == len(want)
    for (gb, gt, gc), (wb, wt, wc) in zip(got, want):
        assert np.array_equal(np.array(gb, dtype=np.float64), np.array(wb, dtype=np.float64))
        assert gt == wt
        assert abs(gc - wc) <= conf_rtol * abs(wc) + 1e-12


@pytest.mark.gpu
def test_gpu_recognize_matches_reference(reader):
    g = load_golden("reader")
    page = load_golden("crops")["page"]
    h_list, f_list = g["rec_h_list"].tolist(), g["rec_f_list"].tolist()
    _assert_results(reader.recognize(page, h_list, f_list, batch_size=1, reformat=False), _result(g, "rec_a"))
    res_b = reader.recognize(page, h_list, f_list, batch_size=4, reformat=False)
    _assert_results(res_b, _result(g, "rec_b"))
    _assert_results(reader.recognize(page, h_list[:4], [], decoder="beamsearch", beamWidth=5, batch_size=4, reformat=False),
                    _result(g, "rec_beam"))
    _assert_results(reader.recognize(page, h_list[:4], [], decoder="wordbeamsearch", beamWidth=5, batch_size=4,
                                     reformat=False), _result(g, "rec_word"))
    _assert_results(reader.recognize(page, h_list[:4], [], batch_size=4, reformat=False, allowlist="0123456789ابپتث"),
                    _result(g, "rec_allow"))
    _assert_results(reader.recognize(page, h_list[:4], [], batch_size=4, reformat=False, blocklist="0123456789دزش"),
                    _result(g, "rec_block"))
    _assert_results(reader.recognize(page[40:90, 30:400], None, None, batch_size=1, reformat=False, img_height=32),
                    _result(g, "rec_c"))
    # reformat=True on a grey array gives the same page
    _assert_results(reader.recognize(page, h_list, f_list, batch_size=4), _result(g, "rec_b"))
    # output formats
    assert reader.recognize(page, h_list, f_list, batch_size=4, reformat=False, detail=0) == _texts(g["fmt_detail0"])
    got_json = reader.recognize(page, h_list, f_list, batch_size=4, reformat=False, output_format="json")
    for a, b in zip(got_json, _texts(g["fmt_json"])):
        ja, jb = json.loads(a), json.loads(b)
        assert ja["boxes"] == jb["boxes"] and ja["text"] == jb["text"]
        assert abs(ja["confident"] - jb["confident"]) <= 5e-3 * abs(jb["confident"]) + 1e-12
    got_dict = reader.recognize(page, h_list, f_list, batch_size=4, reformat=False, output_format="dict")
    want_dict = json.loads(str(g["fmt_dict"]))
    assert [d["text"] for d in got_dict] == [d["text"] for d in want_dict]
    assert [np.array(d["boxes"], dtype=np.float64).tolist() for d in got_dict] == [d["boxes"] for d in want_dict]
    for key, kw in (("para", dict(paragraph=True)), ("para_tight", dict(paragraph=True, x_ths=0.3, y_ths=0.2))):
        para = reader.recognize(page, h_list, f_list, batch_size=4, reformat=False, **kw)
        assert [p[1] for p in para] == _texts(g[key + "_text"])
        assert np.array_equal(np.array([p[0] for p in para], dtype=np.float64).reshape(-1, 4, 2), g[key + "_box"])


@pytest.mark.gpu
def test_gpu_rotation_info_matches_reference(reader):
    """rotation_info test-time augmentation: crops rotated on the GPU (= scipy.ndimage.rotate for right angles), every
    variant recognized, the most confident reading kept (ocr.py:446-473)."""
    g = load_golden("reader")
    page = load_golden("crops")["page"]
    h_list, f_list = g["rec_h_list"].tolist(), g["rec_f_list"].tolist()
    got = reader.recognize(page, h_list[:4], f_list[:1], rotation_info=[90, 180, 270], batch_size=4, reformat=False)
    _assert_results(got, _result(g, "rec_rot"))
    got = reader.recognize(page, h_list[4:6], [], rotation_info=[180], batch_size=1, reformat=False, img_height=32)
    _assert_results(got, _result(g, "rec_rot180"))
    with pytest.raises(NotImplementedError):
        reader.recognize(page, h_list[:1], [], rotation_info=[45], reformat=False)


@pytest.mark.gpu
def test_gpu_rot90_equals_scipy_rotate():
    ndimage = pytest.importorskip("scipy.ndimage")
    import torch
    from ocrfa_b200.crops import DeviceLine
    from ocrfa_b200.reader import make_rotated_img_list

    rng = np.random.default_rng(6)
    shapes = [(64, 200), (32, 77), (1, 1), (64, 64), (33, 101), (5, 300)]
    imgs = [rng.integers(0, 256, s, dtype=np.uint8) for s in shapes]
    offs, off = [], 0
    for im in imgs:
        offs.append(off)
        off += (im.size + 3) & ~3
    buf = torch.zeros(off, dtype=torch.uint8, device="cuda")
    for im, o in zip(imgs, offs):
        buf[o:o + im.size] = torch.from_numpy(im.reshape(-1)).cuda()
    lines = [([[0, 0]] * 4, DeviceLine(buf, o, *im.shape)) for im, o in zip(imgs, offs)]
    angles = [90, 180, 270, -90, 360]
    out = make_rotated_img_list(angles, lines)
    assert len(out) == len(imgs) * (1 + len(angles))
    for i, im in enumerate(imgs):
        assert np.array_equal(out[i][1].numpy(), im)
    for a, angle in enumerate(angles):
        for i, im in enumerate(imgs):
            got = out[(a + 1) * len(imgs) + i][1].numpy()
            want = ndimage.rotate(im, angle, reshape=True)
            assert got.shape == want.shape and np.array_equal(got, want), (angle, im.shape)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [0, 1])
def test_gpu_readtext_matches_reference_behind_fixture_detector(reader, case):
    """readtext end to end - grey conversion, detector input, labeling, boxes, grouping, crops, recognition - with the
    CRAFT forward replaced by the fixture's score maps on both sides (the forward has its own test, tests/test_craft.py)."""
    import torch

    g = load_golden("reader")
    det = load_golden("detection")
    y = torch.from_numpy(np.stack([det[f"map{case}_text"], det[f"map{case}_link"]], axis=-1))[None].cuda()
    rgb = g[f"read{case}_rgb"]
    saved = reader.detector
    reader.detector = lambda x: (y, None)
    try:
        hl, fl = reader.transform_ua(rgb)
        assert np.array_equal(np.array(hl[0], dtype=np.float64).reshape(-1, 4), g[f"read{case}_hlist"])
        assert np.array_equal(np.array(fl[0], dtype=np.float64).reshape(-1, 4, 2), g[f"read{case}_flist"])
        for bs in (1, 8):
            _assert_results(reader.readtext(rgb, batch_size=bs), _result(g, f"read{case}_bs{bs}"))
    finally:
        reader.detector = saved


@pytest.mark.gpu
def test_gpu_readtext_with_the_craft_forward(reader):
    """The whole path including the CRAFT network (synthetic weights): thresholds are set from the score maps the
    network actually produces, so the labeling has something to find; results are structurally checked and the boxes
    are compared with the oracle's post-processing of the same GPU score maps."""
    import torch
    from oracle import det_ref as D
    from ocrfa_b200.detection import detector_input

    rng = np.random.default_rng(4)
    rgb = rng.integers(0, 256, (200, 328, 3), dtype=np.uint8)
    x, ratio = detector_input(rgb, 2560, 1.0)
    y, _ = reader.detector(x)
    text, link = y[0, :, :, 0].cpu().numpy(), y[0, :, :, 1].cpu().numpy()
    low, tt, lt = (float(np.quantile(text, q)) for q in (0.80, 0.90, 0.0))
    lt = float(np.quantile(link, 0.97))
    want = D.textbox_from_maps_ref(text, link, ratio, tt, lt, low)
    got = reader.get_textbox(reader.detector, torch.from_numpy(rgb).cuda(), 2560, 1.0, tt, lt, low, False, "cuda")[0]
    assert len(got) == len(want) and len(got) > 0
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    res = reader.readtext(rgb, batch_size=16, text_threshold=tt, low_text=low, link_threshold=lt, min_size=5)
    assert len(res) > 0
    for box, text_, conf in res:
        assert len(box) == 4 and isinstance(text_, str) and 0.0 <= conf <= 1.0
    pages = reader.readtext_batched([rgb, rgb], batch_size=16, text_threshold=tt, low_text=low, link_threshold=lt, min_size=5)
    assert len(pages) == 2 and [r[1] for r in pages[0]] == [r[1] for r in res] == [r[1] for r in pages[1]]


@pytest.mark.gpu
def test_gpu_default_precision_reader_reads_the_same_boxes(reader, tmp_path, synth_sd):
    """The constructor's defaults (f16x3 recognizer and detector: fp32-grade tensor-core arithmetic): the same
    boxes AND the same strings as the fp32 (CUDA-core) reader; display_order=True additionally applies the RTL display
    order (ocr.py:474-478) to every string."""
    import torch
    import ocrfa_b200 as ob
    from ocrfa_b200.bidi import get_display

    torch.save({k: torch.from_numpy(v) for k, v in synth_sd.items()}, os.path.join(tmp_path, "arabic.pth"))
    torch.save({k: torch.from_numpy(v) for k, v in ob.synth_craft_state_dict(0).items()},
               os.path.join(tmp_path, "craft_mlt_25k.pth"))
    fast = ob.OCR(model_storage_directory=str(tmp_path), verify_md5=False)
    assert fast.precision == "f16x3" and fast.detector_precision == "f16x3" and fast.model_lang == "arabic"
    assert fast.display_order and fast.bidi_backend in ("python-bidi", "ocrfa_b200.bidi")
    g = load_golden("reader")
    page = load_golden("crops")["page"]
    h_list, f_list = g["rec_h_list"].tolist(), g["rec_f_list"].tolist()
    got = fast.recognize(page, h_list, f_list, batch_size=16, reformat=False)
    want = reader.recognize(page, h_list, f_list, batch_size=16, reformat=False)  # logical order (display_order=False)
    assert [r[0] for r in got] == [r[0] for r in want]
    if fast.bidi_backend == "ocrfa_b200.bidi":
        assert [r[1] for r in got] == [get_display(r[1]) for r in want]
    fast.display_order = False
    got = fast.recognize(page, h_list, f_list, batch_size=16, reformat=False)
    assert [r[1] for r in got] == [r[1] for r in want]


@pytest.mark.gpu
def test_gpu_bgr2gray_bit_exact_with_opencv():
    cv2 = pytest.importorskip("cv2")
    import torch
    from ocrfa_b200.reader import bgr2gray_device, reformat_input

    rng = np.random.default_rng(2)
    for h, w in ((1, 1), (7, 13), (300, 521), (1080, 1920)):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(bgr2gray_device(torch.from_numpy(img).cuda()).cpu().numpy(), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
    img = rng.integers(0, 256, (40, 60, 4), dtype=np.uint8)
    colour, grey = reformat_input(img)
    want = cv2.cvtColor(img[:, :, :3], cv2.COLOR_RGB2BGR)
    assert np.array_equal(colour.cpu().numpy(), want) and np.array_equal(grey.cpu().numpy(), cv2.cvtColor(want, cv2.COLOR_BGR2GRAY))
    g2 = rng.integers(0, 256, (40, 60), dtype=np.uint8)
    colour, grey = reformat_input(g2)
    assert np.array_equal(grey.cpu().numpy(), g2) and np.array_equal(colour.cpu().numpy(), cv2.cvtColor(g2, cv2.COLOR_GRAY2BGR))


def test_reformat_input_rejects_what_the_reference_rejects():
    """Argument errors surface before any device work (utils.py:1564-1565; URL fetching is not offered)."""
    from ocrfa_b200.reader import make_rotated_img_list, reformat_input

    with pytest.raises(ValueError):
        reformat_input("https://example.org/page.png")
    with pytest.raises(ValueError):
        reformat_input(12345)
    with pytest.raises(ValueError):
        reformat_input("/nonexistent/page.png")
    with pytest.raises(NotImplementedError):
        make_rotated_img_list([30], [])
    assert make_rotated_img_list([90, 180], []) == []
