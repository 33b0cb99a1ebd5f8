"""get_image_list (crop + OpenCV-exact INTER_LINEAR resize; reference utils.py:1375-1418): the oracle and the host
coefficient tables against the reference-generated fixture on CPU, the CUDA kernel against both on the GPU."""
import numpy as np
import pytest

from conftest import load_golden


def _golden_lists(g, mh):
    n = int(g[f"mh{mh}_n"])
    return [(g[f"mh{mh}_box{i}"].tolist(), g[f"mh{mh}_crop{i}"]) for i in range(n)], int(g[f"mh{mh}_max_width"])


@pytest.mark.parametrize("mh", [32, 64])
def test_oracle_get_image_list_matches_reference_fixture(mh):
    from oracle import recognizer_ref as R

    g = load_golden("crops")
    want, want_mw = _golden_lists(g, mh)
    got, mw = R.get_image_list_ref(g["boxes"].tolist(), g["page"], model_height=mh, free_list=g["free"].tolist())
    assert mw == want_mw and len(got) == len(want)
    for (gb, gc), (wb, wc) in zip(got, want):
        assert gb == wb
        assert gc.shape == wc.shape and np.array_equal(gc, wc)


def test_host_tables_match_oracle_axis_coefficients():
    """ocrfa_b200.crops.linear_table (vectorised) == the oracle's scalar restatement of OpenCV's set-up loop."""
    from oracle import recognizer_ref as R
    from ocrfa_b200.crops import linear_table

    rng = np.random.default_rng(3)
    for _ in range(60):
        src, dst = int(rng.integers(2, 700)), int(rng.integers(1, 900))
        for clamp in (True, False):
            idx, w = R._cv_linear_axis(src, dst, clamp)
            t = linear_table(src, dst, clamp)
            assert np.array_equal(t[:, 0], idx)
            a0 = (t[:, 1] & 0xffff).astype(np.int16).astype(np.int64)
            a1 = ((t[:, 1] >> 16) & 0xffff).astype(np.int16).astype(np.int64)
            assert np.array_equal(a0, w[:, 0]) and np.array_equal(a1, w[:, 1])


def test_oracle_resize_matches_installed_opencv():
    """Extra pin where cv2 is installed (it is in this image): random sizes against cv2.resize(interpolation=1)."""
    cv2 = pytest.importorskip("cv2")
    from oracle import recognizer_ref as R

    rng = np.random.default_rng(0)
    for it in range(60):
        sh, sw = int(rng.integers(2, 90)), int(rng.integers(2, 300))
        dh, dw = int(rng.choice([32, 64, int(rng.integers(2, 100))])), int(rng.integers(2, 400))
        if it % 10 == 0:
            dh, dw = max(1, sh // 2), max(1, sw // 2)
            sh, sw = 2 * dh, 2 * dw
        img = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
        assert np.array_equal(R.cv_resize_linear_u8_ref(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=1))


def test_host_perspective_matrices_match_oracle_and_opencv():
    """ocrfa_b200.crops restates getPerspectiveTransform (LU) + the 3x3 inverse: bit-equal to the oracle's restatement
    and, where cv2 is installed, to cv2 itself."""
    from oracle import recognizer_ref as R
    from ocrfa_b200.crops import get_perspective_transform, invert3

    rng = np.random.default_rng(9)
    try:
        import cv2
    except ImportError:
        cv2 = None
    for _ in range(200):
        quad = (rng.uniform(0, 600, (4, 2))).astype(np.float32)
        mw, mh = int(rng.integers(2, 400)), int(rng.integers(2, 90))
        dst = np.array([[0, 0], [mw - 1, 0], [mw - 1, mh - 1], [0, mh - 1]], dtype="float32")
        M = get_perspective_transform(quad, dst)
        assert np.array_equal(M, R._cv_get_perspective_transform_ref(quad, dst))
        assert np.array_equal(invert3(M), R._cv_invert3_ref(M))
        if cv2 is not None:
            assert np.array_equal(M, cv2.getPerspectiveTransform(quad, dst))
            assert np.array_equal(invert3(M), cv2.invert(M)[1])


def test_oracle_warp_matches_installed_opencv():
    cv2 = pytest.importorskip("cv2")
    from oracle import recognizer_ref as R

    rng = np.random.default_rng(4)
    for _ in range(25):
        img = rng.integers(0, 256, (int(rng.integers(40, 160)), int(rng.integers(60, 300))), dtype=np.uint8)
        quad = np.rint(rng.uniform(-10, 1, (4, 2)) + np.array([[20, 10], [img.shape[1] - 30, 14], [img.shape[1] - 25, img.shape[0] - 5],
                                                                [15, img.shape[0] - 12]])).astype(np.float32)
        got = R.four_point_transform_ref(img, quad)
        mh, mw = got.shape
        dst = np.array([[0, 0], [mw - 1, 0], [mw - 1, mh - 1], [0, mh - 1]], dtype="float32")
        want = cv2.warpPerspective(img, cv2.getPerspectiveTransform(quad, dst), (mw, mh))
        assert np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("mh", [32, 64])
def test_gpu_get_image_list_bit_exact_with_reference_fixture(mh):
    import ocrfa_b200 as ob

    g = load_golden("crops")
    want, want_mw = _golden_lists(g, mh)
    got, mw = ob.get_image_list(g["boxes"].tolist(), g["free"].tolist(), g["page"], model_height=mh)
    assert mw == want_mw and len(got) == len(want)
    for (gb, gc), (wb, wc) in zip(got, want):
        assert gb == wb
        assert gc.dtype == np.uint8 and gc.shape == wc.shape and np.array_equal(gc, wc)


@pytest.mark.gpu
def test_gpu_crops_vs_oracle_random_boxes():
    """Seeded random page and boxes (ragged, clipped, tall) at sizes the oracle finishes in seconds."""
    import ocrfa_b200 as ob
    from oracle import recognizer_ref as R

    rng = np.random.default_rng(5)
    page = rng.integers(0, 256, (500, 900), dtype=np.uint8)
    boxes = []
    for _ in range(120):
        w, h = int(rng.integers(4, 880)), int(rng.integers(4, 200))
        x0, y0 = int(rng.integers(-30, 890)), int(rng.integers(-30, 490))
        boxes.append([x0, x0 + w, y0, y0 + h])
    boxes = [b for b in boxes if min(b[1], 900) - max(b[0], 0) > 0 and min(b[3], 500) - max(b[2], 0) > 0]
    free = []
    for _ in range(40):
        quad = rng.uniform(-20, 20, (4, 2))
        cx, cy, bw, bh = rng.uniform(0, 900), rng.uniform(0, 500), rng.uniform(10, 500), rng.uniform(5, 90)
        quad += np.array([[cx, cy], [cx + bw, cy + rng.uniform(-40, 40)], [cx + bw, cy + bh], [cx, cy + bh]])
        free.append(np.rint(quad).astype(int).tolist())
    for mh in (32, 64):
        want, want_mw = R.get_image_list_ref(boxes, page, model_height=mh, free_list=free)
        got, mw = ob.get_image_list(boxes, free, page, model_height=mh)
        assert mw == want_mw and len(got) == len(want)
        for (gb, gc), (wb, wc) in zip(got, want):
            assert gb == wb and np.array_equal(gc, wc)


@pytest.mark.gpu
def test_device_resident_crops_feed_get_text_identically(synth_sd, charset):
    """page -> get_image_list -> get_text with the crops kept on the GPU (DeviceLine) must give exactly the strings
    and confidences of the host-array route (same kernels, no host round trip in between)."""
    import torch

    import ocrfa_b200 as ob

    g = load_golden("crops")
    chars = str(charset["character"])
    conv = ob.CTCLabelConverter(chars)
    model = ob.Model(1, 512, 512, len(conv.character), precision="f16")
    model.load_state_dict({"module." + k: torch.from_numpy(v) for k, v in synth_sd.items()})
    boxes, free = g["boxes"].tolist(), g["free"].tolist()
    host_list, mw = ob.get_image_list(boxes, free, g["page"], model_height=32)
    dev_list, mw2 = ob.get_image_list(boxes, free, g["page"], model_height=32, as_device=True)
    assert mw == mw2 and len(host_list) == len(dev_list)
    for (hb, hc), (db, dc) in zip(host_list, dev_list):
        assert hb == db and np.array_equal(hc, dc.numpy())
    # tall crops (vertical text) are wider than imgW allows after AlignCollate's resize: both routes clip alike
    kw = dict(ignore_char=ob.default_ignore_char(), decoder="greedy", batch_size=16, device="cuda")
    r_host = ob.get_text(chars, 32, int(mw), model, conv, host_list, **kw)
    r_dev = ob.get_text(chars, 32, int(mw), model, conv, dev_list, **kw)
    assert [(b, s) for b, s, _ in r_host] == [(b, s) for b, s, _ in r_dev]
    np.testing.assert_array_equal(np.array([c for _, _, c in r_host]), np.array([c for _, _, c in r_dev]))
