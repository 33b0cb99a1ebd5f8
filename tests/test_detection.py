"""Detector pre-/post-processing around the CRAFT forward (reference utils.py:106-175, 329-349, 479-524, 1235-1348;
detectors.py:24-115). CPU: the oracle and the host logic against tests/golden/detection.npz (written by
oracle/gen_golden_det.py from the reference's own functions). GPU: the kernels behind ocrfa_det_preprocess_u8 /
ocrfa_det_components against the oracle (bit-exact) and the fixture."""
import numpy as np
import pytest

from conftest import load_golden


def _boxes(g, prefix):
    return [g[f"{prefix}_{i}"] for i in range(int(g[f"{prefix}_n"]))]


def _map_cases(g):
    for i in range(len(g["map_cases"])):
        for j, (tt, lt, low) in enumerate(g["thr_sets"]):
            yield i, j, float(tt), float(lt), float(low)


def _assert_boxes_equal(got, want):
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert np.asarray(a).dtype == np.asarray(b).dtype and np.array_equal(a, b)


# ------------------------------------------------------------------------------------------------------------------
# CPU: oracle vs the reference fixture
# ------------------------------------------------------------------------------------------------------------------
def test_oracle_detector_input_matches_reference_fixture():
    from oracle import det_ref as D

    g = load_golden("detection")
    for i, (h, w, canvas, mag100) in enumerate(g["pre_cases"]):
        x, ratio = D.detector_input_ref(g[f"pre{i}_img"], int(canvas), mag100 / 100.0)
        assert ratio == float(g[f"pre{i}_ratio"])
        assert x.shape == g[f"pre{i}_x"].shape and np.array_equal(x, g[f"pre{i}_x"])


def test_oracle_get_det_boxes_matches_reference_fixture():
    from oracle import det_ref as D

    g = load_golden("detection")
    for i, j, tt, lt, low in _map_cases(g):
        det, _, mapper = D.get_det_boxes_core_ref(g[f"map{i}_text"], g[f"map{i}_link"], tt, lt, low)
        _assert_boxes_equal(det, _boxes(g, f"map{i}_thr{j}_box"))
        assert mapper == g[f"map{i}_thr{j}_mapper"].tolist()
        polys = D.textbox_from_maps_ref(g[f"map{i}_text"], g[f"map{i}_link"], 1.0, tt, lt, low)
        _assert_boxes_equal(polys, _boxes(g, f"map{i}_thr{j}_poly"))


# ------------------------------------------------------------------------------------------------------------------
# CPU: the product's host logic (no kernels involved)
# ------------------------------------------------------------------------------------------------------------------
def test_host_boxes_from_row_extents_match_reference_boxes():
    """cv2.minAreaRect over the dilated row extents == the reference's call over the full dilated mask."""
    from oracle import det_ref as D
    from ocrfa_b200.detection import DET_COMP, boxes_from_components

    g = load_golden("detection")
    for i, j, tt, lt, low in _map_cases(g):
        _, _, comps, ext = D.components_ref(g[f"map{i}_text"], g[f"map{i}_link"], tt, lt, low)
        arr = np.zeros(len(comps), dtype=DET_COMP)
        for n, c in enumerate(comps):
            for k, v in c.items():
                arr[n][k] = v
        det, mapper = boxes_from_components(arr, ext)
        _assert_boxes_equal(det, _boxes(g, f"map{i}_thr{j}_box"))
        assert mapper == g[f"map{i}_thr{j}_mapper"].tolist()


def test_host_boxes_from_row_extents_equal_full_mask_boxes_on_random_maps():
    """The claim the GPU post-processing rests on, beyond the fixture: for 30 random CRAFT-like map pairs (rotated words,
    border contact, tiny and weak blobs) the boxes rebuilt from dilated row extents equal, bit for bit, the oracle's
    get_det_boxes_core (minAreaRect over the full dilated mask)."""
    from oracle import det_ref as D
    from ocrfa_b200.detection import DET_COMP, boxes_from_components

    rng = np.random.default_rng(17)
    n_boxes = 0
    for it in range(30):
        h, w = int(rng.integers(24, 140)), int(rng.integers(24, 220))
        text, link = D.synth_score_maps(h, w, 1000 + it, int(rng.integers(2, 30)))
        tt, lt, low = ((0.7, 0.4, 0.4), (0.5, 0.3, 0.25), (0.9, 0.6, 0.2))[it % 3]
        want, _, want_mapper = D.get_det_boxes_core_ref(text, link, tt, lt, low)
        _, _, comps, ext = D.components_ref(text, link, tt, lt, low)
        arr = np.zeros(len(comps), dtype=DET_COMP)
        for n, c in enumerate(comps):
            for k, v in c.items():
                arr[n][k] = v
        got, mapper = boxes_from_components(arr, ext)
        _assert_boxes_equal(got, want)
        assert mapper == want_mapper
        n_boxes += len(got)
    assert n_boxes > 80


def test_host_resize_geometry_and_coordinates():
    from ocrfa_b200.detection import adjust_result_coordinates, resize_geometry

    g = load_golden("detection")
    for i, (h, w, canvas, mag100) in enumerate(g["pre_cases"]):
        th, tw, ch, cw, ratio = resize_geometry(int(h), int(w), int(canvas), mag100 / 100.0)
        assert (3, ch, cw) == g[f"pre{i}_x"].shape and ratio == float(g[f"pre{i}_ratio"])
        assert (int(cw / 2), int(ch / 2)) == tuple(g[f"pre{i}_heat"].tolist())
    # scaled case: a 3h x 3w page through canvas_size = 2w -> boxes are mapped back with 1 / ratio
    for j in range(len(g["thr_sets"])):
        h, w = (int(v) for v in g["map_cases"][0][:2])
        ratio = resize_geometry(3 * h, 3 * w, 2 * w, 1.0)[4]
        boxes = adjust_result_coordinates(_boxes(g, f"map0_thr{j}_box"), 1 / ratio, 1 / ratio)
        flat = [np.array(b).astype(np.int32).reshape(-1) for b in boxes]
        _assert_boxes_equal(flat, _boxes(g, f"map0_thr{j}_poly_scaled"))
    assert adjust_result_coordinates([], 2.0, 2.0) == []


def test_host_group_text_box_matches_reference_fixture():
    from ocrfa_b200.detection import group_text_box
    from ocrfa_b200.reader import diff

    g = load_golden("detection")
    n_h = n_f = 0
    for i, j, *_ in _map_cases(g):
        polys = _boxes(g, f"map{i}_thr{j}_poly")
        hl, fl = group_text_box(polys, 0.1, 0.5, 0.5, 0.5, 0.1, True)
        hl = [b for b in hl if max(b[1] - b[0], b[3] - b[2]) > 20]
        fl = [b for b in fl if max(diff([c[0] for c in b]), diff([c[1] for c in b])) > 20]
        assert np.array_equal(np.array(hl, dtype=np.float64).reshape(-1, 4), g[f"map{i}_thr{j}_hlist"])
        assert np.array_equal(np.array(fl, dtype=np.float64).reshape(-1, 4, 2), g[f"map{i}_thr{j}_flist"])
        n_h, n_f = n_h + len(hl), n_f + len(fl)
    assert n_h > 10 and n_f > 3  # the fixture exercises both branches


# ------------------------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_detector_input_bit_exact():
    import torch
    from oracle import det_ref as D
    from ocrfa_b200.detection import detector_input

    g = load_golden("detection")
    for i, (h, w, canvas, mag100) in enumerate(g["pre_cases"]):
        x, ratio = detector_input(g[f"pre{i}_img"], int(canvas), mag100 / 100.0)
        assert ratio == float(g[f"pre{i}_ratio"])
        assert np.array_equal(x[0].cpu().numpy(), g[f"pre{i}_x"]), f"case {i}"
    # a page-sized image, shrink and copy, against the oracle
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (1100, 1700, 3), dtype=np.uint8)
    for canvas in (2560, 1280, 850):
        x, _ = detector_input(torch.from_numpy(img).cuda(), canvas, 1.0)
        ref, _ = D.detector_input_ref(img, canvas, 1.0)
        assert np.array_equal(x[0].cpu().numpy(), ref), f"canvas {canvas}"


def _check_components(text, link, tt, lt, low, strided=False):
    import torch
    from oracle import det_ref as D
    from ocrfa_b200.detection import det_components

    if strided:  # the CRAFT output layout: [H, W, 2], maps read in place
        y = torch.from_numpy(np.stack([text, link], axis=-1)).cuda()
        t, l = y[:, :, 0], y[:, :, 1]
    else:
        t, l = torch.from_numpy(text).cuda(), torch.from_numpy(link).cuda()
    n_labels, comps, ext, labels = det_components(t, l, tt, lt, low, want_labels=True)
    r_labels, r_lab, r_comps, r_ext = D.components_ref(text, link, tt, lt, low)
    assert n_labels == r_labels
    assert np.array_equal(labels.cpu().numpy(), r_lab)  # cv2's label numbers
    assert len(comps) == len(r_comps)
    for c, r in zip(comps, r_comps):
        for k, v in r.items():
            assert c[k] == v, (k, c, r)
    assert np.array_equal(ext, r_ext)
    return len(comps)


@pytest.mark.gpu
def test_gpu_components_match_oracle_on_fixture_maps():
    g = load_golden("detection")
    total = 0
    for i, j, tt, lt, low in _map_cases(g):
        total += _check_components(g[f"map{i}_text"], g[f"map{i}_link"], tt, lt, low, strided=bool(j))
    assert total > 50


@pytest.mark.gpu
def test_gpu_get_det_boxes_and_textbox_match_reference_fixture():
    import torch
    from ocrfa_b200.detection import get_det_boxes, get_textbox

    g = load_golden("detection")
    rng = np.random.default_rng(1)
    for i, j, tt, lt, low in _map_cases(g):
        text, link = g[f"map{i}_text"], g[f"map{i}_link"]
        boxes, polys, mapper = get_det_boxes(text, link, tt, lt, low)
        _assert_boxes_equal(boxes, _boxes(g, f"map{i}_thr{j}_box"))
        assert polys == [None] * len(boxes) and mapper == g[f"map{i}_thr{j}_mapper"].tolist()
        y = torch.from_numpy(np.stack([text, link], axis=-1))[None].cuda()
        page = rng.integers(0, 256, (2 * text.shape[0], 2 * text.shape[1], 3), dtype=np.uint8)
        res = get_textbox(lambda x, y=y: (y, None), page, 2560, 1.0, tt, lt, low, False, "cuda")
        assert len(res) == 1
        _assert_boxes_equal(res[0], _boxes(g, f"map{i}_thr{j}_poly"))
        if i == 0:
            big = rng.integers(0, 256, (3 * text.shape[0], 3 * text.shape[1], 3), dtype=np.uint8)
            res = get_textbox(lambda x, y=y: (y, None), big, 2 * text.shape[1], 1.0, tt, lt, low, False, "cuda")
            _assert_boxes_equal(res[0], _boxes(g, f"map{i}_thr{j}_poly_scaled"))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,seed,words", [((368, 640), 21, 160), ((1280, 1280), 22, 500), ((33, 1000), 23, 40)])
def test_gpu_components_page_sized_maps(shape, seed, words):
    """CRAFT's map size for a 736 x 1280 page, the largest map canvas_size=2560 produces, and a ragged strip."""
    from oracle import det_ref as D

    text, link = D.synth_score_maps(shape[0], shape[1], seed, words)
    assert _check_components(text, link, 0.7, 0.4, 0.4, strided=True) > 10


@pytest.mark.gpu
def test_gpu_components_edge_cases():
    """Empty map, everything one component, a 4-connected checkerboard (the label-capacity worst case), components
    that touch every border, single-pixel rows, diagonal neighbours (not connected under 4-connectivity)."""
    rng = np.random.default_rng(9)
    h, w = 48, 80
    zero = np.zeros((h, w), np.float32)
    assert _check_components(zero, zero, 0.7, 0.4, 0.4) == 0
    one = np.ones((h, w), np.float32)
    assert _check_components(one, zero, 0.7, 0.4, 0.4) == 1
    checker = ((np.add.outer(np.arange(h), np.arange(w)) % 2) == 0).astype(np.float32)
    assert _check_components(checker, zero, 0.7, 0.4, 0.4) == 0  # 1920 one-pixel components, all below the size filter
    noise = (rng.random((h, w)) > 0.45).astype(np.float32) * rng.uniform(0.41, 1.0, (h, w)).astype(np.float32)
    linkn = (rng.random((h, w)) > 0.7).astype(np.float32)
    assert _check_components(noise, linkn, 0.7, 0.4, 0.4) >= 1
    diag = np.zeros((h, w), np.float32)
    for k in range(40):
        diag[k, k:k + 2] = 0.9  # a staircase: 4-connected through the overlapping columns
        diag[k, 60 - k if 60 - k != k else 0] = 0.9  # a pure diagonal: isolated pixels
    assert _check_components(diag, zero, 0.7, 0.4, 0.4) == 1
    border = np.zeros((h, w), np.float32)
    border[0:3, :] = border[-2:, :] = 0.95
    border[:, 0:2] = border[:, -4:] = 0.95
    assert _check_components(border, zero, 0.7, 0.4, 0.4) == 1
    # link-only pixels are removed from the mask but still count for the label, its area and its bounding box
    t2, l2 = np.zeros((h, w), np.float32), np.zeros((h, w), np.float32)
    t2[10:16, 10:30] = 0.9
    l2[12:14, 30:50] = 0.9
    t2[10:16, 50:70] = 0.8
    assert _check_components(t2, l2, 0.7, 0.4, 0.4) == 1


def test_labeling_rule_reproduces_opencv_label_numbers():
    """The numbering rule the CUDA labeling implements (det_kernels.cuh): union-find over left / upper neighbours with
    the smaller index as root, so that a component's root is its first pixel in raster order; labels are the ranks of
    the roots. Emulated here in NumPy on random masks (sparse, dense, stripes, checkerboards) and compared with
    cv2.connectedComponentsWithStats(connectivity=4), whose numbers the reference's `mapper` / `labels` carry."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(23)

    def label_by_rule(mask):
        h, w = mask.shape
        parent = np.where(mask.reshape(-1) != 0, np.arange(h * w), -1)

        def find(a):
            while parent[a] != a:
                a = parent[a]
            return a

        for p in np.flatnonzero(mask.reshape(-1)):
            y, x = divmod(int(p), w)
            for q in ((p - 1) if x > 0 and mask[y, x - 1] else -1, (p - w) if y > 0 and mask[y - 1, x] else -1):
                if q >= 0:
                    a, b = find(int(p)), find(int(q))
                    if a != b:
                        parent[max(a, b)] = min(a, b)
        roots = np.array([find(int(p)) if parent[p] >= 0 else -1 for p in range(h * w)])
        is_root = roots == np.arange(h * w)
        rank = np.cumsum(is_root)  # 1-based rank of every root in raster order
        return np.where(roots >= 0, rank[np.maximum(roots, 0)], 0).reshape(h, w), int(is_root.sum()) + 1

    masks = [(rng.random((int(rng.integers(3, 40)), int(rng.integers(3, 60)))) < p).astype(np.uint8)
             for p in (0.05, 0.3, 0.5, 0.5, 0.7, 0.95) for _ in range(4)]
    masks.append((np.add.outer(np.arange(9), np.arange(14)) % 2).astype(np.uint8))
    masks.append(np.tile(np.array([[1], [0]], np.uint8), (6, 17)))
    masks.append(np.zeros((5, 7), np.uint8))
    masks.append(np.ones((5, 7), np.uint8))
    u_shape = np.zeros((8, 9), np.uint8)  # two arms that meet only at the bottom: provisional labels merge late
    u_shape[:, 1] = u_shape[:, 7] = u_shape[7, 1:8] = 1
    masks.append(u_shape)
    for m in masks:
        n_ref, lab_ref = cv2.connectedComponents(m, connectivity=4)
        lab, n = label_by_rule(m)
        assert n == n_ref and np.array_equal(lab, lab_ref)
