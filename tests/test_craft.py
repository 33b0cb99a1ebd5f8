"""CRAFT detector forward (reference: utils.py:48-98 OCRFA over blocks.py:27-82 Vgg16BN): the oracle against the
fixture produced by the reference's own module on CPU, the CUDA path against the same fixture on the GPU."""
import numpy as np
import pytest
import torch

from conftest import load_golden


@pytest.fixture(scope="module")
def craft_sd():
    from ocrfa_b200.craft import synth_craft_state_dict

    return synth_craft_state_dict(0)


@pytest.mark.parametrize("case", ["a", "b"])
def test_oracle_craft_forward_matches_reference_fixture(craft_sd, case):
    from oracle import craft_ref

    g = load_golden("craft_64x96")
    y, feat = craft_ref.craft_forward_ref(craft_sd, g[f"x_{case}"])
    np.testing.assert_allclose(y.numpy(), g[f"y_{case}"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(feat.numpy(), g[f"feature_{case}"], rtol=1e-4, atol=1e-5)


def test_craft_state_dict_has_the_reference_keys(craft_sd):
    """25 convs + 2 fc + 16 BatchNorms: every tensor the reference module owns (minus num_batches_tracked)."""
    convs = [k for k in craft_sd if k.endswith(".weight") and craft_sd[k].ndim == 4]
    assert len(convs) == 27 and craft_sd["basenet.slice5.1.weight"].shape == (1024, 512, 3, 3)
    assert craft_sd["upconv1.conv.0.weight"].shape == (512, 1536, 1, 1) and craft_sd["conv_cls.8.weight"].shape == (2, 16, 1, 1)


def test_craft_rejects_cpu_and_bad_geometry():
    import ocrfa_b200 as ob
    from ocrfa_b200.craft import CraftNet

    with pytest.raises(ob.OcrfaError):
        CraftNet()(torch.zeros(1, 3, 64, 64))
    with pytest.raises(ValueError):
        CraftNet(precision="fp32")


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("f16", 4e-3), ("bf16", 3e-2)])
@pytest.mark.parametrize("case", ["a", "b"])
def test_gpu_craft_forward_vs_reference_fixture(craft_sd, prec, tol, case):
    """16-bit tensor-core forward against the reference module's fp32 output: max error relative to the range of the
    score maps / feature map (fp16: 4e-3, bf16: 3e-2 - 27 conv layers of 16-bit activations)."""
    from ocrfa_b200.craft import CraftNet

    g = load_golden("craft_64x96")
    net = CraftNet(precision=prec)
    net.load_state_dict({"module." + k: torch.from_numpy(v) for k, v in craft_sd.items()})
    y, feat = net(torch.from_numpy(g[f"x_{case}"]).cuda())
    y, feat = y.cpu().numpy(), feat.cpu().numpy()
    assert y.shape == g[f"y_{case}"].shape and feat.shape == g[f"feature_{case}"].shape
    assert np.isfinite(y).all() and np.isfinite(feat).all()
    ey = np.abs(y - g[f"y_{case}"]).max() / np.abs(g[f"y_{case}"]).max()
    ef = np.abs(feat - g[f"feature_{case}"]).max() / np.abs(g[f"feature_{case}"]).max()
    print(f"craft {prec} {case}: score-map error {ey:.3g}, feature error {ef:.3g} (relative to range)")
    assert ey <= tol and ef <= tol


@pytest.mark.gpu
def test_gpu_craft_larger_image_vs_oracle(craft_sd):
    """A 160x224 image (maps that tile into the halo-staged kernel at full resolution, per-tap kernel at 1/16) against
    the oracle, fp16 mode; batch independence: the second image alone gives the same maps bit for bit."""
    from oracle import craft_ref
    from ocrfa_b200.craft import CraftNet

    rng = np.random.default_rng(11)
    x = rng.standard_normal((2, 3, 160, 224)).astype(np.float32)
    want_y, want_f = craft_ref.craft_forward_ref(craft_sd, x)
    net = CraftNet(precision="f16")
    net.load_state_dict({k: torch.from_numpy(v) for k, v in craft_sd.items()})
    y, feat = net(torch.from_numpy(x).cuda())
    ey = np.abs(y.cpu().numpy() - want_y.numpy()).max() / np.abs(want_y.numpy()).max()
    ef = np.abs(feat.cpu().numpy() - want_f.numpy()).max() / np.abs(want_f.numpy()).max()
    print(f"craft f16 160x224: score-map error {ey:.3g}, feature error {ef:.3g}")
    assert ey <= 4e-3 and ef <= 4e-3
    y1, _ = net(torch.from_numpy(x[1:2]).cuda())
    assert torch.equal(y1[0], y[1])


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["a", "b"])
def test_gpu_craft_f16x3_forward_is_fp32_grade(craft_sd, case):
    """The fp32-grade mode (fp16 hi/lo split operands on igemm_x3_kernel, split glue kernels) against the reference
    module's fp32 output: 1e-5 of the range of the score maps / feature map (f16: 1.5e-3, bf16: 1e-2)."""
    from ocrfa_b200.craft import CraftNet

    g = load_golden("craft_64x96")
    net = CraftNet(precision="f16x3")
    net.load_state_dict({k: torch.from_numpy(v) for k, v in craft_sd.items()})
    y, feat = net(torch.from_numpy(g[f"x_{case}"]).cuda())
    y, feat = y.cpu().numpy(), feat.cpu().numpy()
    assert y.shape == g[f"y_{case}"].shape and feat.shape == g[f"feature_{case}"].shape
    ey = np.abs(y - g[f"y_{case}"]).max() / np.abs(g[f"y_{case}"]).max()
    ef = np.abs(feat - g[f"feature_{case}"]).max() / np.abs(g[f"feature_{case}"]).max()
    print(f"craft f16x3 {case}: score-map error {ey:.3g}, feature error {ef:.3g} (relative to range)")
    assert ey <= 1e-5 and ef <= 1e-5


@pytest.mark.gpu
def test_gpu_real_forward_boxes_equal_the_reference_boxes(craft_sd):
    """SURVEY section 8f row 4, end to end with the REAL forward: `get_textbox` (resize + normalise on the GPU, CRAFT
    forward in f16x3, thresholds / connected components / dilation on the GPU, minAreaRect on the row extents, coordinate
    adjustment) returns exactly the boxes the reference's `get_textbox` returns with its own fp32 forward of the same
    weights on the same pages (tests/golden/craft_boxes.npz, oracle/gen_golden_craft_boxes.py; reference:
    detectors.py:24-115, utils.py:48-98, 106-175). The score maps agree to 1e-5 of their range and the thresholds sit
    in gaps of the reference's maps at least 4e-5 of the range wide, so no pixel can change sides."""
    import ocrfa_b200 as ob
    from ocrfa_b200.craft import CraftNet

    g = load_golden("craft_boxes")
    net = CraftNet(precision="f16x3")
    net.load_state_dict({k: torch.from_numpy(v) for k, v in craft_sd.items()})
    for i, (h, w, canvas, mag, _) in enumerate(g["cases"].tolist()):
        tt, lt, low = g[f"c{i}_thr"].tolist()
        res = ob.get_textbox(net, g[f"c{i}_page"], int(canvas), float(mag), tt, lt, low, False, "cuda")
        got = np.array(res[0], dtype=np.int32).reshape(-1, 8)
        want = g[f"c{i}_boxes"]
        print(f"case {i}: {len(got)} boxes (reference {len(want)})")
        assert got.shape == want.shape and np.array_equal(got, want)
