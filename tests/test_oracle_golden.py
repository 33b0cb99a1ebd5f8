"""Pins the oracle (oracle/recognizer_ref.py) against outputs of the reference itself (tests/golden/, written
by oracle/gen_golden.py from /root/reference). CPU only."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import recognizer_ref as R

FWD = ["c1_32x128", "c2_32x512", "c3_64x256"]


@pytest.mark.parametrize("name", FWD + ["c3_64x2048"])
def test_preprocess_matches_reference(name):
    g = load_golden(name)
    imgs = g["images"]
    x = R.align_collate_ref(list(imgs), imgs.shape[1], imgs.shape[2])
    assert x.dtype == np.float32 and np.array_equal(x, g["x"])


@pytest.mark.parametrize("name", FWD)
def test_forward_logits_match_reference(name, synth_sd):
    g = load_golden(name)
    logits = R.model_forward_ref(synth_sd, g["x"]).numpy()
    # same torch CPU kernels as the reference except the LSTM (hoisted projection): fp32 round-off only
    np.testing.assert_allclose(logits, g["logits"], rtol=1e-3, atol=1e-3)
    assert np.array_equal(logits.argmax(2), g["logits"].argmax(2))


@pytest.mark.parametrize("name", FWD + ["c3_64x2048"])
def test_probs_greedy_conf_match_reference(name, charset):
    g = load_golden(name)
    ign = charset["ignore_idx"].tolist()
    chars = charset["converter_character"].tolist()
    probs = R.probs_ref(g["logits"], ign)
    assert np.array_equal(probs, g["probs"])
    assert R.greedy_decode_ref(probs, chars) == g["str_greedy_5"].tolist()
    np.testing.assert_allclose(np.array(R.confidence_ref(probs), dtype=np.float64), g["conf_greedy_5"], rtol=1e-6)


@pytest.mark.parametrize("name", FWD)
@pytest.mark.parametrize("bw", [5, 8])
def test_beam_matches_reference_on_model_probs(name, bw, charset):
    g = load_golden(name)
    chars = charset["converter_character"].tolist()
    assert R.beam_decode_ref(g["probs"], chars, bw) == g[f"str_beamsearch_{bw}"].tolist()


@pytest.mark.parametrize("name", FWD)
def test_wordbeam_matches_reference_on_model_probs(name, charset):
    g = load_golden(name)
    chars = charset["converter_character"].tolist()
    assert R.wordbeam_decode_ref(g["probs"], chars, 8) == g["str_wordbeamsearch_8"].tolist()


def test_decode_edge_cases_match_reference(charset):
    d = load_golden("decode_cases")
    chars = charset["converter_character"].tolist()
    mat = d["mat"]
    assert R.greedy_decode_ref(mat, chars) == d["greedy"].tolist()
    for bw in (1, 2, 5, 8, 25):
        assert R.beam_decode_ref(mat, chars, bw) == d[f"beam_{bw}"].tolist(), bw
    for bw in (5, 8):
        assert R.wordbeam_decode_ref(mat, chars, bw) == d[f"wordbeam_{bw}"].tolist(), bw
    assert R.greedy_decode_ref(d["mat_t1"], chars) == d["greedy_t1"].tolist()
    assert R.beam_decode_ref(d["mat_t1"], chars, 8) == d["beam_t1_8"].tolist()


def test_ragged_collate_and_contrast_match_reference():
    g = load_golden("ragged_get_text")
    imgs = [g[f"img{i}"] for i in range(4)]
    h, w = int(g["img_h"]), int(g["img_w"])
    assert np.array_equal(R.align_collate_ref(imgs, h, w), g["x"])
    assert np.array_equal(R.align_collate_ref(imgs, h, w, adjust_contrast=0.5), g["x_contrast"])


def test_ignore_idx_matches_reference(charset):
    ign = R.ignore_idx_ref(str(charset["character"]), str(charset["ignore_char"]))
    assert sorted(ign) == charset["ignore_idx"].tolist()
    assert len(ign) == 69 and 0 not in ign
