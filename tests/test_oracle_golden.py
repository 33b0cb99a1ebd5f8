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


def test_oracle_strings_match_reference_fp32_column_int8_column_recorded(synth_sd, charset):
    """tests/golden/quantized_int8.npz (oracle/gen_golden_quant.py): the reference's get_text strings with
    quantize=False - the oracle must reproduce them - and with quantize=True (dynamic int8, the OCR(gpu=False)
    default), kept as a recorded second column."""
    from oracle import recognizer_ref as R
    from oracle.synth import synth_lines

    g = load_golden("quantized_int8")
    imgs = synth_lines(64, 32, 128, seed=int(g["seed"]))
    chars = str(charset["character"])
    res = R.recognizer_predict_ref(synth_sd, R.align_collate_ref(list(imgs), 32, 128), ["[blank]"] + list(chars),
                                   charset["ignore_idx"].tolist(), "greedy", 5)
    want32, want8 = g["fp32_text"][:-1].tolist(), g["int8_text"][:-1].tolist()
    assert [r[0] for r in res] == want32
    np.testing.assert_allclose([r[1] for r in res], g["fp32_conf"], rtol=1e-4, atol=1e-12)
    assert len(want8) == 64 and 0 < sum(a == b for a, b in zip(want8, want32)) < 64  # int8 is a different arithmetic
    # the oracle's int8 leg (bench.py's quantize_true baseline row) reproduces that column; the reference ran batches of
    # 32 and dynamic quantization picks its activation scales per batch
    got8 = []
    for i in (0, 32):
        x = R.align_collate_ref(list(imgs[i:i + 32]), 32, 128)
        got8 += [r[0] for r in R.recognizer_predict_ref(synth_sd, x, ["[blank]"] + list(chars),
                                                        charset["ignore_idx"].tolist(), "greedy", 5, quantize=True)]
    assert got8 == want8
