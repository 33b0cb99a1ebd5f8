"""`CTCLabelConverter` built with language separators and dictionary files (utils.py:1058-1081) - options the default
Farsi recognizer leaves empty. tests/golden/converter_options.npz holds the outputs of the reference's own converter
(oracle/gen_golden_converter.py). CPU: the oracle restatement and the drop-in's host logic (constructor, word
segmentation) against it. GPU: the drop-in's three decoders through the C ABI (collapse masks, span search, the candidate
list of `ocrfa_ctc_beam_candidates`) - identical strings."""
import os

import numpy as np
import pytest

from conftest import load_golden
from oracle import recognizer_ref as R

SEPS = "‹›«»"
LETTERS = "abcdefghijkl"
CHARSET_SEP = SEPS + " " + LETTERS
CHARSET_PLAIN = " " + LETTERS
SEPARATORS = {"th": [SEPS[0], SEPS[1]], "en": [SEPS[2], SEPS[3]]}
LANG = {0: "", 1: "th", 2: "en"}


def _strs(a):
    return a[:-1].tolist()


def _write(tmp_path, name, words):
    path = os.path.join(tmp_path, name)
    with open(path, "w", encoding="utf-8") as f:
        f.write("\n".join(words) + ("\n" if words else ""))
    return path


# ------------------------------------------------------------------------------------------------------------------
# CPU: oracle and host logic against the reference's outputs
# ------------------------------------------------------------------------------------------------------------------
def test_oracle_decoders_with_separators_match_reference():
    g = load_golden("converter_options")
    chars = ["[blank]"] + list(CHARSET_SEP)
    ign = R.converter_ignore_idx_ref(SEPARATORS)
    assert ign == [0, 1, 2, 3, 4]
    got, o = [], 0
    for n in g["greedy_len"].tolist():
        got.append("".join(chars[k] for k in R.greedy_labels_ref(g["greedy_idx"][o:o + n], ign)))
        o += n
    assert got == _strs(g["greedy_text"])
    probs = g["sep_probs"]
    assert R.beam_decode_ref(probs, chars, 5, ign) == _strs(g["sep_beam5"])
    dicts = {"th": _strs(g["dict_th"]), "en": _strs(g["dict_en"])}
    assert R.wordbeam_decode_full_ref(probs, chars, 5, ign, SEPARATORS, dicts) == _strs(g["sep_wordbeam5_dict"])
    assert R.wordbeam_decode_full_ref(probs, chars, 5, ign, SEPARATORS, {"th": [], "en": []}) == _strs(g["sep_wordbeam5_nodict"])
    assert _strs(g["sep_wordbeam5_dict"]) != _strs(g["sep_wordbeam5_nodict"])  # the lookup matters in the fixture


@pytest.mark.parametrize("width", [5, 25])
def test_oracle_dictionary_word_search_matches_reference(width):
    g = load_golden("converter_options")
    chars = ["[blank]"] + list(CHARSET_PLAIN)
    probs = g["plain_probs"]
    words = _strs(g[f"dict_plain{width}"])
    assert R.wordbeam_decode_full_ref(probs, chars, width, [0], {}, words) == _strs(g[f"plain_wordbeam{width}_dict"])
    assert R.wordbeam_decode_full_ref(probs, chars, width, [0], {}, []) == _strs(g[f"plain_wordbeam{width}_nodict"])
    assert R.wordbeam_decode_ref(probs, chars, width) == _strs(g[f"plain_wordbeam{width}_nodict"])


def test_word_segmentation_oracle_and_dropin_match_reference():
    from ocrfa_b200.converter import word_segmentation

    g = load_golden("converter_options")
    rows, want = g["seg_rows"], g["seg_result"]
    for fn in (R.word_segmentation_ref, word_segmentation):
        got = [[i, lang, t0, t1] for i in range(rows.shape[0]) for lang, (t0, t1) in fn(rows[i])]
        assert [[i, LANG[c], a, b] for i, c, a, b in want.tolist()] == got, fn.__module__
    assert word_segmentation(np.zeros(0, dtype=np.int64)) == []


def test_converter_constructor_options(tmp_path):
    """utils.py:1044-1081: ignore_idx counts the separator characters; without separators the dictionaries are one flat
    list and unreadable files are skipped; with separators they are kept per language and a missing file raises."""
    import ocrfa_b200 as ob

    conv = ob.CTCLabelConverter(CHARSET_SEP, SEPARATORS, {})
    assert conv.ignore_idx == [0, 1, 2, 3, 4] and conv.dict_list == {} and conv.separator_list == SEPARATORS
    assert conv.character[0] == "[blank]" and conv.character[1:5] == list(SEPS) and conv.dict[" "] == 5
    a = _write(tmp_path, "a.txt", ["abc", "de"])
    b = _write(tmp_path, "b.txt", ["fg"])
    flat = ob.CTCLabelConverter(CHARSET_PLAIN, {}, {"x": a, "gone": os.path.join(tmp_path, "nope.txt"), "y": b})
    assert flat.ignore_idx == [0] and flat.dict_list == ["abc", "de", "fg"]
    by_lang = ob.CTCLabelConverter(CHARSET_SEP, SEPARATORS, {"th": a, "en": b})
    assert by_lang.dict_list == {"th": ["abc", "de"], "en": ["fg"]}
    with pytest.raises(FileNotFoundError):
        ob.CTCLabelConverter(CHARSET_SEP, SEPARATORS, {"th": os.path.join(tmp_path, "nope.txt")})
    with pytest.raises(ValueError):
        ob.CTCLabelConverter("ab", {"th": ["a", "b"], "en": ["c", "d"]}, {})
    default = ob.CTCLabelConverter(ob.default_character())
    assert default.ignore_idx == [0] and default.dict_list == [] and default._collapse_mask("cpu") is None
    assert conv._collapse_mask("cpu").tolist() == [1, 1, 1, 1, 1] + [0] * 13


# ------------------------------------------------------------------------------------------------------------------
# GPU: the drop-in converter through the C ABI
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_decoders_with_separators_match_reference(tmp_path):
    import torch

    import ocrfa_b200 as ob

    g = load_golden("converter_options")
    conv = ob.CTCLabelConverter(CHARSET_SEP, SEPARATORS, {})
    assert conv.decode_greedy(g["greedy_idx"], torch.IntTensor(g["greedy_len"])) == _strs(g["greedy_text"])
    probs = g["sep_probs"]
    assert conv.decode_beamsearch(probs, beamWidth=5) == _strs(g["sep_beam5"])
    paths = {lang: _write(tmp_path, lang + ".txt", _strs(g["dict_" + lang])) for lang in ("th", "en")}
    conv_d = ob.CTCLabelConverter(CHARSET_SEP, SEPARATORS, paths)
    assert conv_d.decode_wordbeamsearch(probs, beamWidth=5) == _strs(g["sep_wordbeam5_dict"])
    empty = {lang: _write(tmp_path, lang + "_empty.txt", []) for lang in ("th", "en")}
    conv_e = ob.CTCLabelConverter(CHARSET_SEP, SEPARATORS, empty)
    assert conv_e.decode_wordbeamsearch(probs, beamWidth=5) == _strs(g["sep_wordbeam5_nodict"])
    with pytest.raises(KeyError):  # a language-tagged segment without that language's dictionary, as in the reference
        conv.decode_wordbeamsearch(probs, beamWidth=5)


@pytest.mark.gpu
@pytest.mark.parametrize("width", [5, 25])
def test_gpu_dictionary_word_search_matches_reference(tmp_path, width):
    import ocrfa_b200 as ob

    g = load_golden("converter_options")
    probs = g["plain_probs"]
    path = _write(tmp_path, "plain.txt", _strs(g[f"dict_plain{width}"]))
    conv = ob.CTCLabelConverter(CHARSET_PLAIN, {}, {"fa": path})
    assert conv.decode_wordbeamsearch(probs, beamWidth=width) == _strs(g[f"plain_wordbeam{width}_dict"])
    plain = ob.CTCLabelConverter(CHARSET_PLAIN, {}, {})
    assert plain.decode_wordbeamsearch(probs, beamWidth=width) == _strs(g[f"plain_wordbeam{width}_nodict"])


@pytest.mark.gpu
def test_gpu_candidate_list_against_the_oracle():
    """ocrfa_ctc_beam_candidates on whole items (no spans): ranked, collapsed candidates equal the oracle's; its first
    candidate equals ocrfa_ctc_beam's result; an n_cand below the beam width is a prefix of the ranking."""
    import ctypes as C

    import torch

    from ocrfa_b200 import _lib

    g = load_golden("converter_options")
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    probs = np.ascontiguousarray(g["plain_probs"][:6, :24])
    B, T, Cn = probs.shape
    m = torch.from_numpy(probs).to(dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for bw, K in ((5, 20), (25, 20), (8, 3), (1, 20)):
        ws_bytes = lib.ocrfa_beam_candidates_workspace_bytes(B, T, Cn, bw, K)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        labels = torch.full((B, K, T), -1, dtype=torch.int32, device=dev)
        lens = torch.full((B, K), -1, dtype=torch.int32, device=dev)
        count = torch.full((B,), -1, dtype=torch.int32, device=dev)
        _lib.check(lib.ocrfa_ctc_beam_candidates(C.c_void_p(m.data_ptr()), None, None, B, T, Cn, bw, K, None,
                                                 C.c_void_p(labels.data_ptr()), C.c_void_p(lens.data_ptr()),
                                                 C.c_void_p(count.data_ptr()), C.c_void_p(ws.data_ptr()), ws_bytes, stream))
        h_lab, h_len, h_cnt = labels.cpu().numpy(), lens.cpu().numpy(), count.cpu().numpy()
        ws1 = torch.empty(lib.ocrfa_beam_workspace_bytes(B, T, Cn, bw), dtype=torch.uint8, device=dev)
        lab1 = torch.empty((B, T), dtype=torch.int32, device=dev)
        len1 = torch.empty(B, dtype=torch.int32, device=dev)
        _lib.check(lib.ocrfa_ctc_beam(C.c_void_p(m.data_ptr()), B, T, Cn, bw, None, C.c_void_p(lab1.data_ptr()),
                                      C.c_void_p(len1.data_ptr()), C.c_void_p(ws1.data_ptr()), ws1.numel(), stream))
        lab1, len1 = lab1.cpu().numpy(), len1.cpu().numpy()
        for b in range(B):
            want = [R.beam_collapse_ref(lab) for lab in R.beam_candidates_ref(probs[b], bw, K)]
            assert h_cnt[b] == len(want), (bw, K, b)
            got = [h_lab[b, k, :h_len[b, k]].tolist() for k in range(h_cnt[b])]
            assert got == want, (bw, K, b)
            assert lab1[b, :len1[b]].tolist() == got[0]


# ------------------------------------------------------------------------------------------------------------------
# CPU: host span bookkeeping of the drop-in against the oracle's independent restatement, random inputs
# ------------------------------------------------------------------------------------------------------------------
def test_word_spans_match_the_oracle_on_random_rows():
    from ocrfa_b200.converter import word_spans

    rng = np.random.default_rng(5)
    for T in (1, 2, 7, 33):
        am = rng.integers(0, 4, size=(9, T))
        am[0, :] = 2          # a line of spaces only: no span
        am[1, :] = 1          # no space at all: one span
        lines, row0, slen = word_spans(am, 2)
        want = [(i, span[0], len(span)) for i in range(am.shape[0]) for span in R.word_spans_ref(am[i], 2)]
        assert [(b, int(r) - b * T, int(n)) for b, r, n in zip(lines, row0, slen)] == want
        assert 0 not in lines and lines.count(1) == 1


def test_word_segmentation_dropin_equals_oracle_on_random_rows():
    from ocrfa_b200.converter import word_segmentation

    rng = np.random.default_rng(6)
    for _ in range(300):
        T = int(rng.integers(1, 40))
        row = np.repeat(rng.integers(0, 7, size=T), rng.integers(1, 4, size=T))[:T]
        assert word_segmentation(row) == R.word_segmentation_ref(row)
    custom = {"fa": [5, 6]}
    row = np.array([0, 5, 5, 1, 2, 6, 6, 3, 5, 6])
    assert word_segmentation(row, custom, [5, 6]) == R.word_segmentation_ref(row, custom, [5, 6]) == \
        [["", [0, 1]], ["fa", [3, 4]], ["", [6, 7]], ["fa", [9, 8]]]
