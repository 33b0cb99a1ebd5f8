"""CPU tests of ocrfa_b200.bidi.get_display (the RTL display-order step of OCR.recognize, ocr.py:474-478).

python-bidi is absent from this image and from /root/reference, so its own output cannot be recorded. The restatement is
pinned against an INDEPENDENT implementation of the same published algorithm instead: tests/golden/bidi_icu.npz holds
ICU's display order and levels for 6 000 seeded cases (oracle/gen_golden_bidi.py); the hand-derived UAX #9 cases below
cover what that fixture leaves out."""
import unicodedata

import pytest

from conftest import load_golden
from ocrfa_b200.bidi import base_level, get_display, resolve_levels
from ocrfa_b200.charset import default_character


def test_display_order_and_levels_match_icu():
    """6 000 cases (2 000 random strings over the charset + Hebrew, tab, more brackets; paragraph level from the first
    strong character, forced LTR, forced RTL): identical display strings; identical levels wherever ICU reports
    them (it flattens a paragraph that needs no reordering to its paragraph level)."""
    g = load_golden("bidi_icu")
    logical, visual, levels = g["logical"][:-1].tolist(), g["visual"][:-1].tolist(), g["levels"][:-1].tolist()
    base = g["base"].tolist()
    assert len(logical) == 6000 and len(set(logical)) >= 1900
    n_levels = 0
    for s, b, v, l in zip(logical, base, visual, levels):
        bd = (None, "L", "R")[b]
        assert get_display(s, bd) == v, (s, bd)
        want = [ord(c) - 48 for c in l]
        if len(set(want)) > 1:
            assert resolve_levels(s, bd) == want, (s, bd)
            n_levels += 1
    assert n_levels > 3000


@pytest.mark.parametrize("logical, base_dir, levels", [
    # N0: a pair takes the embedding direction when it encloses it, the opposite one only if the text before agrees
    ("a(ب)b", "L", [0, 0, 1, 0, 0]),            # encloses R only, L before -> embedding direction L
    ("ب(ب)a", "L", [1, 1, 1, 1, 0]),            # R before the pair -> R
    ("ب(a)ب", "R", [1, 1, 2, 1, 1]),            # encloses L only, R before -> embedding direction R
    ("ب [a ب] a", "L", [1, 0, 0, 0, 0, 1, 0, 0, 0]),  # encloses L = embedding direction
    ("ب (!) a", "L", [1, 0, 0, 0, 0, 0, 0]),    # no strong type inside: left to N1 / N2
    ("ب(a]ب)", "R", [1, 1, 2, 1, 1, 1]),        # the unmatched ']' is skipped, '(' pairs with ')'
    # a combining mark after a bracket that N0 changed follows the bracket (UAX #9 since Unicode 8.0; ICU keeps it neutral)
    ("ب(ب)ًa", "L", [1, 1, 1, 1, 1, 0]),
    ("٤(ب)ً", "L", [2, 1, 1, 1, 1]),
])
def test_paired_brackets_hand_derived(logical, base_dir, levels):
    assert resolve_levels(logical, base_dir) == levels


def test_without_the_paired_bracket_rule():
    """paired_brackets=False: UAX #9 before Unicode 6.3 (what the pure-Python python-bidi 0.4 line implements)."""
    s = "a(ب)b"
    assert resolve_levels(s, "L", paired_brackets=False) == [0, 0, 1, 0, 0]
    s = "ب(ب)a"
    assert resolve_levels(s, "L") == [1, 1, 1, 1, 0] and resolve_levels(s, "L", paired_brackets=False) == [1, 1, 1, 0, 0]
    assert get_display(s, "L") == "(ب)بa" and get_display(s, "L", paired_brackets=False) == "ب)ب)a"


@pytest.mark.parametrize("logical, visual", [
    ("", ""),
    ("abc", "abc"),                                   # all L: untouched
    ("سلام", "مالس"),                                  # all AL: reversed
    ("سال 1402 است", "تسا 1402 لاس"),                  # EN after AL -> AN (W2): digits keep their order (level 2)
    ("٣٤ کتاب", "باتک ٣٤"),                            # AN digits first: paragraph level from the first STRONG char
    ("1 س 23", "23 س 1"),
    ("abc سلام", "abc مالس"),                          # L paragraph with an embedded RTL word
    ("hello سلام world", "hello مالس world"),
    ("سلام abc دنیا", "ایند abc مالس"),                # RTL paragraph with an embedded LTR word
    ("(سلام)", "(مالس)"),                              # neutrals between R and eor -> R; brackets mirrored (L4)
    ("سلام!", "!مالس"),                                # trailing ON takes the paragraph direction (N2 / eor)
    ("قیمت: 12.50$", "$12.50 :تمیق"),                  # CS between AN digits -> AN (W4); ET next to AN -> ON (W6)
    ("a (b) c", "a (b) c"),
    ("12-34", "12-34"),                               # no strong char: level 0, ES between EN -> EN, numbers level 2
])
def test_get_display_hand_derived(logical, visual):
    assert get_display(logical) == visual


def test_levels_and_base_level():
    assert base_level("123 سلام") == 1 and base_level("x سلام") == 0 and base_level("12 .") == 0
    assert resolve_levels("سال 1402") == [1, 1, 1, 1, 2, 2, 2, 2]
    assert resolve_levels("ab 12") == [0, 0, 0, 0, 0]          # W7: EN after L -> L
    assert resolve_levels("سلام  ") == [1, 1, 1, 1, 1, 1]      # L1: trailing whitespace at the paragraph level
    assert resolve_levels("ab  ", base_dir="R") == [2, 2, 1, 1]


def test_combining_marks_travel_with_the_reversal():
    s = "مَن"  # meem + fatha (NSM) + noon: W1 gives the mark its base's type; no L3 reordering (as python-bidi)
    assert get_display(s) == s[::-1]
    assert unicodedata.bidirectional(s[1]) == "NSM"


def test_involution_on_single_direction_text_and_charset_coverage():
    chars = default_character()
    # every class of the recognizer's charset is handled (no explicit formatting characters in it)
    assert {unicodedata.bidirectional(c) for c in chars} <= {"AL", "L", "EN", "AN", "ES", "ET", "CS", "NSM", "ON", "WS"}
    rtl = "".join(c for c in chars if unicodedata.bidirectional(c) == "AL")
    assert get_display(rtl) == rtl[::-1] and get_display(get_display(rtl)) == rtl
    # the display string is a permutation of the logical one up to mirroring
    unmirror = str.maketrans("()[]{}<>«»", "(([[{{<<««")
    assert sorted(get_display(chars).translate(unmirror)) == sorted(chars.translate(unmirror))


def test_explicit_formatting_is_rejected():
    with pytest.raises(ValueError):
        get_display("a‮b")
    with pytest.raises(ValueError):
        get_display("ab", base_dir="x")
