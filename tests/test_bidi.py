"""CPU tests of ocrfa_b200.bidi.get_display (the RTL display-order step of OCR.recognize, ocr.py:474-478).

python-bidi is absent from this image and from /root/reference, so there is no reference output to pin against
(PARITY UNPINNED, see the module docstring): the expected strings below are derived by hand from UAX #9."""
import unicodedata

import pytest

from ocrfa_b200.bidi import base_level, get_display, resolve_levels
from ocrfa_b200.charset import default_character


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
