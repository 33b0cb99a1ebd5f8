"""Generates tests/golden/bidi_icu.npz: display order and embedding levels of seeded random strings computed by ICU's
`ubidi` (a conformant implementation of the Unicode Bidirectional Algorithm, UAX #9 with the paired-bracket rule N0),
to pin `ocrfa_b200.bidi.get_display` - the restatement of the algorithm behind `bidi.algorithm.get_display`, the
last step of `OCR.recognize` for Arabic-script models (src/ocrfa/ocr.py:474-478). python-bidi itself is absent from
this image and from /root/reference (it is not even listed in requirements.txt), so the pin is an independent
implementation of the same published algorithm rather than that package's output.

ICU is not a dependency of anything here either: the script borrows the copy that ships inside Nsight Compute
(libicuuc.so.71 under /opt/nvidia/nsight-compute/*/host/*/) through ctypes. TEST INFRASTRUCTURE, dev container only:

    python oracle/gen_golden_bidi.py

Strings: the recognizer's 184-character charset (classes AL, L, EN, AN, ES, ET, CS, NSM, ON, WS) plus Hebrew letters (R),
tab (S), no-break space, Arabic decimal / thousands separators, more bracket pairs and guillemets; each string is resolved
with the paragraph level taken from its first strong character and with both forced levels (get_display's base_dir).
One family of strings is left out and counted: a combining mark (NSM) directly after a paired bracket. UAX #9 (since
Unicode 8.0) lets such a mark follow the bracket when N0 changes the bracket's type; ICU 71 keeps it neutral (its
bracket code turns a mark after a neutral into a neutral before the pair is resolved), so the two differ there by
design, and the restatement follows the published rule (hand-derived cases in tests/test_bidi.py).
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import random
import re
import sys
import unicodedata

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ocrfa_b200.charset import default_character  # noqa: E402

EXTRA = "אבגד\t ٫٬⟨⟩「」⁅⁆⟦⟧<>«»±°%‰"
BRACKETS = "()[]{}⟨⟩「」⁅⁆⟦⟧"
UBIDI_DEFAULT_LTR = 0xFE
UBIDI_DO_MIRRORING = 2


def load_icu():
    hits = sorted(glob.glob("/opt/nvidia/nsight-compute/*/host/*/libicuuc.so.*"))
    if not hits:
        raise SystemExit("no ICU library found (expected the copy inside Nsight Compute)")
    path = hits[-1]
    ver = path.rsplit(".", 1)[-1]
    C.CDLL(os.path.join(os.path.dirname(path), "libicudata.so." + ver), mode=C.RTLD_GLOBAL)
    uc = C.CDLL(path, mode=C.RTLD_GLOBAL)
    f = lambda name: getattr(uc, f"{name}_{ver}")  # noqa: E731  (ICU exports versioned names)
    f("ubidi_open").restype = C.c_void_p
    f("ubidi_close").argtypes = [C.c_void_p]
    f("ubidi_setPara").argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_uint8, C.c_void_p, C.POINTER(C.c_int)]
    f("ubidi_getLevels").argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    f("ubidi_getLevels").restype = C.POINTER(C.c_uint8)
    f("ubidi_writeReordered").argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_uint16, C.POINTER(C.c_int)]
    f("ubidi_writeReordered").restype = C.c_int32
    return f, ver


def icu_display(f, text: str, para_level: int):
    """(visual string with mirroring, embedding levels) of one paragraph of BMP text."""
    n = len(text)
    src = C.create_string_buffer(text.encode("utf-16-le"), 2 * n + 2)
    err = C.c_int(0)
    bidi = f("ubidi_open")()
    f("ubidi_setPara")(bidi, src, n, para_level, None, C.byref(err))
    assert err.value <= 0, err.value
    lv = f("ubidi_getLevels")(bidi, C.byref(err))
    levels = [lv[i] for i in range(n)]
    dst = C.create_string_buffer(2 * n + 2)
    m = f("ubidi_writeReordered")(bidi, dst, n, UBIDI_DO_MIRRORING, C.byref(err))
    assert err.value <= 0 and m == n, (err.value, m)
    f("ubidi_close")(bidi)
    return dst.raw[:2 * m].decode("utf-16-le"), levels


def main():
    f, ver = load_icu()
    chars = default_character()
    nsm = "".join(sorted({c for c in chars if unicodedata.bidirectional(c) == "NSM"}))
    mark_after_bracket = re.compile("[" + re.escape(BRACKETS) + "][" + nsm + "]")
    rnd = random.Random(20261019)
    pool = chars + EXTRA * 3 + "()[]{}" * 4 + " " * 10
    logical, base, visual, levels = [], [], [], []
    skipped = 0
    while len(logical) < 6000:
        s = "".join(rnd.choice(pool) for _ in range(rnd.randint(1, 30)))
        if mark_after_bracket.search(s):
            skipped += 1
            continue
        for code, lvl in ((0, UBIDI_DEFAULT_LTR), (1, 0), (2, 1)):   # base_dir None / 'L' / 'R'
            v, l = icu_display(f, s, lvl)
            logical.append(s)
            base.append(code)
            visual.append(v)
            levels.append("".join(chr(48 + x) for x in l))
    out = {"icu_version": np.array(ver), "logical": np.array(logical + [""]), "base": np.asarray(base, dtype=np.int8),
           "visual": np.array(visual + [""]), "levels": np.array(levels + [""]),
           "skipped_mark_after_bracket": np.int64(skipped)}
    path = os.path.join(ROOT, "tests", "golden", "bidi_icu.npz")
    np.savez_compressed(path, **out)
    print(f"ICU {ver}: {len(logical)} cases, {skipped} strings with a mark after a bracket left out; wrote {path} "
          f"({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    main()
