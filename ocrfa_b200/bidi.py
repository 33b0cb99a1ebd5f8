"""Right-to-left display order for the recognizer's output strings: the step `OCR.recognize` applies to every result
when the model language is Arabic-script (reference: `item[1] = get_display(item[1])`, src/ocrfa/ocr.py:474-478, with
`from bidi.algorithm import get_display`, ocr.py:30).

The arithmetic lives in a third-party dependency that is ABSENT from /root/reference and from this image: python-bidi
(the reference's requirements.txt does not even pin it). This module restates the published algorithm that package
implements - the Unicode Bidirectional Algorithm, UAX #9 - for the text this path can produce:

  * P2-P3  paragraph level from the first strong character (L -> 0, R / AL -> 1, none -> 0), or `base_dir`
  * W1-W7  weak types (NSM, EN after AL -> AN, AL -> R, ES / CS between numbers, ET next to EN, EN after L -> L)
  * N0     paired brackets (BD14-BD16: a 63-deep stack over the characters that are still ON; a pair takes the embedding
           direction when it encloses it, the opposite one only when the text before the pair agrees; marks that followed
           a changed bracket follow it) - `paired_brackets=False` skips the rule
  * N1-N2  neutral types (B, S, WS, ON) from the surrounding strong text, numbers counting as R
  * I1-I2  implicit levels
  * L1     trailing whitespace back to the paragraph level
  * L2     reversal of every contiguous run at level >= k, from the highest level down to the lowest odd level
  * L4     mirrored characters at odd levels

Not restated: explicit embeddings / overrides / isolates (X1-X10, no such format character exists in the recognizer's
184-character charset, utils.py:414-415: its classes are AL, L, EN, AN, ES, ET, CS, NSM, ON, WS) and L3 (combining marks
stay where the reversal puts them, as in python-bidi). A string with an explicit formatting character raises.

Which python-bidi: the reference does not pin one. The releases since 0.5 wrap the Rust `unicode-bidi` crate (current
UAX #9, N0 included) - what `pip install python-bidi` resolves to - while the pure-Python 0.4 line predates N0; the
default here is the current algorithm, `paired_brackets=False` gives the older one. The two differ only on text with
brackets (()[]{} in the charset) around mixed-direction content.

PARITY: pinned against an independent implementation of the same published algorithm, not against python-bidi itself
(absent here). tests/golden/bidi_icu.npz holds ICU 71's display strings and levels for 6 000 seeded cases over the charset
plus Hebrew letters, tabs, further bracket pairs and both forced paragraph directions (oracle/gen_golden_bidi.py, ICU
borrowed from the Nsight Compute installation through ctypes); tests/test_bidi.py requires identical results on all of
them, and a differential run of 420 000 random cases during development found no difference either. One family is left
out of the fixture: a combining mark directly after a paired bracket whose type N0 changes - UAX #9 (since Unicode 8.0)
lets the mark follow the bracket, ICU keeps it neutral; this module follows the published rule (hand-derived tests).
When python-bidi is importable `ocrfa_b200.reader` calls it instead (the reference's own dependency wins);
`OCR.bidi_backend` says which one ran.
"""
from __future__ import annotations

import unicodedata
from typing import List, Optional

_EXPLICIT = {"LRE", "RLE", "LRO", "RLO", "PDF", "LRI", "RLI", "FSI", "PDI"}
_NEUTRAL = ("B", "S", "WS", "ON")

# BidiMirroring.txt pairs for the characters unicodedata.mirrored() flags in Basic Latin, Latin-1, General Punctuation,
# Super/Subscripts and the common mathematical relations (the recognizer's charset only has the first five pairs)
_MIRROR_PAIRS = "()[]{}<>«»‹›⁅⁆⁽⁾₍₎≤≥≦≧≨≩≪≫≮≯≰≱≲≳≴≵≶≷≸≹≺≻≼≽≾≿⊀⊁⊂⊃⊄⊅⊆⊇⊈⊉⊊⊋⊏⊐⊑⊒⟦⟧⟨⟩⟪⟫⟬⟭⟮⟯⦃⦄⦅⦆「」『』【】〔〕〖〗〘〙〚〛"
_MIRROR = {}
for _i in range(0, len(_MIRROR_PAIRS), 2):
    _MIRROR[_MIRROR_PAIRS[_i]] = _MIRROR_PAIRS[_i + 1]
    _MIRROR[_MIRROR_PAIRS[_i + 1]] = _MIRROR_PAIRS[_i]


def _direction(level: int) -> str:
    return "R" if level & 1 else "L"


def base_level(text: str) -> int:
    """P2-P3: level of the first strong character, 0 when there is none."""
    for ch in text:
        t = unicodedata.bidirectional(ch)
        if t in ("AL", "R"):
            return 1
        if t == "L":
            return 0
    return 0


# BidiBrackets.txt: the mirrored pairs above whose members are opening / closing punctuation (gc = Ps / Pe); quotation
# marks such as the guillemets (Pi / Pf) and the relations (Sm) mirror but do not pair
_BRACKET_CLOSE = {a: b for a, b in zip(_MIRROR_PAIRS[0::2], _MIRROR_PAIRS[1::2])
                  if unicodedata.category(a) == "Ps" and unicodedata.category(b) == "Pe"}
_BRACKET_OPEN = {b: a for a, b in _BRACKET_CLOSE.items()}


def _bracket_pairs(text: str, types: List[str]) -> List[tuple]:
    """BD16: (opening, closing) positions of the paired brackets, ordered by opening position. Only characters whose
    type is still ON take part; a stack deeper than 63 abandons the search."""
    stack: List[tuple] = []
    pairs: List[tuple] = []
    for i, ch in enumerate(text):
        if types[i] != "ON":
            continue
        if ch in _BRACKET_CLOSE:
            if len(stack) == 63:
                return []
            stack.append((_BRACKET_CLOSE[ch], i))
        elif ch in _BRACKET_OPEN:
            for k in range(len(stack) - 1, -1, -1):
                if stack[k][0] == ch:
                    pairs.append((stack[k][1], i))
                    del stack[k:]
                    break
    pairs.sort()
    return pairs


def _resolve_paired_brackets(text: str, types: List[str], orig: List[str], e: str, sos: str) -> None:
    """N0: a bracket pair takes the embedding direction `e` when it encloses a strong type of that direction (numbers
    count as R); when it encloses only the opposite direction it takes the opposite direction if the text before the
    opening bracket establishes it, else `e`; a pair around no strong type is left to N1-N2. Marks that followed a
    changed bracket (NSM before W1) follow it."""
    o = "L" if e == "R" else "R"

    def strong(t: str) -> Optional[str]:
        return "R" if t in ("R", "AN", "EN") else ("L" if t == "L" else None)

    for a, b in _bracket_pairs(text, types):
        inside = {strong(types[k]) for k in range(a + 1, b)}
        if e in inside:
            new = e
        elif o in inside:
            before = sos
            for k in range(a - 1, -1, -1):
                if strong(types[k]):
                    before = strong(types[k])
                    break
            new = o if before == o else e
        else:
            continue
        for pos in (a, b):
            types[pos] = new
            k = pos + 1
            while k < len(types) and orig[k] == "NSM":
                types[k] = new
                k += 1


def resolve_levels(text: str, base_dir: Optional[str] = None, paired_brackets: bool = True) -> List[int]:
    """Embedding level of every character after W1-W7, N0-N2, I1-I2 and L1 (one paragraph, no explicit formatting).
    paired_brackets=False skips N0 (UAX #9 before Unicode 6.3; python-bidi 0.4.x)."""
    types = [unicodedata.bidirectional(ch) for ch in text]
    for t in types:
        if t in _EXPLICIT:
            raise ValueError("explicit bidi formatting characters are not supported by this restatement")
    if base_dir is None:
        base = base_level(text)
    elif base_dir in ("L", "R"):
        base = 1 if base_dir == "R" else 0
    else:
        raise ValueError("base_dir must be 'L', 'R' or None")
    n = len(text)
    orig = list(types)
    sor = eor = _direction(base)  # the whole paragraph is one level run

    # W1: NSM takes the type of the previous character (sor at the start); W2: EN after AL becomes AN
    prev_strong = prev_type = sor
    for i in range(n):
        t = types[i]
        if t == "NSM":
            types[i] = t = prev_type
        if t == "EN" and prev_strong == "AL":
            types[i] = "AN"
        if t in ("R", "L", "AL"):
            prev_strong = t
        prev_type = types[i]
    # W3: AL -> R
    for i in range(n):
        if types[i] == "AL":
            types[i] = "R"
    # W4: a single ES between two EN -> EN; a single CS between two numbers of the same type -> that type
    for i in range(1, n - 1):
        t, p, q = types[i], types[i - 1], types[i + 1]
        if t == "ES" and p == q == "EN":
            types[i] = "EN"
        if t == "CS" and p == q and p in ("AN", "EN"):
            types[i] = p
    # W5: a sequence of ET adjacent to EN -> EN
    for i in range(n):
        if types[i] == "EN":
            j = i - 1
            while j >= 0 and types[j] == "ET":
                types[j] = "EN"
                j -= 1
            j = i + 1
            while j < n and types[j] == "ET":
                types[j] = "EN"
                j += 1
    # W6: remaining separators and terminators -> ON
    for i in range(n):
        if types[i] in ("ET", "ES", "CS"):
            types[i] = "ON"
    # W7: EN -> L when the previous strong type is L
    prev_strong = sor
    for i in range(n):
        if types[i] == "EN" and prev_strong == "L":
            types[i] = "L"
        if types[i] in ("L", "R"):
            prev_strong = types[i]

    # N0: paired brackets
    if paired_brackets:
        _resolve_paired_brackets(text, types, orig, _direction(base), sor)

    # N1-N2: runs of neutrals take the direction both neighbours agree on (numbers count as R), else the embedding's
    ext = [sor] + types + [eor]
    start = None
    prev_t = sor
    for i in range(len(ext)):
        if ext[i] in _NEUTRAL:
            if start is None:
                start = i
                prev_t = ext[i - 1]
        elif start is not None:
            a = "R" if prev_t in ("AN", "EN") else prev_t
            b = "R" if ext[i] in ("AN", "EN") else ext[i]
            fill = a if a == b else _direction(base)
            for j in range(start, i):
                ext[j] = fill
            start = None
    types = ext[1:-1]

    # I1-I2
    levels = [base] * n
    for i in range(n):
        t = types[i]
        if base & 1 == 0:
            if t == "R":
                levels[i] += 1
            elif t != "L":
                levels[i] += 2
        elif t != "R":
            levels[i] += 1
    # L1: segment / paragraph separators and the whitespace before them or at the end of the line -> paragraph level
    reset = True
    for i in range(n - 1, -1, -1):
        if orig[i] in ("B", "S"):
            levels[i] = base
            reset = True
        elif reset and orig[i] in ("BN", "WS"):
            levels[i] = base
        else:
            reset = False
    return levels


def get_display(text: str, base_dir: Optional[str] = None, paired_brackets: bool = True) -> str:
    """Logical -> visual order of one paragraph (the `bidi.algorithm.get_display(text)` call of ocr.py:477)."""
    if not text:
        return text
    levels = resolve_levels(text, base_dir, paired_brackets)
    chars = list(text)
    # L4 first (it depends on the resolved level, not on the position)
    for i, lv in enumerate(levels):
        if lv & 1 and chars[i] in _MIRROR:
            chars[i] = _MIRROR[chars[i]]
    items = list(zip(levels, chars))
    highest = max(levels)
    odd = [lv for lv in levels if lv & 1]
    if not odd:
        return "".join(chars)
    # L2: from the highest level down to the lowest odd level, reverse every contiguous run at that level or higher
    for level in range(highest, min(odd) - 1, -1):
        i = 0
        while i < len(items):
            if items[i][0] >= level:
                j = i
                while j + 1 < len(items) and items[j + 1][0] >= level:
                    j += 1
                items[i:j + 1] = items[i:j + 1][::-1]
                i = j + 1
            else:
                i += 1
    return "".join(ch for _, ch in items)
