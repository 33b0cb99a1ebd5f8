"""Character tables of the reference's default Farsi recognizer (drop-in data, not code).

Values mirror /root/reference/src/ocrfa/utils.py:408-416 (recognition_models["gen1"]["arabic_g1"]) and the
hard-coded Farsi letter list of OCR.setLanguageList (/root/reference/src/ocrfa/ocr.py:273-297). They are kept as
\\u escapes so the file is encoding-proof; tests/test_charset.py checks them against the golden fixture
written by oracle/gen_golden.py from the reference itself.
"""

ARABIC_G1_CHARACTERS = (
    "0123456789!\u0022#$%&'()*+,-./:;<=>?@[\u005c]^_`{|}~ abcdefghijklmnopqrstuvwxyzABCDEFGHIJK"
    "LMNOPQRSTUVWXYZ\u0660\u0661\u0662\u0663\u0664\u0665\u0666\u0667\u0668\u0669\u00ab\u00bb"
    "\u061f\u060c\u061b\u0621\u0622\u0623\u0624\u0625\u0626\u0627\u0627\u064b\u0628\u0629\u062a"
    "\u062b\u062c\u062d\u062e\u062f\u0630\u0631\u0632\u0633\u0634\u0635\u0636\u0637\u0638\u0639"
    "\u063a\u0641\u0642\u0643\u0644\u0645\u0646\u0647\u0648\u0649\u064a\u064b\u064c\u064d\u064e"
    "\u064f\u0650\u0651\u0652\u0653\u0654\u0670\u0671\u0679\u067e\u0686\u0688\u0691\u0698\u06a9"
    "\u06ad\u06af\u06ba\u06be\u06c0\u06c1\u06c2\u06c3\u06c6\u06c7\u06c8\u06cb\u06cc\u06d0\u06d2"
    "\u06d3\u06d5"
)

ARABIC_G1_SYMBOLS = (
    "\u00ab\u00bb\u061f\u060c\u061b\u0660\u0661\u0662\u0663\u0664\u0665\u0666\u0667\u0668\u0669"
    "0123456789!\u0022#$%&'()*+,-./:;<=>?@[\u005c]^_`{|}~ "
)

FA_LANG_LETTERS = (
    "\u0622\u0627\u0628\u067e\u062a\u062b\u062c\u0686\u062d\u062e\u062f\u0630\u0631\u0632\u0698"
    "\u0633\u0634\u0635\u0636\u0637\u0638\u0639\u063a\u0641\u0642\u06a9\u06af\u0644\u0645\u0646"
    "\u0648\u0647\u06cc\u0621\u0626\u0650\u0643\u064a\u064b\u0651\u064e\u0629\u0623\u0654\u0624"
    "\u064f\u0649\u0625\u06c0\u064d\u0652\u064c\u0670\u0653\u0671"
)


def default_character() -> str:
    """`OCR.character` for the default model (ocr.py:180-185 -> model["characters"])."""
    return ARABIC_G1_CHARACTERS


def default_lang_char() -> str:
    """`OCR.lang_char` after setLanguageList(["fa"], arabic_g1): Farsi letters union the model symbols."""
    return "".join(sorted(set(FA_LANG_LETTERS).union(set(ARABIC_G1_SYMBOLS))))


def default_ignore_char() -> str:
    """`ignore_char` of OCR.recognize with no allowlist/blocklist (ocr.py:384-389)."""
    return "".join(sorted(set(ARABIC_G1_CHARACTERS) - set(default_lang_char())))
