"""Generates tests/golden/converter_options.npz: outputs of the reference's own `CTCLabelConverter` built WITH language
separators and dictionary files (utils.py:1044-1081) - the options the default Farsi recognizer leaves empty - on
seeded synthetic inputs: `decode_greedy` and `decode_beamsearch` with the separators' ignore_idx, and
`decode_wordbeamsearch` in its three shapes (separators + per-language dictionaries, :1152-1172; no separators + one
dictionary, :1130-1150 with BeamState.wordsearch, :853-872; separators + empty dictionaries). The dictionaries are
chosen so that the lookup matters: they hold lower-ranked candidates of about half of the words (picked with the
oracle's candidate list) next to words that never occur. TEST INFRASTRUCTURE, dev container only (needs /root/reference):

    python oracle/gen_golden_converter.py
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import recognizer_ref as R  # noqa: E402
from oracle._refimport import import_reference  # noqa: E402

SEPS = "‹›«»"          # th start / end, en start / end: classes 1-4, as word_segmentation assumes
LETTERS = "abcdefghijkl"
CHARSET_SEP = SEPS + " " + LETTERS           # separators first (utils.py:1058-1062 counts them, it does not look them up)
CHARSET_PLAIN = " " + LETTERS
SEPARATORS = {"th": [SEPS[0], SEPS[1]], "en": [SEPS[2], SEPS[3]]}


def synth_probs(rng, B, T, C, word_classes, structure, strength=2.2):
    """Peaked but ambiguous probabilities: a target class per timestep (blank runs, repeated letters, `structure`
    classes such as separators / spaces at word boundaries) under softmax(strength * onehot + noise)."""
    probs = np.zeros((B, T, C), dtype=np.float32)
    for b in range(B):
        target = []
        while len(target) < T:
            kind = rng.random()
            if kind < 0.25:
                target += [0] * int(rng.integers(1, 3))
            elif kind < 0.45 and structure:
                target += [int(rng.choice(structure))] * int(rng.integers(1, 3))
            else:
                target += [int(rng.choice(word_classes))] * int(rng.integers(1, 4))
        z = rng.normal(0.0, 1.0, size=(T, C))
        z[np.arange(T), target[:T]] += strength * rng.uniform(0.6, 1.6, size=T)
        e = np.exp(z - z.max(axis=1, keepdims=True))
        probs[b] = (e / e.sum(axis=1, keepdims=True)).astype(np.float32)
    return probs


def pick_dictionary(rng, spans_probs, character, beam_width, ignore, decoys):
    """Words for a dictionary file: for about half of the spans a candidate of rank 1-6 (so the lookup changes the
    result), plus decoys."""
    words = list(decoys)
    for m in spans_probs:
        if m.shape[0] == 0 or rng.random() < 0.5:
            continue
        cands = R.beam_candidates_ref(m, beam_width, 20)
        texts = ["".join(character[k] for k in R.beam_collapse_ref(lab, ignore)) for lab in cands]
        r = int(rng.integers(1, 7))
        if r < len(texts) and texts[r]:
            words.append(texts[r])
    return sorted(set(words))


def main():
    utils, _ = import_reference()
    rng = np.random.default_rng(20261019)
    out = {}
    bw = 5
    with tempfile.TemporaryDirectory() as td:
        # ---------------- separators, no dictionary: greedy + plain beam search with ignore_idx = [0..4]
        conv = utils.CTCLabelConverter(CHARSET_SEP, SEPARATORS, {})
        C = len(conv.character)
        assert conv.ignore_idx == [0, 1, 2, 3, 4] and conv.dict_list == {}
        lengths = rng.integers(1, 40, size=16).astype(np.int32)
        idx = np.concatenate([np.repeat(rng.integers(0, C, size=n), rng.integers(1, 3, size=n))[:n] for n in lengths])
        out["greedy_idx"] = idx.astype(np.int64)
        out["greedy_len"] = lengths
        import torch

        out["greedy_text"] = np.array(conv.decode_greedy(idx, torch.IntTensor(lengths)) + [""])
        letters = [conv.dict[c] for c in LETTERS]
        probs = synth_probs(rng, 12, 36, C, letters, [1, 2, 3, 4, conv.dict[" "]])
        out["sep_probs"] = probs
        out["sep_beam5"] = np.array(conv.decode_beamsearch(probs, beamWidth=bw) + [""])

        # ---------------- separators + per-language dictionaries
        am = probs.argmax(axis=2)
        per_lang = {"th": [], "en": []}
        for i in range(probs.shape[0]):
            for lang, (t0, t1) in R.word_segmentation_ref(am[i]):
                if lang:
                    per_lang[lang].append(probs[i, t0:t1 + 1, :])
        n_seg = {k: len(v) for k, v in per_lang.items()}
        print("language-tagged segments:", n_seg)
        paths = {}
        for lang, decoys in (("th", ["abc", "kkk"]), ("en", ["lid", "beef"])):
            words = pick_dictionary(rng, per_lang[lang], conv.character, bw, conv.ignore_idx, decoys)
            paths[lang] = os.path.join(td, lang + ".txt")
            with open(paths[lang], "w", encoding="utf-8") as f:
                f.write("\n".join(words) + "\n")
            out["dict_" + lang] = np.array(words + [""])
        conv_d = utils.CTCLabelConverter(CHARSET_SEP, SEPARATORS, paths)
        out["sep_wordbeam5_dict"] = np.array(conv_d.decode_wordbeamsearch(probs, beamWidth=bw) + [""])
        # empty dictionary files: every segment takes the dictionary-free ending
        empty = {}
        for lang in ("th", "en"):
            empty[lang] = os.path.join(td, lang + "_empty.txt")
            open(empty[lang], "w").close()
        conv_e = utils.CTCLabelConverter(CHARSET_SEP, SEPARATORS, empty)
        out["sep_wordbeam5_nodict"] = np.array(conv_e.decode_wordbeamsearch(probs, beamWidth=bw) + [""])
        changed = sum(a != b for a, b in zip(out["sep_wordbeam5_dict"], out["sep_wordbeam5_nodict"]))
        print(f"separators: dictionary changes {changed} of {probs.shape[0]} lines")

        # ---------------- no separators + one dictionary (the list form, utils.py:1064-1073)
        conv_p0 = utils.CTCLabelConverter(CHARSET_PLAIN, {}, {})
        Cp = len(conv_p0.character)
        letters_p = [conv_p0.dict[c] for c in LETTERS]
        probs_p = synth_probs(rng, 12, 40, Cp, letters_p, [conv_p0.dict[" "]])
        out["plain_probs"] = probs_p
        am_p = probs_p.argmax(axis=2)
        spans = [probs_p[i, span, :] for i in range(probs_p.shape[0]) for span in R.word_spans_ref(am_p[i], conv_p0.dict[" "])]
        for width in (5, 25):
            words = pick_dictionary(rng, spans, conv_p0.character, width, [0], ["jail", "fig"])
            path = os.path.join(td, f"plain{width}.txt")
            with open(path, "w", encoding="utf-8") as f:
                f.write("\n".join(words) + "\n")
            conv_p = utils.CTCLabelConverter(CHARSET_PLAIN, {}, {"fa": path, "missing": os.path.join(td, "nope.txt")})
            assert conv_p.dict_list == words
            out[f"dict_plain{width}"] = np.array(words + [""])
            got = conv_p.decode_wordbeamsearch(probs_p, beamWidth=width)
            base = conv_p0.decode_wordbeamsearch(probs_p, beamWidth=width)
            out[f"plain_wordbeam{width}_dict"] = np.array(got + [""])
            out[f"plain_wordbeam{width}_nodict"] = np.array(base + [""])
            print(f"plain, beam {width}: dictionary changes {sum(a != b for a, b in zip(got, base))} of {len(got)} lines")
    # ---------------- word_segmentation itself (utils.py:785-820) on random argmax rows, runs of separators included
    rows = np.stack([np.repeat(rng.integers(0, 9, size=30), rng.integers(1, 4, size=30))[:48] for _ in range(40)])
    rows[0, :] = 7                # no separator at all
    rows[1, :] = 1                # only a start separator
    rows[2, :4] = [1, 2, 3, 4]    # empty pairs
    code = {"": 0, "th": 1, "en": 2}
    seg = [[i, code[lang], t0, t1] for i in range(rows.shape[0]) for lang, (t0, t1) in utils.word_segmentation(rows[i])]
    out["seg_rows"] = rows.astype(np.int64)
    out["seg_result"] = np.asarray(seg, dtype=np.int64)
    print("word_segmentation:", len(seg), "segments on", rows.shape[0], "rows")
    path = os.path.join(ROOT, "tests", "golden", "converter_options.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
