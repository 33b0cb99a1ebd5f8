"""Drop-in for `CTCLabelConverter` (/root/reference/src/ocrfa/utils.py:1041-1173).

Same constructor, attributes (`character`, `dict`, `ignore_idx`, `separator_list`, `dict_list`) and method
signatures. The collapse / beam searches run on the GPU through libocrfa_b200.so; only the index -> character
lookup, the word segmentation bookkeeping (integer work on B*T argmax flags) and the dictionary membership test
(strings) stay on the host, as in the reference.
Language separators and dictionary files are not used by the default Farsi recognizer (`separator_list = {}`,
no dict/ directory in the repo: ocr.py:102, 172-174); they are supported with the reference's semantics all the same
(utils.py:1058-1081, :785-820, :853-872, :1152-1172): the separators' indices are dropped by every collapse, word beam
search segments by separator pairs and looks the 20 best candidates of a word up in its language's dictionary
(`ocrfa_ctc_beam_candidates`).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib


def _dev() -> torch.device:
    return _lib.resolve_device("cuda")


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class CTCLabelConverter:
    """Convert between text-label and text-index (index 0 = "[blank]")."""

    def __init__(self, character, separator_list=None, dict_pathlist=None):
        separator_list = separator_list if separator_list is not None else {}
        dict_pathlist = dict_pathlist if dict_pathlist is not None else {}
        chars = list(character)
        self.dict = {ch: i + 1 for i, ch in enumerate(chars)}  # later duplicates win, like the reference loop
        self.character = ["[blank]"] + chars
        self.separator_list = separator_list
        # utils.py:1058-1062: one ignored index per separator character, COUNTED from 1 (the separators are taken to be
        # the first characters of the charset, as word_segmentation's fixed class numbers 1-4 also assume)
        n_sep = sum(len(sep) for sep in separator_list.values())
        self.ignore_idx = [0] + [i + 1 for i in range(n_sep)]
        if self.ignore_idx[-1] >= len(self.character):
            raise ValueError("more separator characters than characters")
        if len(separator_list) == 0:
            # utils.py:1065-1073: one flat word list; unreadable dictionary files are skipped silently
            words: List[str] = []
            for _lang, path in dict_pathlist.items():
                try:
                    with open(path, "r", encoding="utf-8-sig") as f:
                        words += f.read().splitlines()
                except Exception:
                    pass
            self.dict_list = words
        else:
            # utils.py:1074-1079: language -> word list; a missing file is an error here, as in the reference
            by_lang = {}
            for lang, path in dict_pathlist.items():
                with open(path, "r", encoding="utf-8-sig") as f:
                    by_lang[lang] = f.read().splitlines()
            self.dict_list = by_lang
        self._collapse_mask_cache = {}

    # ------------------------------------------------------------------------------------------------------
    def encode(self, text, batch_max_length=25):
        """utils.py:1083-1097."""
        length = [len(s) for s in text]
        flat = [self.dict[ch] for ch in "".join(text)]
        return torch.IntTensor(flat), torch.IntTensor(length)

    def labels_to_text(self, labels: np.ndarray, lengths: np.ndarray) -> List[str]:
        """Rows of collapsed label indices (first lengths[b] valid) -> strings."""
        chars = self.character
        return ["".join(chars[k] for k in row[:n]) for row, n in zip(labels.tolist(), lengths.tolist())]

    def _collapse_mask(self, device):
        """uint8 [C] device mask of converter.ignore_idx, or None when it is [0] (the kernels always drop blank)."""
        if len(self.ignore_idx) <= 1:
            return None
        key = (str(device), tuple(self.ignore_idx))
        m = self._collapse_mask_cache.get(key)
        if m is None:
            host = np.zeros(len(self.character), dtype=np.uint8)
            host[np.asarray(self.ignore_idx, dtype=np.int64)] = 1
            m = torch.from_numpy(host).to(device)
            self._collapse_mask_cache = {key: m}
        return m

    # ------------------------------------------------------------------------------------------------------
    def decode_greedy(self, text_index, length) -> List[str]:
        """utils.py:1099-1115: flat argmax indices + per-line lengths -> collapsed strings."""
        dev = _dev()
        lib = _lib.load()
        idx = torch.as_tensor(np.asarray(text_index)).to(dev, torch.int32).contiguous().view(-1)
        lens = [int(v) for v in (length.tolist() if hasattr(length, "tolist") else length)]
        if not lens:
            return []
        offs = np.zeros(len(lens) + 1, dtype=np.int32)
        offs[1:] = np.cumsum(lens)
        if offs[-1] > idx.numel():
            raise ValueError("lengths exceed the number of indices")
        d_off = torch.from_numpy(offs).to(dev)
        labels = torch.empty(max(1, int(offs[-1])), dtype=torch.int32, device=dev)
        out_len = torch.empty(len(lens), dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.ocrfa_ctc_collapse(_ptr(idx), _ptr(d_off), len(lens), len(self.character),
                                          _ptr(self._collapse_mask(dev)), _ptr(labels), _ptr(out_len), C.c_void_p(stream)))
        lab = labels.cpu().numpy()
        n = out_len.cpu().numpy()
        chars = self.character
        return ["".join(chars[k] for k in lab[o:o + m].tolist()) for o, m in zip(offs[:-1].tolist(), n.tolist())]

    # ------------------------------------------------------------------------------------------------------
    def beam_labels(self, mat: torch.Tensor, beamWidth: int):
        """mat: CUDA float32 [B, T, C] probabilities -> (labels int32 [B, T], lengths int32 [B]) on the device."""
        lib = _lib.load()
        B, T, Cn = mat.shape
        ws_bytes = lib.ocrfa_beam_workspace_bytes(B, T, Cn, int(beamWidth))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=mat.device)
        labels = torch.empty((B, T), dtype=torch.int32, device=mat.device)
        lens = torch.empty(B, dtype=torch.int32, device=mat.device)
        stream = torch.cuda.current_stream(mat.device).cuda_stream
        _lib.check(lib.ocrfa_ctc_beam(_ptr(mat), B, T, Cn, int(beamWidth), _ptr(self._collapse_mask(mat.device)),
                                      _ptr(labels), _ptr(lens), _ptr(ws), ws_bytes, C.c_void_p(stream)))
        return labels, lens

    def _as_device_mat(self, mat) -> torch.Tensor:
        dev = _dev()
        if isinstance(mat, torch.Tensor):
            m = mat.to(dev, torch.float32)
        else:
            m = torch.from_numpy(np.ascontiguousarray(mat, dtype=np.float32)).to(dev)
        if m.dim() != 3 or m.shape[2] != len(self.character):
            raise ValueError(f"expected [B, T, {len(self.character)}] probabilities, got {tuple(m.shape)}")
        return m.contiguous()

    def decode_beamsearch(self, mat, beamWidth=5) -> List[str]:
        """utils.py:1117-1122: CTC prefix beam search per line (lm=None, no dictionary)."""
        m = self._as_device_mat(mat)
        if m.shape[0] == 0:
            return []
        labels, lens = self.beam_labels(m, beamWidth)
        return self.labels_to_text(labels.cpu().numpy(), lens.cpu().numpy())

    # ------------------------------------------------------------------------------------------------------
    MAX_CANDIDATE = 20  # BeamState.wordsearch's maxCandidate as ctcBeamSearch passes it (utils.py:1037)

    def _search_spans(self, m: torch.Tensor, row0: np.ndarray, slen: np.ndarray, beamWidth: int,
                      dicts: Sequence[Optional[Sequence[str]]]) -> List[str]:
        """ctcBeamSearch (utils.py:949-1038) over row spans of `m` seen as [B*T, C]: span i covers rows row0[i] ..
        + slen[i]. dicts[i] empty / None -> the best labeling's text; else the first of the 20 best candidates that is
        a word of dicts[i], the best one when none is (BeamState.wordsearch, utils.py:853-872). A span without rows
        decodes to "" either way."""
        n = len(row0)
        out = [""] * n
        live = [i for i in range(n) if slen[i] > 0]
        if not live:
            return out
        lib = _lib.load()
        dev = m.device
        Cn = m.shape[-1]
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        mask = self._collapse_mask(dev)
        plain = [i for i in live if not dicts[i]]
        looked = [i for i in live if dicts[i]]
        if plain:
            t_max = int(max(slen[i] for i in plain))
            d_row0 = torch.from_numpy(np.asarray([row0[i] for i in plain], dtype=np.int32)).to(dev)
            d_len = torch.from_numpy(np.asarray([slen[i] for i in plain], dtype=np.int32)).to(dev)
            ws_bytes = lib.ocrfa_beam_workspace_bytes(len(plain), t_max, Cn, int(beamWidth))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            labels = torch.empty((len(plain), t_max), dtype=torch.int32, device=dev)
            lens = torch.empty(len(plain), dtype=torch.int32, device=dev)
            _lib.check(lib.ocrfa_ctc_beam_spans(_ptr(m), _ptr(d_row0), _ptr(d_len), len(plain), t_max, Cn, int(beamWidth),
                                                _ptr(mask), _ptr(labels), _ptr(lens), _ptr(ws), ws_bytes, stream))
            for i, w in zip(plain, self.labels_to_text(labels.cpu().numpy(), lens.cpu().numpy())):
                out[i] = w
        if looked:
            K = self.MAX_CANDIDATE
            t_max = int(max(slen[i] for i in looked))
            d_row0 = torch.from_numpy(np.asarray([row0[i] for i in looked], dtype=np.int32)).to(dev)
            d_len = torch.from_numpy(np.asarray([slen[i] for i in looked], dtype=np.int32)).to(dev)
            ws_bytes = lib.ocrfa_beam_candidates_workspace_bytes(len(looked), t_max, Cn, int(beamWidth), K)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            labels = torch.empty((len(looked), K, t_max), dtype=torch.int32, device=dev)
            lens = torch.empty((len(looked), K), dtype=torch.int32, device=dev)
            count = torch.empty(len(looked), dtype=torch.int32, device=dev)
            _lib.check(lib.ocrfa_ctc_beam_candidates(_ptr(m), _ptr(d_row0), _ptr(d_len), len(looked), t_max, Cn,
                                                     int(beamWidth), K, _ptr(mask), _ptr(labels), _ptr(lens), _ptr(count),
                                                     _ptr(ws), ws_bytes, stream))
            h_lab, h_len, h_cnt = labels.cpu().numpy(), lens.cpu().numpy(), count.cpu().numpy()
            sets = {}
            for j, i in enumerate(looked):
                words = sets.get(id(dicts[i]))
                if words is None:
                    words = sets[id(dicts[i])] = set(dicts[i])
                texts = self.labels_to_text(h_lab[j, :h_cnt[j]], h_len[j, :h_cnt[j]])
                out[i] = next((t for t in texts if t in words), texts[0])
        return out

    def decode_wordbeamsearch(self, mat, beamWidth=5) -> List[str]:
        """utils.py:1124-1173. Without separators: split each line into maximal runs of timesteps whose argmax is not
        the space class, search every run (against the flat dictionary when there is one), join with single spaces.
        With separators: segments from `word_segmentation` (its default class numbers 1-4, as the reference calls it),
        each searched against its language's dictionary (none for text outside separator pairs), concatenated."""
        m = self._as_device_mat(mat)
        B, T, Cn = m.shape
        if B == 0:
            return []
        lib = _lib.load()
        dev = m.device
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        am = torch.empty((B, T), dtype=torch.int32, device=dev)
        _lib.check(lib.ocrfa_argmax_rows(_ptr(m), B * T, Cn, _ptr(am), stream))
        am_h = am.cpu().numpy()
        per_line: List[List[str]] = [[] for _ in range(B)]
        if len(self.separator_list) == 0:
            lines, row0, slen = word_spans(am_h, self.dict[" "])
            dicts = [self.dict_list] * len(lines)
            joiner = " "
        else:
            lines, r0, sl, dicts = [], [], [], []
            for b in range(B):
                for lang, (t0, t1) in word_segmentation(am_h[b]):
                    lines.append(b)
                    r0.append(b * T + t0)
                    sl.append(max(0, t1 - t0 + 1))
                    dicts.append(None if lang == "" else self.dict_list[lang])  # KeyError without that dictionary, as in the reference
            row0 = np.asarray(r0, dtype=np.int64)
            slen = np.asarray(sl, dtype=np.int64)
            joiner = ""
        for line, w in zip(lines, self._search_spans(m, row0, slen, beamWidth, dicts)):
            per_line[line].append(w)
        return [joiner.join(ws_) for ws_ in per_line]


def word_spans(argmax: np.ndarray, space_idx: int):
    """utils.py:1132-1136 for a whole batch: the maximal runs of timesteps whose argmax is not the space class. `argmax`
    is [B, T]; returns (line of every span, its first row in the [B*T, C] view of the probabilities, its length), spans
    ordered by line then time (host integer work on B*T flags, as in the reference)."""
    B, T = argmax.shape
    pad = np.zeros((B, T + 2), dtype=np.int8)
    pad[:, 1:-1] = argmax != space_idx
    d = np.diff(pad, axis=1)
    starts = np.argwhere(d == 1)   # (line, t0) sorted by line then t0
    ends = np.argwhere(d == -1)    # (line, t1 + 1)
    row0 = (starts[:, 0] * T + starts[:, 1]).astype(np.int64)
    slen = (ends[:, 1] - starts[:, 1]).astype(np.int64)
    return starts[:, 0].tolist(), row0, slen


def _run_edges(pos: np.ndarray, last: bool) -> List[int]:
    """utils.py:773-782 (`consecutive`): the first or last element of every run of consecutive integers."""
    if pos.size == 0:
        return []
    brk = np.flatnonzero(np.diff(pos) != 1)
    return pos[np.concatenate([brk, [pos.size - 1]])].tolist() if last else pos[np.concatenate([[0], brk + 1])].tolist()


def word_segmentation(mat, separator_idx=None, separator_idx_list=None):
    """utils.py:785-820: `mat` is one line's per-timestep argmax. A language's start separator (odd class number) counts
    at the LAST timestep of each of its runs, its end separator (even) at the FIRST; a start followed by the matching
    end gives [lang, [t0, t1]] for the timesteps strictly between them, the text outside such pairs ['', [t0, t1]]."""
    if separator_idx_list is None:
        separator_idx_list = [1, 2, 3, 4]
    if separator_idx is None:
        separator_idx = {"th": [1, 2], "en": [3, 4]}
    mat = np.asarray(mat)
    marks = []
    for sep in separator_idx_list:
        marks += [(t, sep) for t in _run_edges(np.flatnonzero(mat == sep), last=sep % 2 != 0)]
    marks.sort(key=lambda mk: mk[0])
    result = []
    start_idx = 0
    open_lang, open_at = "", 0
    for t, sep in marks:
        for lang, pair in separator_idx.items():
            if sep == pair[0]:
                open_lang, open_at = lang, t
            elif sep == pair[1]:
                if open_lang == lang:
                    if open_at > start_idx:
                        result.append(["", [start_idx, open_at - 1]])
                    start_idx = t + 1
                    result.append([lang, [open_at + 1, t - 1]])
                open_lang = ""
    if start_idx <= len(mat) - 1:
        result.append(["", [start_idx, len(mat) - 1]])
    return result
