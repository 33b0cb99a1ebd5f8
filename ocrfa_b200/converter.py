"""Drop-in for `CTCLabelConverter` (/root/reference/src/ocrfa/utils.py:1041-1173).

Same constructor, attributes (`character`, `dict`, `ignore_idx`, `separator_list`, `dict_list`) and method
signatures. The collapse / beam searches run on the GPU through libocrfa_b200.so; only the index -> character
lookup and the (language-separator-free) word segmentation bookkeeping stay on the host, as in the reference.
Language separators and dictionary files are not used by the default Farsi recognizer (`separator_list = {}`,
no dict/ directory in the repo: ocr.py:102, 172-174) and are rejected here instead of being half-supported.
"""
from __future__ import annotations

import ctypes as C
from typing import List

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
        separator_list = separator_list or {}
        dict_pathlist = dict_pathlist or {}
        chars = list(character)
        self.dict = {ch: i + 1 for i, ch in enumerate(chars)}  # later duplicates win, like the reference loop
        self.character = ["[blank]"] + chars
        self.separator_list = separator_list
        if len(separator_list) != 0:
            raise NotImplementedError("language separators are not part of the Farsi recognition path")
        self.ignore_idx = [0]
        # utils.py:1065-1073: unreadable dictionary files are skipped silently; none ship with the reference
        words: List[str] = []
        for _lang, path in dict_pathlist.items():
            try:
                with open(path, "r", encoding="utf-8-sig") as f:
                    words += f.read().splitlines()
            except OSError:
                pass
        if words:
            raise NotImplementedError("dictionary-constrained word beam search is outside the hot path")
        self.dict_list = words
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
        # converter.ignore_idx is [0] without separators: blank is always dropped by the kernels -> no mask needed
        return None

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
        _lib.check(lib.ocrfa_ctc_collapse(_ptr(idx), _ptr(d_off), len(lens), len(self.character), None, _ptr(labels),
                                          _ptr(out_len), C.c_void_p(stream)))
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
        _lib.check(lib.ocrfa_ctc_beam(_ptr(mat), B, T, Cn, int(beamWidth), None, _ptr(labels), _ptr(lens), _ptr(ws),
                                      ws_bytes, C.c_void_p(stream)))
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

    def decode_wordbeamsearch(self, mat, beamWidth=5) -> List[str]:
        """utils.py:1124-1151 with no separators and an empty dictionary: split each line into maximal runs of
        timesteps whose argmax is not the space class, beam-search every run, join with single spaces."""
        m = self._as_device_mat(mat)
        B, T, Cn = m.shape
        if B == 0:
            return []
        lib = _lib.load()
        dev = m.device
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        am = torch.empty((B, T), dtype=torch.int32, device=dev)
        _lib.check(lib.ocrfa_argmax_rows(_ptr(m), B * T, Cn, _ptr(am), stream))
        space = self.dict[" "]
        not_space = (am.cpu().numpy() != space)
        # span bookkeeping (host integer work on B*T flags, as in the reference)
        pad = np.zeros((B, T + 2), dtype=np.int8)
        pad[:, 1:-1] = not_space
        d = np.diff(pad, axis=1)
        starts = np.argwhere(d == 1)   # (line, t0) sorted by line then t0
        ends = np.argwhere(d == -1)    # (line, t1 + 1)
        n_spans = starts.shape[0]
        per_line: List[List[str]] = [[] for _ in range(B)]
        if n_spans:
            row0 = (starts[:, 0] * T + starts[:, 1]).astype(np.int32)
            slen = (ends[:, 1] - starts[:, 1]).astype(np.int32)
            t_max = int(slen.max())
            d_row0 = torch.from_numpy(row0).to(dev)
            d_len = torch.from_numpy(slen).to(dev)
            ws_bytes = lib.ocrfa_beam_workspace_bytes(n_spans, t_max, Cn, int(beamWidth))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            labels = torch.empty((n_spans, t_max), dtype=torch.int32, device=dev)
            lens = torch.empty(n_spans, dtype=torch.int32, device=dev)
            _lib.check(lib.ocrfa_ctc_beam_spans(_ptr(m), _ptr(d_row0), _ptr(d_len), n_spans, t_max, Cn, int(beamWidth),
                                                None, _ptr(labels), _ptr(lens), _ptr(ws), ws_bytes, stream))
            words = self.labels_to_text(labels.cpu().numpy(), lens.cpu().numpy())
            for line, w in zip(starts[:, 0].tolist(), words):
                per_line[line].append(w)
        return [" ".join(ws_) for ws_ in per_line]
