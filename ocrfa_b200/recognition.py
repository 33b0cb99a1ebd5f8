"""Host-side mirror of the reference recognition pipeline (/root/reference/src/ocrfa/utils.py:533-770):
`get_recognizer`, `get_text`, `recognizer_predict`, `AlignCollate`, `ListDataset`, `custom_mean` keep their names,
argument meaning and error behaviour; the arithmetic behind them runs in libocrfa_b200.so.

What stays on the host, exactly where the reference has it: list plumbing, packing of the ragged uint8 lines and
the size-only coefficient tables of the resize (ocrfa_b200/resize.py), and the index -> character lookup. The
Pillow-exact BICUBIC resize (utils.py:598) and the percentile contrast stretch of the second pass
(utils.py:450-463) run on the GPU (ocrfa_align_collate_u8).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .converter import CTCLabelConverter
from .model import Model
from .resize import align_collate_device, resized_width


def custom_mean(x):
    """utils.py:446-447."""
    return x.prod() ** (2.0 / np.sqrt(len(x)))


def contrast_grey(img):
    """utils.py:450-453."""
    high = np.percentile(img, 90)
    low = np.percentile(img, 10)
    return (high - low) / np.maximum(10, high + low), high, low


def adjust_contrast_grey(img, target=0.4):
    """utils.py:456-463."""
    contrast, high, low = contrast_grey(img)
    if contrast < target:
        img = img.astype(int)
        ratio = 200.0 / np.maximum(10, high - low)
        img = (img - low + 25) * ratio
        img = np.clip(img, 0, 255).astype(np.uint8)
    return img


class LineBatch:
    """One collated batch before normalisation: uint8 [B, imgH, imgW] (pinned when possible) + valid widths."""

    def __init__(self, images_u8: torch.Tensor, widths: torch.Tensor):
        self.images_u8 = images_u8
        self.widths = widths

    def size(self, dim=None):
        s = (self.images_u8.shape[0], 1, self.images_u8.shape[1], self.images_u8.shape[2])
        return s if dim is None else s[dim]


class ListDataset:
    """utils.py:553-564 (indexable list of uint8 grayscale lines)."""

    def __init__(self, image_list):
        self.image_list = image_list
        self.nSamples = len(image_list)

    def __len__(self):
        return self.nSamples

    def __getitem__(self, index):
        return self.image_list[index]


class AlignCollate:
    """utils.py:567-602. Produces a LineBatch (uint8 at model height + widths); the /255, (x-0.5)/0.5 and the
    replicate-last-column padding happen on the GPU (ocrfa_preprocess_u8 / the fused stem)."""

    def __init__(self, imgH=32, imgW=100, keep_ratio_with_pad=False, adjust_contrast=0.0):
        self.imgH = imgH
        self.imgW = imgW
        self.keep_ratio_with_pad = keep_ratio_with_pad
        self.adjust_contrast = adjust_contrast

    def __call__(self, batch) -> LineBatch:
        dev_lines = len(batch) > 0 and all(hasattr(b, "buffer") and hasattr(b, "off") for b in batch)  # crops.DeviceLine
        images = list(batch) if dev_lines else [np.asarray(b) for b in batch if b is not None]
        for img in images:
            if img.ndim != 2 or img.dtype != np.uint8:
                raise ValueError("lines must be uint8 grayscale arrays [h, w]")
        rws = [resized_width(img.shape[1], img.shape[0], self.imgH, self.imgW) for img in images]
        identity = (not dev_lines and self.adjust_contrast <= 0 and
                    all(img.shape == (self.imgH, rw) for img, rw in zip(images, rws)))
        if not identity:
            # contrast stretch and/or Pillow-exact BICUBIC resize: GPU kernels (ocrfa_align_collate_u8)
            if not torch.cuda.is_available():
                raise _lib.OcrfaError("AlignCollate needs a CUDA device for lines that are not at model height "
                                      "(no CPU fallback)")
            dev = images[0].buffer.device if dev_lines else torch.device("cuda", torch.cuda.current_device())
            out, widths = align_collate_device(images, self.imgH, self.imgW, self.adjust_contrast, dev)
            return LineBatch(out, widths)
        out = torch.empty((len(images), self.imgH, self.imgW), dtype=torch.uint8)
        if torch.cuda.is_available():
            out = out.pin_memory()
        out_np = out.numpy()
        for i, (img, rw) in enumerate(zip(images, rws)):
            out_np[i, :, :rw] = img
            out_np[i, :, rw:] = 0  # overwritten on the device by the replicate pad
        return LineBatch(out, torch.tensor(rws, dtype=torch.int32))


def _batches(dataset: ListDataset, batch_size: int, collate) -> Iterable[LineBatch]:
    for i in range(0, len(dataset), batch_size):
        yield collate([dataset[j] for j in range(i, min(i + batch_size, len(dataset)))])


def _ignore_mask(num_class: int, ignore_idx: Sequence[int]) -> np.ndarray:
    m = np.zeros(num_class, dtype=np.uint8)
    for k in ignore_idx:
        if 0 < int(k) < num_class:
            m[int(k)] = 1
    return m


def _decode_device(logits: torch.Tensor, converter: CTCLabelConverter, mask_np: np.ndarray, decoder: str,
                   beamWidth: int):
    """logits CUDA float32 [B,T,C] -> (strings, confidences) via the decode kernels."""
    lib = _lib.load()
    B, T, Cn = logits.shape
    dev = logits.device
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    mask = torch.from_numpy(mask_np).to(dev)
    labels = torch.empty((B, T), dtype=torch.int32, device=dev)
    lens = torch.empty(B, dtype=torch.int32, device=dev)
    conf = torch.empty(B, dtype=torch.float64, device=dev)
    need_probs = decoder != "greedy"
    probs = torch.empty((B, T, Cn), dtype=torch.float32, device=dev) if need_probs else None
    _lib.check(lib.ocrfa_ctc_greedy(C.c_void_p(logits.data_ptr()), B, T, Cn, C.c_void_p(mask.data_ptr()), None,
                                    C.c_void_p(probs.data_ptr()) if need_probs else None, None, None,
                                    C.c_void_p(labels.data_ptr()), C.c_void_p(lens.data_ptr()),
                                    C.c_void_p(conf.data_ptr()), stream))
    if decoder == "greedy":
        strs = converter.labels_to_text(labels.cpu().numpy(), lens.cpu().numpy())
    elif decoder == "beamsearch":
        strs = converter.decode_beamsearch(probs, beamWidth=beamWidth)
    elif decoder == "wordbeamsearch":
        strs = converter.decode_wordbeamsearch(probs, beamWidth=beamWidth)
    else:
        raise ValueError(f"unknown decoder {decoder!r}")
    return strs, conf.cpu().numpy()


def recognizer_predict(model: Model, converter: CTCLabelConverter, test_loader, batch_max_length, ignore_idx,
                       char_group_idx, decoder="greedy", beamWidth=5, device="cuda"):
    """utils.py:605-662. `test_loader` yields LineBatch objects (this package's AlignCollate) or, like the
    reference, float tensors [B,1,H,W] already normalised. Returns [[text, confidence], ...]."""
    if str(device) == "cpu":
        raise _lib.OcrfaError("ocrfa_b200 has no CPU path; use device='cuda'")
    dev = _lib.resolve_device(device)
    mask_np = _ignore_mask(len(converter.character), ignore_idx)
    result = []
    lib = _lib.load()
    for batch in test_loader:
        if isinstance(batch, LineBatch):
            if decoder in ("greedy", "beamsearch") and not batch.images_u8.is_cuda:
                strs, conf = recognize_host(model, converter, batch.images_u8, batch.widths, mask_np, decoder, beamWidth, dev)
            else:
                logits = model.forward_u8(batch.images_u8.to(dev, non_blocking=True), batch.widths.to(dev))
                strs, conf = _decode_device(logits, converter, mask_np, decoder, beamWidth)
        else:
            image = batch.to(dev)
            logits = model(image, None)
            strs, conf = _decode_device(logits, converter, mask_np, decoder, beamWidth)
        for s, c in zip(strs, conf.tolist()):
            result.append([s, c])
    del lib
    return result


# A greedy string is CERTIFIED when the line's decision margin (ocrfa_ctc_margin: smallest top-2 logit gap over the
# timesteps, relative to the rms of the row's logits) exceeds this bound: ~9 standard deviations of the f16x3 mode's
# measured logit-difference error (2.4e-6 of the rms per logit, tools/error_budget.py), so every evaluation of the network
# within fp32-grade error - the reference's own - takes the same argmax at every timestep of such a line. Lines below it
# hold a near-tie that the reference itself resolves differently when only its batch size changes
# (profiles/r02_ref_self_consistency.json).
CERTIFY_MARGIN = 3e-5


def recognize_host(model: Model, converter: CTCLabelConverter, images_u8: torch.Tensor, widths: Optional[torch.Tensor],
                   mask_np: Optional[np.ndarray], decoder: str, beamWidth: int, dev: torch.device,
                   return_margin: bool = False):
    """One batch through ocrfa_recognize_host_ex: host uint8 lines in, strings + confidences out (the end-to-end call:
    H2D, preprocess, forward, decode and D2H all inside). return_margin=True also returns the per-line decision margin
    of the greedy decode (float32 [B]); the last call's margins stay on `model.last_margins` either way."""
    lib = _lib.load()
    if images_u8.is_cuda or images_u8.dtype != torch.uint8 or images_u8.dim() != 3:
        raise ValueError("images_u8 must be a host uint8 tensor [B, H, W]")
    img = images_u8.contiguous()
    B, H, W = img.shape
    T = lib.ocrfa_out_timesteps(H, W)
    if T == 0:
        raise ValueError(f"unsupported line geometry H={H} W={W}")
    with torch.cuda.device(dev):
        handle = model._ensure_handle(dev)
        bw = int(beamWidth) if decoder == "beamsearch" else 1
        ws_bytes = lib.ocrfa_recognize_workspace_bytes(handle, B, H, W, bw)
        ws = model.workspace(ws_bytes, dev)
        labels = torch.empty((B, T), dtype=torch.int32).pin_memory()
        lens = torch.empty(B, dtype=torch.int32).pin_memory()
        conf = torch.empty(B, dtype=torch.float64).pin_memory()
        margin = torch.empty(B, dtype=torch.float32).pin_memory()
        w32 = widths.to(torch.int32).contiguous() if widths is not None else None
        mask = np.ascontiguousarray(mask_np, dtype=np.uint8) if mask_np is not None else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.ocrfa_recognize_host_ex(
            handle, C.c_void_p(img.data_ptr()), C.c_void_p(w32.data_ptr()) if w32 is not None else None, B, H, W,
            mask.ctypes.data_as(C.c_void_p) if mask is not None else None, 1 if decoder == "beamsearch" else 0, bw,
            C.c_void_p(labels.data_ptr()), C.c_void_p(lens.data_ptr()), C.c_void_p(conf.data_ptr()),
            C.c_void_p(margin.data_ptr()), 0, C.c_void_p(ws.data_ptr()), ws_bytes, C.c_void_p(stream)))
    model.last_margins = margin.numpy().copy()
    strs = converter.labels_to_text(labels.numpy(), lens.numpy())
    if return_margin:
        return strs, conf.numpy().copy(), model.last_margins
    return strs, conf.numpy().copy()


def get_recognizer(recog_network, network_params, character, separator_list, dict_list, model_path, device="cuda",
                   quantize=True, precision="f16x3"):
    """utils.py:665-695. `quantize` (dynamic int8 of the reference's CPU path) has no meaning here and is ignored;
    `precision`: "f16x3" (default: fp32-grade tensor-core arithmetic, the mode whose strings are checked against the
    reference), "bf16" / "f16" (single-pass tensor-core modes, 3.5x faster, 16-bit rounding noise in the logits) or
    "fp32" (CUDA-core arithmetic with fp64 accumulation, 25x slower; the reference point of the parity runs)."""
    if str(device) == "cpu":
        raise _lib.OcrfaError("ocrfa_b200 has no CPU path; use device='cuda'")
    converter = CTCLabelConverter(character, separator_list, dict_list)
    num_class = len(converter.character)
    if recog_network != "generation1":
        raise ValueError("only the reference's reachable recognizer ('generation1': ResNet + BiLSTM + CTC) exists")
    model = Model(num_class=num_class, precision=precision, **network_params)
    state = torch.load(model_path, map_location="cpu")
    model.load_state_dict(state)
    return model, converter


def get_text(character, imgH, imgW, recognizer, converter, image_list, ignore_char="", decoder="greedy", beamWidth=5,
             batch_size=1, contrast_ths=0.1, adjust_contrast=0.5, filter_ths=0.003, workers=1, device="cuda"):
    """utils.py:698-770: first pass, contrast-adjusted second pass over low-confidence lines, keep the better."""
    batch_max_length = int(imgW / 10)
    ignore_idx = []
    for ch in ignore_char:
        k = character.find(ch) if isinstance(character, str) else (list(character).index(ch) if ch in character else -1)
        if k >= 0:
            ignore_idx.append(k + 1)
    coord = [item[0] for item in image_list]
    img_list = [item[1] for item in image_list]
    loader = _batches(ListDataset(img_list), batch_size, AlignCollate(imgH=imgH, imgW=imgW, keep_ratio_with_pad=True))
    result1 = recognizer_predict(recognizer, converter, loader, batch_max_length, ignore_idx, {}, decoder, beamWidth,
                                 device=device)
    low = [i for i, item in enumerate(result1) if item[1] < contrast_ths]
    result2 = []
    if low:
        loader2 = _batches(ListDataset([img_list[i] for i in low]), batch_size,
                           AlignCollate(imgH=imgH, imgW=imgW, keep_ratio_with_pad=True, adjust_contrast=adjust_contrast))
        result2 = recognizer_predict(recognizer, converter, loader2, batch_max_length, ignore_idx, {}, decoder,
                                     beamWidth, device=device)
    pos = {i: n for n, i in enumerate(low)}
    result = []
    for i, (box, pred1) in enumerate(zip(coord, result1)):
        if i in pos:
            pred2 = result2[pos[i]]
            best = pred1 if pred1[1] > pred2[1] else pred2
            result.append((box, best[0], best[1]))
        else:
            result.append((box, pred1[0], pred1[1]))
    return result
