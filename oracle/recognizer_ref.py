"""CPU restatement of the OCR-FA recognition hot path (TEST INFRASTRUCTURE - the ORACLE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this module,
and only as the checker or the timed CPU baseline. The product path (ocrfa_b200/) never does.

Parity pin: the reference has no tests or golden vectors, so this restatement is pinned
against outputs of the reference ITSELF, produced in the dev container by oracle/gen_golden.py and committed
under tests/golden/ (tests/test_oracle_golden.py). Floating-point stages use torch fp32 CPU ops exactly as the
reference does (the arithmetic lives in torch/numpy/Pillow, not in the reference tree); byte/integer stages
(CTC collapse, beam bookkeeping) are plain numpy/Python.

Each function cites the reference lines it follows (paths relative to /root/reference/).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm2d default, src/ocrfa/modules/blocks.py:201


# --------------------------------------------------------------------------------------------------------------
# a1: pre-processing  (src/ocrfa/utils.py:533-602)
# --------------------------------------------------------------------------------------------------------------
def adjust_contrast_grey_ref(img: np.ndarray, target: float = 0.4) -> np.ndarray:
    """utils.py:450-463: percentile 10/90 contrast stretch when contrast < target."""
    high = np.percentile(img, 90)
    low = np.percentile(img, 10)
    contrast = (high - low) / np.maximum(10, high + low)
    if contrast < target:
        img = img.astype(int)
        ratio = 200.0 / np.maximum(10, high - low)
        img = (img - low + 25) * ratio
        img = np.maximum(np.full(img.shape, 0), np.minimum(np.full(img.shape, 255), img)).astype(np.uint8)
    return img


def resized_width_ref(w: int, h: int, img_h: int, img_w: int) -> int:
    """utils.py:592-596."""
    ratio = w / float(h)
    return img_w if math.ceil(img_h * ratio) > img_w else math.ceil(img_h * ratio)


# Pillow's Image.resize(..., BICUBIC) for 8-bit images, restated (third-party arithmetic: Pillow 10.1 pinned by the
# reference's requirements.txt:4, 12.2 in this container; src/libImaging/Resample.c: precompute_coeffs,
# normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc). Two passes, horizontal first; every pass
# rounds to uint8: out = clip8((2^21 + sum_k pixel_k * round(coef_k * 2^22)) >> 22).
PRECISION_BITS = 32 - 8 - 2


def _bicubic_filter(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pillow_coeffs_ref(in_size: int, out_size: int):
    """Per output position: first input index, number of taps, fixed-point coefficients (precompute_coeffs with
    box (0, in_size), filter support 2.0). Returns (bounds int32 [out,2], kk int32 [out,ksize], ksize)."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ww = 0.0
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [0.0] * ksize
        for x in range(xmax):
            w = _bicubic_filter((x + xmin - center + 0.5) * ss)
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        for x in range(ksize):  # normalize_coeffs_8bpc
            v = k[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if k[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _resample_axis_ref(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    in_size = img.shape[axis]
    bounds, kk, _ = pillow_coeffs_ref(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        xmin, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(n):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pillow_bicubic_resize_ref(img_u8: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """uint8 [h, w] -> uint8 [out_h, out_w], bit-exact with PIL.Image.resize((out_w, out_h), Image.BICUBIC)."""
    h, w = img_u8.shape
    out = img_u8
    if out_w != w:
        out = _resample_axis_ref(out, out_w, axis=1)
    if out_h != h:
        out = _resample_axis_ref(out, out_h, axis=0)
    return np.ascontiguousarray(out)


def normalize_pad_ref(img_u8: np.ndarray, img_h: int, img_w: int, adjust_contrast: float = 0.0) -> np.ndarray:
    """One line: uint8 [h, w] -> float32 [1, img_h, img_w] in [-1, 1], right-padded with the last column.

    utils.py:562-564 (PIL 'L' image), :587-590 (optional contrast), :592-598 (BICUBIC resize, identity when the
    size already matches), :542-543 (ToTensor /255 then (x-0.5)/0.5), :545-548 (replicate-last-column pad).
    """
    from PIL import Image

    if adjust_contrast > 0:
        img_u8 = adjust_contrast_grey_ref(img_u8, target=adjust_contrast)
    h, w = img_u8.shape
    rw = resized_width_ref(w, h, img_h, img_w)
    if (rw, img_h) != (w, h):
        img_u8 = np.asarray(Image.fromarray(img_u8, "L").resize((rw, img_h), Image.BICUBIC))
    x = img_u8.astype(np.float32) / np.float32(255.0)
    x = (x - np.float32(0.5)) / np.float32(0.5)
    out = np.empty((1, img_h, img_w), dtype=np.float32)
    out[0, :, :rw] = x
    if rw != img_w:
        out[0, :, rw:] = x[:, rw - 1:rw]
    return out


def align_collate_ref(imgs: Sequence[np.ndarray], img_h: int, img_w: int, adjust_contrast: float = 0.0) -> np.ndarray:
    """utils.py:575-602: list of uint8 lines -> float32 [B, 1, img_h, img_w]."""
    return np.stack([normalize_pad_ref(im, img_h, img_w, adjust_contrast) for im in imgs], 0)


# --------------------------------------------------------------------------------------------------------------
# a2-a5: model forward  (src/ocrfa/modules/blocks.py:186-391, 85-102, 422-447)
# --------------------------------------------------------------------------------------------------------------
def _t(sd, key):
    v = sd[key]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))


def _conv_bn(sd, x, conv, bn, stride=1, padding=0, relu=True):
    y = F.conv2d(x, _t(sd, conv + ".weight"), None, stride=stride, padding=padding)
    y = F.batch_norm(y, _t(sd, bn + ".running_mean"), _t(sd, bn + ".running_var"), _t(sd, bn + ".weight"),
                     _t(sd, bn + ".bias"), False, 0.0, BN_EPS)
    return F.relu(y) if relu else y


def _block_b(sd, x, pre):
    """BlockB.forward, blocks.py:219-223 (+ downsample from _make_layer, blocks.py:324-331)."""
    residual = x
    if pre + "downsample.0.weight" in sd:
        residual = _conv_bn(sd, x, pre + "downsample.0", pre + "downsample.1", relu=False)
    y = _conv_bn(sd, x, pre + "conv1", pre + "bn1", padding=1)
    y = _conv_bn(sd, y, pre + "conv2", pre + "bn2", padding=1, relu=False)
    return F.relu(y + residual)


def resnet_ref(sd: Dict[str, np.ndarray], x: torch.Tensor, prefix: str = "FeatureExtraction.ConvNet.") -> torch.Tensor:
    """ResNet.forward, blocks.py:340-391. x: [B,1,H,W] -> [B,512,H',T]."""
    p = prefix
    x = _conv_bn(sd, x, p + "conv0_1", p + "bn0_1", padding=1)
    x = _conv_bn(sd, x, p + "conv0_2", p + "bn0_2", padding=1)
    x = F.max_pool2d(x, 2, 2, 0)                                       # blocks.py:259
    for b in range(1):
        x = _block_b(sd, x, f"{p}layer1.{b}.")
    x = _conv_bn(sd, x, p + "conv1", p + "bn1", padding=1)
    x = F.max_pool2d(x, 2, 2, 0)                                       # blocks.py:271-275
    for b in range(2):
        x = _block_b(sd, x, f"{p}layer2.{b}.")
    x = _conv_bn(sd, x, p + "conv2", p + "bn2", padding=1)
    x = F.max_pool2d(x, 2, (2, 1), (0, 1))                             # blocks.py:292
    for b in range(5):
        x = _block_b(sd, x, f"{p}layer3.{b}.")
    x = _conv_bn(sd, x, p + "conv3", p + "bn3", padding=1)
    for b in range(3):
        x = _block_b(sd, x, f"{p}layer4.{b}.")
    x = _conv_bn(sd, x, p + "conv4_1", p + "bn4_1", stride=(2, 1), padding=(0, 1))   # blocks.py:305-313
    x = _conv_bn(sd, x, p + "conv4_2", p + "bn4_2", stride=1, padding=0)             # blocks.py:314-322
    return x


def lstm_dir_ref(x: torch.Tensor, w_ih, w_hh, b_ih, b_hh, reverse: bool) -> torch.Tensor:
    """One direction of nn.LSTM (blocks.py:89), PyTorch gate order i,f,g,o, h0=c0=0. x: [B,T,I] -> [B,T,H]."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    xp = x @ w_ih.t() + b_ih + b_hh  # hoisted input projection (same math, different association than cuDNN)
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    out = x.new_empty(B, T, H)
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        g = xp[:, t] + h @ w_hh.t()
        i, f, gg, o = g.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[:, t] = h
    return out


def bilstm_ref(sd, x: torch.Tensor, pre: str) -> torch.Tensor:
    """BidirectionalLSTM.forward, blocks.py:92-102: LSTM(bidirectional) -> Linear(1024 -> 512)."""
    fwd = lstm_dir_ref(x, _t(sd, pre + "rnn.weight_ih_l0"), _t(sd, pre + "rnn.weight_hh_l0"),
                       _t(sd, pre + "rnn.bias_ih_l0"), _t(sd, pre + "rnn.bias_hh_l0"), False)
    bwd = lstm_dir_ref(x, _t(sd, pre + "rnn.weight_ih_l0_reverse"), _t(sd, pre + "rnn.weight_hh_l0_reverse"),
                       _t(sd, pre + "rnn.bias_ih_l0_reverse"), _t(sd, pre + "rnn.bias_hh_l0_reverse"), True)
    return F.linear(torch.cat([fwd, bwd], 2), _t(sd, pre + "linear.weight"), _t(sd, pre + "linear.bias"))


def model_forward_ref(sd: Dict[str, np.ndarray], x, return_features: bool = False):
    """Model.forward, blocks.py:440-447. x: float32 [B,1,H,W] -> logits float32 [B,T,C]."""
    with torch.no_grad():
        x = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
        f = resnet_ref(sd, x.float())                       # [B,512,H',T]
        v = f.permute(0, 3, 1, 2).mean(dim=3)               # AdaptiveAvgPool2d((None,1)) + squeeze(3): [B,T,512]
        s = bilstm_ref(sd, v, "SequenceModeling.0.")
        s = bilstm_ref(sd, s, "SequenceModeling.1.")
        logits = F.linear(s, _t(sd, "Prediction.weight"), _t(sd, "Prediction.bias"))
    if return_features:
        return logits, v
    return logits


class _QuantizedHead(torch.nn.Module):
    """The recognizer's nn.LSTM / nn.Linear modules (blocks.py:85-102, 438), the ones get_recognizer's
    torch.quantization.quantize_dynamic(model, dtype=torch.qint8) rewrites when quantize=True (utils.py:686-690)."""

    def __init__(self, sd):
        super().__init__()
        num_class = _t(sd, "Prediction.weight").shape[0]
        self.rnn0 = torch.nn.LSTM(512, 512, bidirectional=True, batch_first=True)
        self.lin0 = torch.nn.Linear(1024, 512)
        self.rnn1 = torch.nn.LSTM(512, 512, bidirectional=True, batch_first=True)
        self.lin1 = torch.nn.Linear(1024, 512)
        self.pred = torch.nn.Linear(512, num_class)
        with torch.no_grad():
            for li, (rnn, lin) in enumerate(((self.rnn0, self.lin0), (self.rnn1, self.lin1))):
                pre = f"SequenceModeling.{li}."
                for name, param in rnn.named_parameters():
                    param.copy_(_t(sd, pre + "rnn." + name))
                lin.weight.copy_(_t(sd, pre + "linear.weight"))
                lin.bias.copy_(_t(sd, pre + "linear.bias"))
            self.pred.weight.copy_(_t(sd, "Prediction.weight"))
            self.pred.bias.copy_(_t(sd, "Prediction.bias"))

    def forward(self, v):
        s = self.lin0(self.rnn0(v)[0])
        s = self.lin1(self.rnn1(s)[0])
        return self.pred(s)


_QUANT_HEADS: Dict[int, torch.nn.Module] = {}


def model_forward_int8_ref(sd: Dict[str, np.ndarray], x):
    """Model.forward as the reference's CPU default runs it (quantize=True): fp32 ResNet, dynamically quantized int8
    LSTM / Linear layers. A recorded baseline column, not a parity target."""
    head = _QUANT_HEADS.get(id(sd))
    if head is None:
        head = _QuantizedHead(sd).eval()
        torch.quantization.quantize_dynamic(head, dtype=torch.qint8, inplace=True)
        _QUANT_HEADS.clear()
        _QUANT_HEADS[id(sd)] = head
    with torch.no_grad():
        x = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
        f = resnet_ref(sd, x.float())
        v = f.permute(0, 3, 1, 2).mean(dim=3)
        return head(v)


# --------------------------------------------------------------------------------------------------------------
# bf16 emulation of the same forward: identical math, with the tensor-core path's rounding points
# (weights with BN folded -> bf16; every stored activation -> bf16; fp32 accumulate, bias, gate math, cell state).
# Used to check the bf16 kernels tightly: against the fp32 reference a random-weight network amplifies ANY bf16
# evaluation's rounding noise beyond 2e-2 (tools/bf16_sensitivity.py), which says nothing about the kernels.
# --------------------------------------------------------------------------------------------------------------
def _bf(t: torch.Tensor) -> torch.Tensor:
    return t.bfloat16().float()


def _fold(sd, conv, bn):
    w = _t(sd, conv + ".weight").double()
    s = _t(sd, bn + ".weight").double() / torch.sqrt(_t(sd, bn + ".running_var").double() + BN_EPS)
    b = _t(sd, bn + ".bias").double() - _t(sd, bn + ".running_mean").double() * s
    return (w * s[:, None, None, None]).float(), b.float()


def _conv_q(sd, x, conv, bn, stride=1, padding=0, extra=None):
    w, b = _fold(sd, conv, bn)
    y = F.conv2d(x, _bf(w), None, stride=stride, padding=padding)
    if extra is not None:
        y = y + extra
    return _bf(F.relu(y + b[None, :, None, None]))


def _block_q(sd, x, pre):
    y = _conv_q(sd, x, pre + "conv1", pre + "bn1", padding=1)
    if pre + "downsample.0.weight" in sd:  # fused into conv2's accumulator: never rounded on its own
        wd, bd = _fold(sd, pre + "downsample.0", pre + "downsample.1")
        extra = F.conv2d(x, _bf(wd), None) + bd[None, :, None, None]
    else:
        extra = x
    return _conv_q(sd, y, pre + "conv2", pre + "bn2", padding=1, extra=extra)


def _lstm_dir_q(x, w_ih, w_hh, b_ih, b_hh, reverse):
    B, T, _ = x.shape
    H = w_hh.shape[1]
    xp = x @ _bf(w_ih).t() + (b_ih + b_hh)
    whh = _bf(w_hh).t()
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    out = x.new_empty(B, T, H)
    for t in (range(T - 1, -1, -1) if reverse else range(T)):
        i, f, gg, o = (xp[:, t] + h @ whh).chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = _bf(torch.sigmoid(o) * torch.tanh(c))
        out[:, t] = h
    return out


def features_bf16emu_ref(sd, x) -> torch.Tensor:
    """float32 [B,1,H,W] -> visual features [B,T,512] (bf16-valued float32) with the kernel's rounding points."""
    p = "FeatureExtraction.ConvNet."
    with torch.no_grad():
        x = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
        w0, b0 = _fold(sd, p + "conv0_1", p + "bn0_1")          # stem runs in fp32, output stored as bf16
        a = _bf(F.relu(F.conv2d(x.float(), w0, b0, padding=1)))
        a = _conv_q(sd, a, p + "conv0_2", p + "bn0_2", padding=1)
        a = F.max_pool2d(a, 2, 2, 0)
        a = _block_q(sd, a, p + "layer1.0.")
        a = _conv_q(sd, a, p + "conv1", p + "bn1", padding=1)
        a = F.max_pool2d(a, 2, 2, 0)
        for b in range(2):
            a = _block_q(sd, a, f"{p}layer2.{b}.")
        a = _conv_q(sd, a, p + "conv2", p + "bn2", padding=1)
        a = F.max_pool2d(a, 2, (2, 1), (0, 1))
        for b in range(5):
            a = _block_q(sd, a, f"{p}layer3.{b}.")
        a = _conv_q(sd, a, p + "conv3", p + "bn3", padding=1)
        for b in range(3):
            a = _block_q(sd, a, f"{p}layer4.{b}.")
        a = _conv_q(sd, a, p + "conv4_1", p + "bn4_1", stride=(2, 1), padding=(0, 1))
        a = _conv_q(sd, a, p + "conv4_2", p + "bn4_2")
        s = a.permute(0, 3, 1, 2)
        return _bf(s.mean(dim=3)) if s.shape[3] > 1 else s[..., 0].contiguous()


def sequence_bf16emu_ref(sd, feat) -> torch.Tensor:
    """features [B,T,512] -> logits [B,T,C]: 2 x BiLSTM(+Linear) + Prediction with bf16 storage, fp32 accumulate."""
    with torch.no_grad():
        s = feat if isinstance(feat, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(feat))
        s = _bf(s.float())
        for li in range(2):
            pre = f"SequenceModeling.{li}."
            fw = _lstm_dir_q(s, *[_t(sd, pre + "rnn." + n + "_l0") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")], False)
            bw = _lstm_dir_q(s, *[_t(sd, pre + "rnn." + n + "_l0_reverse") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")], True)
            s = _bf(F.linear(torch.cat([fw, bw], 2), _bf(_t(sd, pre + "linear.weight")), _t(sd, pre + "linear.bias")))
        return F.linear(s, _bf(_t(sd, "Prediction.weight")), _t(sd, "Prediction.bias"))


def model_forward_bf16emu_ref(sd, x) -> torch.Tensor:
    """float32 [B,1,H,W] -> logits float32 [B,T,C] with bf16 storage / fp32 accumulation at the kernel's points."""
    return sequence_bf16emu_ref(sd, features_bf16emu_ref(sd, x))


# --------------------------------------------------------------------------------------------------------------
# a6: softmax / mask / renormalise / argmax  (src/ocrfa/utils.py:628-638)
# --------------------------------------------------------------------------------------------------------------
def probs_ref(logits, ignore_idx: Sequence[int]) -> np.ndarray:
    """utils.py:628-632: softmax over classes, zero the ignored classes, divide by the row sum. float32 [B,T,C]."""
    lg = logits if isinstance(logits, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(logits))
    p = F.softmax(lg.float(), dim=2).cpu().numpy()
    p[:, :, list(ignore_idx)] = 0.0
    norm = p.sum(axis=2)
    return (p / np.expand_dims(norm, axis=-1)).astype(np.float32)


# --------------------------------------------------------------------------------------------------------------
# a7: greedy CTC decode  (src/ocrfa/utils.py:1099-1115)
# --------------------------------------------------------------------------------------------------------------
def greedy_labels_ref(idx_row: np.ndarray, ignore: Sequence[int] = (0,)) -> List[int]:
    """Keep t iff idx[t] != idx[t-1] (t=0 always kept) and idx[t] not in `ignore`."""
    out = []
    prev = None
    for t, k in enumerate(idx_row.tolist()):
        if (t == 0 or k != prev) and k not in ignore:
            out.append(k)
        prev = k
    return out


def greedy_decode_ref(probs: np.ndarray, character: Sequence[str], ignore: Sequence[int] = (0,)) -> List[str]:
    """utils.py:637-639 + decode_greedy: argmax over probabilities (lowest index wins ties), collapse, map."""
    idx = probs.argmax(axis=2)
    return ["".join(character[k] for k in greedy_labels_ref(row, ignore)) for row in idx]


# --------------------------------------------------------------------------------------------------------------
# a10: confidence  (src/ocrfa/utils.py:647-660, custom_mean :446-447)
# --------------------------------------------------------------------------------------------------------------
def confidence_ref(probs: np.ndarray) -> List[float]:
    values = probs.max(axis=2)
    indices = probs.argmax(axis=2)
    out = []
    for v, i in zip(values, indices):
        mp = v[i != 0]
        if len(mp) == 0:
            mp = np.array([0])
        out.append(mp.prod() ** (2.0 / np.sqrt(len(mp))))
    return out


# --------------------------------------------------------------------------------------------------------------
# a8: CTC prefix beam search  (src/ocrfa/utils.py:949-1038 with :824-851, :904-946; lm=None, dict_list=[])
# --------------------------------------------------------------------------------------------------------------
def _extend_label(lab: Tuple[int, ...], c: int, blank: int = 0) -> Tuple[int, ...]:
    """fast_simplify_label, utils.py:904-940, for the reachable cases (a labeling never starts with blank)."""
    if not lab:
        return lab if c == blank else (c,)
    if c == blank:
        return lab if lab[-1] == blank else lab + (c,)
    if lab[-1] == blank:
        return lab + (c,) if lab[-2] == c else lab[:-1] + (c,)
    return lab + (c,)


def beam_labels_ref(mat: np.ndarray, beam_width: int, blank: int = 0) -> Tuple[int, ...]:
    """Best raw labeling (blanks and repeats still inside) after the last timestep.

    State per labeling: [prTotal, prNonBlank, prBlank], float32 arithmetic in exactly the reference's order:
    copy-beam first (utils.py:966-990), then every class c with mat[t,c] >= float32(0.5/C) in ascending order,
    blank included (utils.py:997-1017); equal labelings merge by += in dict insertion order; candidates for the
    next step are the top `beam_width` by prTotal under a stable descending sort (utils.py:847-851, :964).
    """
    mat = np.asarray(mat, dtype=np.float32)
    T, C = mat.shape
    thr = np.float32(0.5 / C)
    one, zero = np.float32(1.0), np.float32(0.0)
    last: Dict[Tuple[int, ...], List[np.float32]] = {(): [one, zero, one]}
    for t in range(T):
        curr: Dict[Tuple[int, ...], List[np.float32]] = {}
        order = sorted(last.keys(), key=lambda k: last[k][0], reverse=True)[:beam_width]
        row = mat[t]
        cand = np.nonzero(row >= thr)[0].tolist()
        for lab in order:
            p_tot, p_nb, p_b = last[lab]
            nb = p_nb * row[lab[-1]] if lab else zero
            bl = p_tot * row[blank]
            e = curr.setdefault(lab, [zero, zero, zero])
            e[1] = e[1] + nb
            e[2] = e[2] + bl
            e[0] = e[0] + (bl + nb)
            for c in cand:
                new = _extend_label(lab, c, blank)
                pr = row[c] * (p_b if (lab and lab[-1] == c) else p_tot)
                e2 = curr.setdefault(new, [zero, zero, zero])
                e2[1] = e2[1] + pr
                e2[0] = e2[0] + pr
        last = curr
    return sorted(last.keys(), key=lambda k: last[k][0], reverse=True)[0]


def beam_collapse_ref(lab: Sequence[int], ignore: Sequence[int] = (0,)) -> List[int]:
    """utils.py:1031-1035: drop ignored indices and adjacent duplicates of the RAW labeling."""
    return [l for i, l in enumerate(lab) if l not in ignore and not (i > 0 and lab[i - 1] == lab[i])]


def beam_decode_ref(probs: np.ndarray, character: Sequence[str], beam_width: int = 5,
                    ignore: Sequence[int] = (0,)) -> List[str]:
    """CTCLabelConverter.decode_beamsearch, utils.py:1117-1122."""
    return ["".join(character[k] for k in beam_collapse_ref(beam_labels_ref(m, beam_width), ignore)) for m in probs]


# --------------------------------------------------------------------------------------------------------------
# a9: word beam search, no separators, empty dictionary  (src/ocrfa/utils.py:1124-1151)
# --------------------------------------------------------------------------------------------------------------
def word_spans_ref(argmax_row: np.ndarray, space_idx: int) -> List[List[int]]:
    """Maximal runs of consecutive timesteps whose argmax is not the space class (utils.py:1134-1136)."""
    spans: List[List[int]] = []
    for t, k in enumerate(argmax_row.tolist()):
        if k == space_idx:
            continue
        if spans and spans[-1][-1] == t - 1:
            spans[-1].append(t)
        else:
            spans.append([t])
    return spans


def wordbeam_decode_ref(probs: np.ndarray, character: Sequence[str], beam_width: int = 5,
                        ignore: Sequence[int] = (0,)) -> List[str]:
    space_idx = list(character).index(" ")  # converter.dict[" "] == index in `character` (utils.py:1052-1056)
    am = np.argmax(probs, axis=2)
    out = []
    for i in range(probs.shape[0]):
        words = []
        for span in word_spans_ref(am[i], space_idx):
            sub = probs[i, span, :]
            words.append("".join(character[k] for k in beam_collapse_ref(beam_labels_ref(sub, beam_width), ignore)))
        out.append(" ".join(words))
    return out


# --------------------------------------------------------------------------------------------------------------
# a9, the converter's options: language separators and dictionaries  (src/ocrfa/utils.py:773-820, :853-872, :1058-1081,
# :1152-1172). Inert in the default Farsi setup (separator_list = {}, no dictionary files), restated for completeness.
# --------------------------------------------------------------------------------------------------------------
def converter_ignore_idx_ref(separator_list: Dict[str, Sequence[str]]) -> List[int]:
    """utils.py:1058-1062: blank plus one index per separator character - the separators are ASSUMED to be the first
    characters of the charset (the indices count them, they are not looked up)."""
    n = sum(len(sep) for sep in separator_list.values())
    return [0] + [i + 1 for i in range(n)]


def _run_edges(pos: Sequence[int], mode: str) -> List[int]:
    """consecutive(), utils.py:773-782: first / last element of every run of consecutive integers in `pos`."""
    out: List[int] = []
    for k, v in enumerate(pos):
        starts = k == 0 or v != pos[k - 1] + 1
        ends = k == len(pos) - 1 or pos[k + 1] != v + 1
        if (mode == "first" and starts) or (mode == "last" and ends):
            out.append(int(v))
    return out


def word_segmentation_ref(argmax_row: np.ndarray, separator_idx=None, separator_idx_list=None):
    """word_segmentation, utils.py:785-820 (decode_wordbeamsearch calls it with its DEFAULT tables, :1153): a start
    separator class (odd index) marks with the LAST timestep of each of its runs, an end separator (even index) with the
    FIRST; a start followed by the matching end yields [lang, [t0, t1]] (t1 < t0 when nothing lies between them), text
    outside such pairs yields ['', [t0, t1]]."""
    if separator_idx_list is None:
        separator_idx_list = [1, 2, 3, 4]
    if separator_idx is None:
        separator_idx = {"th": [1, 2], "en": [3, 4]}
    row = np.asarray(argmax_row).tolist()
    marks = []
    for sep in separator_idx_list:
        pos = [t for t, k in enumerate(row) if k == sep]
        marks += [(t, sep) for t in _run_edges(pos, "first" if sep % 2 == 0 else "last")]
    marks.sort(key=lambda m: m[0])  # stable, like sorted(..., key=position)
    result = []
    start_idx = 0
    open_lang, open_at = "", None
    for t, sep in marks:
        for lang, (s_idx, e_idx) in separator_idx.items():
            if sep == s_idx:
                open_lang, open_at = lang, t
            elif sep == e_idx:
                if open_lang == lang:
                    if open_at > start_idx:
                        result.append(["", [start_idx, open_at - 1]])
                    start_idx = t + 1
                    result.append([lang, [open_at + 1, t - 1]])
                open_lang = ""
    if start_idx <= len(row) - 1:
        result.append(["", [start_idx, len(row) - 1]])
    return result


def beam_candidates_ref(mat: np.ndarray, beam_width: int, max_candidate: int = 20, blank: int = 0):
    """The ranked raw labelings of the final beam state (sortedBeams[:maxCandidate] of BeamState.wordsearch,
    utils.py:855-857): beam_labels_ref's search, all final labelings by prTotal under the stable descending sort."""
    mat = np.asarray(mat, dtype=np.float32)
    T, C = mat.shape if mat.ndim == 2 else (0, 0)
    thr = np.float32(0.5 / C) if C else np.float32(0)
    one, zero = np.float32(1.0), np.float32(0.0)
    last: Dict[Tuple[int, ...], List[np.float32]] = {(): [one, zero, one]}
    for t in range(T):
        curr: Dict[Tuple[int, ...], List[np.float32]] = {}
        order = sorted(last.keys(), key=lambda k: last[k][0], reverse=True)[:beam_width]
        row = mat[t]
        cand = np.nonzero(row >= thr)[0].tolist()
        for lab in order:
            p_tot, p_nb, p_b = last[lab]
            nb = p_nb * row[lab[-1]] if lab else zero
            bl = p_tot * row[blank]
            e = curr.setdefault(lab, [zero, zero, zero])
            e[1] = e[1] + nb
            e[2] = e[2] + bl
            e[0] = e[0] + (bl + nb)
            for c in cand:
                new = _extend_label(lab, c, blank)
                pr = row[c] * (p_b if (lab and lab[-1] == c) else p_tot)
                e2 = curr.setdefault(new, [zero, zero, zero])
                e2[1] = e2[1] + pr
                e2[0] = e2[0] + pr
        last = curr
    return sorted(last.keys(), key=lambda k: last[k][0], reverse=True)[:max_candidate]


def wordsearch_ref(mat: np.ndarray, character: Sequence[str], beam_width: int, ignore: Sequence[int],
                   dict_list: Sequence[str]) -> str:
    """ctcBeamSearch's two endings, utils.py:1029-1038: the best labeling when the dictionary is empty, else the first
    of the 20 best candidates whose text is a dictionary word (the best one when none is)."""
    cands = beam_candidates_ref(mat, beam_width, 20 if len(dict_list) else 1)
    texts = ["".join(character[k] for k in beam_collapse_ref(lab, ignore)) for lab in cands]
    if not len(dict_list):
        return texts[0]
    for text in texts:
        if text in dict_list:
            return text
    return texts[0]


def wordbeam_decode_full_ref(probs: np.ndarray, character: Sequence[str], beam_width: int = 5,
                             ignore: Sequence[int] = (0,), separator_list=None, dict_list=None) -> List[str]:
    """decode_wordbeamsearch with every option, utils.py:1124-1173. Without separators `dict_list` is one word list and
    the words of a line are joined with spaces; with separators it maps language -> word list, segments come from
    word_segmentation_ref and are concatenated as they are."""
    separator_list = separator_list or {}
    am = np.argmax(probs, axis=2)
    out = []
    for i in range(probs.shape[0]):
        if len(separator_list) == 0:
            space_idx = list(character).index(" ")
            words = [wordsearch_ref(probs[i, span, :], character, beam_width, ignore, dict_list or [])
                     for span in word_spans_ref(am[i], space_idx)]
            out.append(" ".join(words))
        else:
            text = ""
            for lang, (t0, t1) in word_segmentation_ref(am[i]):
                words = [] if lang == "" else dict_list[lang]
                text += wordsearch_ref(probs[i, t0:t1 + 1, :], character, beam_width, ignore, words)
            out.append(text)
    return out


# --------------------------------------------------------------------------------------------------------------
# recognizer_predict restated end to end  (src/ocrfa/utils.py:605-662)
# --------------------------------------------------------------------------------------------------------------
def recognizer_predict_ref(sd, x: np.ndarray, character: Sequence[str], ignore_idx: Sequence[int],
                           decoder: str = "greedy", beam_width: int = 5, quantize: bool = False):
    """x: float32 [B,1,H,W] (already collated). Returns [[str, conf], ...] like the reference. quantize=True runs the
    LSTM / Linear layers dynamically quantized to int8, as get_recognizer(device="cpu", quantize=True) does."""
    logits = model_forward_int8_ref(sd, x) if quantize else model_forward_ref(sd, x)
    probs = probs_ref(logits, ignore_idx)
    if decoder == "greedy":
        strs = greedy_decode_ref(probs, character)
    elif decoder == "beamsearch":
        strs = beam_decode_ref(probs, character, beam_width)
    elif decoder == "wordbeamsearch":
        strs = wordbeam_decode_ref(probs, character, beam_width)
    else:
        raise ValueError(decoder)
    return [[s, c] for s, c in zip(strs, confidence_ref(probs))]


def ignore_idx_ref(character: str, ignore_char: str) -> List[int]:
    """get_text, utils.py:716-722: 1-based index of the FIRST occurrence of each ignored character."""
    out = []
    for ch in ignore_char:
        k = character.find(ch)
        if k >= 0:
            out.append(k + 1)
    return out


# ------------------------------------------------------------------------------------------------------------------
# get_image_list / compute_ratio_and_resize for axis-aligned boxes (utils.py:1351-1418). The reference's
# cv2.resize(..., interpolation=Image.Resampling.LANCZOS) is OpenCV INTER_LINEAR (the enum's value is 1); OpenCV
# 4.x is a third-party dependency (requirements.txt: opencv-python~=4.9.0.80), so its published 8-bit arithmetic
# (imgproc/resize.cpp: HResizeLinear / VResizeLinear with INTER_RESIZE_COEF_BITS = 11) is restated here and pinned
# by tests/golden/crops.npz, which oracle/gen_golden_crops.py produced with the reference's own get_image_list.
# ------------------------------------------------------------------------------------------------------------------
def _cv_linear_axis(src: int, dst: int, clamp: bool):
    scale = 1.0 / (float(dst) / src)
    idx = np.zeros(dst, np.int64)
    w = np.zeros((dst, 2), np.int64)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(math.floor(f))
        f = np.float32(f - np.float32(s))
        if clamp:
            if s < 0:
                f, s = np.float32(0), 0
            if s >= src - 1:
                f, s = np.float32(0), src - 1
        idx[d] = s
        w[d, 0] = int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048))))
        w[d, 1] = int(np.rint(np.float32(f * np.float32(2048))))
    return idx, w


def cv_resize_linear_u8_ref(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    sh, sw = img.shape
    if (dw, dh) == (sw, sh):
        return img.copy()
    if sw == 2 * dw and sh == 2 * dh:  # exact 2x shrink: resize() switches INTER_LINEAR to INTER_AREA
        s = img.astype(np.int64)
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    xo, xa = _cv_linear_axis(sw, dw, True)
    yo, yb = _cv_linear_axis(sh, dh, False)
    S = img.astype(np.int64)
    x1 = np.minimum(xo + 1, sw - 1)
    H = S[:, xo] * xa[:, 0] + S[:, x1] * xa[:, 1]
    out = np.zeros((dh, dw), np.uint8)
    for d in range(dh):
        r0 = min(max(int(yo[d]), 0), sh - 1)
        r1 = min(max(int(yo[d]) + 1, 0), sh - 1)
        v = ((yb[d, 0] * (H[r0] >> 4)) >> 16) + ((yb[d, 1] * (H[r1] >> 4)) >> 16) + 2
        out[d] = np.clip(v >> 2, 0, 255).astype(np.uint8)
    return out


def _cv_get_perspective_transform_ref(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """cv2.getPerspectiveTransform (OpenCV imgproc/imgwarp.cpp): 8x8 system with float32 cross terms, solved in
    float64 by OpenCV's own LU (core/matrix_decomp.cpp LUImpl: partial pivoting, d = -1/pivot, row updates, back
    substitution)."""
    src = src.astype(np.float32)
    dst = dst.astype(np.float32)
    A = np.zeros((8, 8))
    b = np.zeros(8)
    for i in range(4):
        A[i, 0:3] = (src[i, 0], src[i, 1], 1.0)
        A[i + 4, 3:6] = (src[i, 0], src[i, 1], 1.0)
        A[i, 6], A[i, 7] = np.float32(-src[i, 0] * dst[i, 0]), np.float32(-src[i, 1] * dst[i, 0])
        A[i + 4, 6], A[i + 4, 7] = np.float32(-src[i, 0] * dst[i, 1]), np.float32(-src[i, 1] * dst[i, 1])
        b[i], b[i + 4] = dst[i, 0], dst[i, 1]
    eps = np.finfo(np.float64).eps * 100
    for i in range(8):
        k = i + int(np.argmax(np.abs(A[i:, i])))  # first maximum, like the strict '>' scan
        if abs(A[k, i]) < eps:
            return np.append(np.zeros(8), 1.0).reshape(3, 3)
        if k != i:
            A[[i, k]] = A[[k, i]]
            b[[i, k]] = b[[k, i]]
        d = -1.0 / A[i, i]
        for j in range(i + 1, 8):
            alpha = A[j, i] * d
            for c in range(i + 1, 8):
                A[j, c] += alpha * A[i, c]
            b[j] += alpha * b[i]
    for i in range(7, -1, -1):
        acc = b[i]
        for c in range(i + 1, 8):
            acc -= A[i, c] * b[c]
        b[i] = acc / A[i, i]
    return np.append(b, 1.0).reshape(3, 3)


def _cv_invert3_ref(S: np.ndarray) -> np.ndarray:
    """cv::invert, 3x3 float64 closed form (core/lapack.cpp)."""
    c00 = S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]
    c01 = S[1, 0] * S[2, 2] - S[1, 2] * S[2, 0]
    c02 = S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0]
    det = S[0, 0] * c00 - S[0, 1] * c01 + S[0, 2] * c02
    if det == 0.0:
        return np.zeros((3, 3))
    d = 1.0 / det
    return np.array([[c00 * d, (S[0, 2] * S[2, 1] - S[0, 1] * S[2, 2]) * d, (S[0, 1] * S[1, 2] - S[0, 2] * S[1, 1]) * d],
                     [(S[1, 2] * S[2, 0] - S[1, 0] * S[2, 2]) * d, (S[0, 0] * S[2, 2] - S[0, 2] * S[2, 0]) * d,
                      (S[0, 2] * S[1, 0] - S[0, 0] * S[1, 2]) * d],
                     [c02 * d, (S[0, 1] * S[2, 0] - S[0, 0] * S[2, 1]) * d, (S[0, 0] * S[1, 1] - S[0, 1] * S[1, 0]) * d]])


def cv_warp_perspective_u8_ref(src: np.ndarray, M: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.warpPerspective(src, M, (dw, dh)) - INTER_LINEAR, BORDER_CONSTANT 0 (imgproc/imgwarp.cpp
    WarpPerspectiveInvoker + remapBilinear): float64 inverse map per 32x32-area block, 1/32-pixel fixed point,
    integer bilinear weights summing to 32768."""
    Mi = _cv_invert3_ref(np.asarray(M, dtype=np.float64))
    sh, sw = src.shape
    bh0 = min(16, dh)
    bw0 = min(1024 // bh0, dw)
    ys, xs = np.meshgrid(np.arange(dh), np.arange(dw), indexing="ij")
    xb = (xs // bw0) * bw0
    x1 = (xs - xb).astype(np.float64)
    xbf, yf = xb.astype(np.float64), ys.astype(np.float64)
    X0 = (Mi[0, 0] * xbf + Mi[0, 1] * yf) + Mi[0, 2]
    Y0 = (Mi[1, 0] * xbf + Mi[1, 1] * yf) + Mi[1, 2]
    W0 = (Mi[2, 0] * xbf + Mi[2, 1] * yf) + Mi[2, 2]
    W = W0 + Mi[2, 0] * x1
    with np.errstate(divide="ignore"):
        W = np.where(W != 0.0, 32.0 / W, 0.0)
    fX = np.clip((X0 + Mi[0, 0] * x1) * W, -2147483648.0, 2147483647.0)
    fY = np.clip((Y0 + Mi[1, 0] * x1) * W, -2147483648.0, 2147483647.0)
    X = np.rint(fX).astype(np.int64)
    Y = np.rint(fY).astype(np.int64)
    sx, sy = np.clip(X >> 5, -32768, 32767), np.clip(Y >> 5, -32768, 32767)
    ax, ay = X & 31, Y & 31
    S = np.zeros((sh + 2, sw + 2), np.int64)  # zero border
    S[1:-1, 1:-1] = src

    def px(yy, xx):
        ok = (yy >= 0) & (yy < sh) & (xx >= 0) & (xx < sw)
        return np.where(ok, S[np.clip(yy, -1, sh) + 1, np.clip(xx, -1, sw) + 1], 0)

    acc = (px(sy, sx) * ((32 - ay) * (32 - ax) * 32) + px(sy, sx + 1) * ((32 - ay) * ax * 32) +
           px(sy + 1, sx) * (ay * (32 - ax) * 32) + px(sy + 1, sx + 1) * (ay * ax * 32))
    return np.clip((acc + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


def four_point_transform_ref(image: np.ndarray, rect: np.ndarray) -> np.ndarray:
    """utils.py:1218-1231."""
    (tl, tr, br, bl) = rect
    width_a = np.sqrt(((br[0] - bl[0]) ** 2) + ((br[1] - bl[1]) ** 2))
    width_bs = np.sqrt(((tr[0] - tl[0]) ** 2) + ((tr[1] - tl[1]) ** 2))
    mw = max(int(width_a), int(width_bs))
    height_a = np.sqrt(((tr[0] - br[0]) ** 2) + ((tr[1] - br[1]) ** 2))
    height_bs = np.sqrt(((tl[0] - bl[0]) ** 2) + ((tl[1] - bl[1]) ** 2))
    mh = max(int(height_a), int(height_bs))
    dst = np.array([[0, 0], [mw - 1, 0], [mw - 1, mh - 1], [0, mh - 1]], dtype="float32")
    return cv_warp_perspective_u8_ref(image, _cv_get_perspective_transform_ref(rect, dst), mw, mh)


def _resize_to_model_height_ref(crop: np.ndarray, model_height: int):
    """compute_ratio_and_resize, utils.py:1361-1372."""
    height, width = crop.shape
    ratio = width / height
    if ratio < 1.0:
        ratio = 1.0 / ratio
        return cv_resize_linear_u8_ref(crop, model_height, int(model_height * ratio)), ratio
    return cv_resize_linear_u8_ref(crop, int(model_height * ratio), model_height), ratio


def get_image_list_ref(horizontal_list, img: np.ndarray, model_height: int = 64, sort_output: bool = True, free_list=()):
    """utils.py:1375-1418."""
    maximum_y, maximum_x = img.shape
    image_list = []
    max_ratio_hori, max_ratio_free = 1, 1
    for box in free_list:
        rect = np.array(box, dtype="float32")
        warped = four_point_transform_ref(img, rect)
        h, w = warped.shape
        ratio = w / h
        if ratio < 1.0:
            ratio = 1.0 / ratio
        if int(model_height * ratio) == 0:
            continue
        crop, ratio = _resize_to_model_height_ref(warped, model_height)
        image_list.append((box, crop))
        max_ratio_free = max(ratio, max_ratio_free)
    max_ratio_free = math.ceil(max_ratio_free)
    for box in horizontal_list:
        x_min, x_max = max(0, box[0]), min(box[1], maximum_x)
        y_min, y_max = max(0, box[2]), min(box[3], maximum_y)
        crop = img[y_min:y_max, x_min:x_max]
        width, height = x_max - x_min, y_max - y_min
        ratio = width / height
        if ratio < 1.0:
            ratio = 1.0 / ratio
        if int(model_height * ratio) == 0:
            continue
        crop, ratio = _resize_to_model_height_ref(crop, model_height)
        image_list.append(([[x_min, y_min], [x_max, y_min], [x_max, y_max], [x_min, y_max]], crop))
        max_ratio_hori = max(ratio, max_ratio_hori)
    max_width = math.ceil(max(math.ceil(max_ratio_hori), max_ratio_free)) * model_height
    if sort_output:
        image_list = sorted(image_list, key=lambda item: item[0][0][1])
    return image_list, max_width
