"""Host mirror of the reference's CRAFT text detector network (the forward only).

    net = CraftNet(precision="f16"); net.load_state_dict(torch.load("craft_mlt_25k.pth")); y, feature = net(x)

mirrors `OCRFA.forward` (/root/reference/src/ocrfa/utils.py:48-98; backbone `Vgg16BN`, modules/blocks.py:27-82) as
`detectors.get_detector` / `test_net` use it (/root/reference/src/ocrfa/detectors.py:43-47, 76-92): x float32
[N,3,H,W] on a CUDA device, H and W multiples of 32 (resize_aspect_ratio pads to that, utils.py:513-517) ->
(y float32 [N,H/2,W/2,2], feature float32 [N,32,H/2,W/2]). The 25 convolutions run in libocrfa_b200.so
(ocrfa_craft_forward); the reference's resize / box post-processing (cv2 on the host) is not part of this package.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib

# (state-dict prefix, cin, cout, kernel, has BatchNorm) in forward order
CRAFT_CONVS = (
    ("basenet.slice1.0", 3, 64, 3, "basenet.slice1.1"), ("basenet.slice1.3", 64, 64, 3, "basenet.slice1.4"),
    ("basenet.slice1.7", 64, 128, 3, "basenet.slice1.8"), ("basenet.slice1.10", 128, 128, 3, "basenet.slice1.11"),
    ("basenet.slice2.14", 128, 256, 3, "basenet.slice2.15"), ("basenet.slice2.17", 256, 256, 3, "basenet.slice2.18"),
    ("basenet.slice3.20", 256, 256, 3, "basenet.slice3.21"), ("basenet.slice3.24", 256, 512, 3, "basenet.slice3.25"),
    ("basenet.slice3.27", 512, 512, 3, "basenet.slice3.28"), ("basenet.slice4.30", 512, 512, 3, "basenet.slice4.31"),
    ("basenet.slice4.34", 512, 512, 3, "basenet.slice4.35"), ("basenet.slice4.37", 512, 512, 3, "basenet.slice4.38"),
    ("basenet.slice5.1", 512, 1024, 3, None), ("basenet.slice5.2", 1024, 1024, 1, None),
    ("upconv1.conv.0", 1536, 512, 1, "upconv1.conv.1"), ("upconv1.conv.3", 512, 256, 3, "upconv1.conv.4"),
    ("upconv2.conv.0", 768, 256, 1, "upconv2.conv.1"), ("upconv2.conv.3", 256, 128, 3, "upconv2.conv.4"),
    ("upconv3.conv.0", 384, 128, 1, "upconv3.conv.1"), ("upconv3.conv.3", 128, 64, 3, "upconv3.conv.4"),
    ("upconv4.conv.0", 192, 64, 1, "upconv4.conv.1"), ("upconv4.conv.3", 64, 32, 3, "upconv4.conv.4"),
    ("conv_cls.0", 32, 32, 3, None), ("conv_cls.2", 32, 32, 3, None), ("conv_cls.4", 32, 16, 3, None),
    ("conv_cls.6", 16, 16, 1, None), ("conv_cls.8", 16, 2, 1, None),
)


def synth_craft_state_dict(seed: int = 0) -> "OrderedDict[str, np.ndarray]":
    """Seed-fixed synthetic detector weights with the reference's state-dict keys and shapes (the real
    craft_mlt_25k.pth is a network download): He-normal convolutions, small biases, non-trivial BatchNorm statistics."""
    rng = np.random.Generator(np.random.PCG64(seed + 77))
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for key, cin, cout, k, bn in CRAFT_CONVS:
        fan_in = cin * k * k
        sd[key + ".weight"] = (rng.standard_normal((cout, cin, k, k)) * np.sqrt(2.0 / fan_in)).astype(np.float32)
        sd[key + ".bias"] = (rng.standard_normal(cout) * 0.05).astype(np.float32)
        if bn:
            sd[bn + ".weight"] = rng.uniform(0.5, 1.0, cout).astype(np.float32)
            sd[bn + ".bias"] = (rng.standard_normal(cout) * 0.1).astype(np.float32)
            sd[bn + ".running_mean"] = (rng.standard_normal(cout) * 0.1).astype(np.float32)
            sd[bn + ".running_var"] = rng.uniform(0.8, 1.2, cout).astype(np.float32)
    return sd


class CraftNet:
    """The reference's `OCRFA` detector network, forward only, on the B200 kernels. precision: "f16" (default) | "bf16"
    (single-pass tensor-core modes) | "f16x3" (fp32-grade: fp16 hi/lo split operands, the recognizer's igemm_x3 kernel;
    score maps within 1e-5 of the reference's fp32 output, so the thresholded components - and hence the boxes - are
    the reference's)."""

    def __init__(self, precision: str = "f16"):
        if precision not in ("bf16", "f16", "f16x3"):
            raise ValueError("the CRAFT forward runs in the tensor-core modes only: 'f16', 'bf16' or 'f16x3'")
        self.precision = precision
        self._state: Optional[Dict[str, torch.Tensor]] = None
        self._handle: Optional[C.c_void_p] = None
        self._handle_device: Optional[int] = None
        self._ws: Optional[torch.Tensor] = None

    def load_state_dict(self, state_dict, strict: bool = True):
        """Accepts the reference's checkpoints with or without the DataParallel `module.` prefix (detectors.py:12-21)."""
        sd = {}
        for k, v in state_dict.items():
            k = k[7:] if k.startswith("module.") else k
            t = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))
            if t.dtype.is_floating_point:
                sd[k] = t.detach().to("cpu", torch.float32).contiguous()
        if strict:
            for key, _, _, _, bn in CRAFT_CONVS:
                need = [key + ".weight", key + ".bias"] + ([bn + s for s in (".weight", ".bias", ".running_mean", ".running_var")] if bn else [])
                missing = [n for n in need if n not in sd]
                if missing:
                    raise KeyError(f"missing detector weights: {missing[:4]}")
        self._state = sd
        self._release()
        return self

    def eval(self):
        return self

    def to(self, *_args, **_kwargs):
        return self

    def _release(self):
        if self._handle is not None:
            _lib.load().ocrfa_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _ensure_handle(self, device: torch.device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if self._handle is not None and self._handle_device == idx:
            return self._handle
        if self._state is None:
            raise _lib.OcrfaError("load_state_dict first")
        self._release()
        lib = _lib.load()
        handle = C.c_void_p()
        _lib.check(lib.ocrfa_craft_create(_lib.PRECISIONS[self.precision], idx, C.byref(handle)))
        try:
            for name, t in self._state.items():
                shape = (C.c_int64 * max(1, t.dim()))(*t.shape)
                _lib.check(lib.ocrfa_load_weight(handle, name.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()))
            _lib.check(lib.ocrfa_finalize_weights(handle, None))
        except Exception:
            lib.ocrfa_destroy(handle)
            raise
        self._handle, self._handle_device = handle, idx
        return handle

    def __call__(self, x: torch.Tensor):
        """OCRFA.forward: (y [N,H/2,W/2,2], feature [N,32,H/2,W/2]), both float32 on x's device."""
        if not x.is_cuda:
            raise _lib.OcrfaError("ocrfa_b200.CraftNet runs on CUDA only (no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] % 32 or x.shape[3] % 32:
            raise ValueError(f"expected [N, 3, H, W] with H, W multiples of 32, got {tuple(x.shape)}")
        x = x.detach().to(torch.float32).contiguous()
        N, _, H, W = x.shape
        lib = _lib.load()
        with torch.cuda.device(x.device):
            handle = self._ensure_handle(x.device)
            ws_bytes = lib.ocrfa_craft_workspace_bytes(handle, N, H, W)
            if self._ws is None or self._ws.numel() < ws_bytes + 1024 or self._ws.device != x.device:
                self._ws = torch.empty(int(ws_bytes) + 1024, dtype=torch.uint8, device=x.device)
            off = (-self._ws.data_ptr()) % 1024
            y = torch.empty((N, H // 2, W // 2, 2), dtype=torch.float32, device=x.device)
            feat = torch.empty((N, 32, H // 2, W // 2), dtype=torch.float32, device=x.device)
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(lib.ocrfa_craft_forward(handle, C.c_void_p(x.data_ptr()), N, H, W, C.c_void_p(y.data_ptr()),
                                               C.c_void_p(feat.data_ptr()), C.c_void_p(self._ws.data_ptr() + off),
                                               ws_bytes, C.c_void_p(stream)))
        return y, feat

    forward = __call__


def get_detector(trained_model, device="cuda", quantize=True, cudnn_benchmark=False, precision="f16"):
    """detectors.py:76-92: builds the network and loads `trained_model` (a .pth path or a state dict)."""
    if str(device) == "cpu":
        raise _lib.OcrfaError("ocrfa_b200 has no CPU path; use device='cuda'")
    net = CraftNet(precision=precision)
    state = trained_model if isinstance(trained_model, dict) else torch.load(trained_model, map_location="cpu")
    net.load_state_dict(state)
    return net
