"""Drop-in for the reference recognizer `Model` (/root/reference/src/ocrfa/modules/blocks.py:422-447).

Same constructor, same state-dict keys (FeatureExtraction.ConvNet.*, SequenceModeling.{0,1}.*, Prediction.*),
same call signature `model(image, text) -> logits [B, T, num_class]`; the arithmetic runs in libocrfa_b200.so
(tcgen05 implicit-GEMM convolutions, LSTM step kernel, ...). The torch modules created here only HOLD parameters
so that `load_state_dict` / `state_dict` / `.to()` behave like the reference; they are never executed.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import nn

from . import _lib

_BLOCKS = (1, 2, 5, 3)  # ResNet(input_channel, output_channel, BlockB, [1, 2, 5, 3]), blocks.py:189


def _conv(cin, cout, k):
    return nn.Conv2d(cin, cout, kernel_size=k, bias=False)


def _resnet_holder(cin: int, cout: int) -> nn.Module:
    """Parameter container with the attribute names of the reference ResNet (blocks.py:226-338)."""
    net = nn.Module()
    planes = [cout // 4, cout // 2, cout, cout]
    net.conv0_1, net.bn0_1 = _conv(cin, cout // 16, 3), nn.BatchNorm2d(cout // 16)
    net.conv0_2, net.bn0_2 = _conv(cout // 16, cout // 8, 3), nn.BatchNorm2d(cout // 8)
    inpl = cout // 8
    for li, (pl, nblocks) in enumerate(zip(planes, _BLOCKS), start=1):
        layer = nn.Sequential()
        for b in range(nblocks):
            blk = nn.Module()
            blk.conv1, blk.bn1 = _conv(inpl if b == 0 else pl, pl, 3), nn.BatchNorm2d(pl)
            blk.conv2, blk.bn2 = _conv(pl, pl, 3), nn.BatchNorm2d(pl)
            if b == 0 and inpl != pl:
                blk.downsample = nn.Sequential(_conv(inpl, pl, 1), nn.BatchNorm2d(pl))
            layer.add_module(str(b), blk)
        setattr(net, f"layer{li}", layer)
        inpl = pl
        if li < 4:
            setattr(net, f"conv{li}", _conv(pl, pl, 3))
            setattr(net, f"bn{li}", nn.BatchNorm2d(pl))
    net.conv4_1, net.bn4_1 = _conv(cout, cout, 2), nn.BatchNorm2d(cout)
    net.conv4_2, net.bn4_2 = _conv(cout, cout, 2), nn.BatchNorm2d(cout)
    return net


def _bilstm_holder(cin: int, hidden: int, cout: int) -> nn.Module:
    m = nn.Module()
    m.rnn = nn.LSTM(cin, hidden, bidirectional=True, batch_first=True)
    m.linear = nn.Linear(hidden * 2, cout)
    return m


class Model(nn.Module):
    """ResNet-29 -> mean over H' -> 2 x BiLSTM(+Linear) -> Linear(num_class), executed by the CUDA library.

    precision: "f16x3" (default; tcgen05 tensor cores on fp16 hi/lo split operands, three products per term, hi.hi
    accumulated per K chunk in registers: fp32-grade logits, the mode held to the reference's strings), "bf16" (single-pass
    tensor cores, the BASELINE config's dtype: 3.5x the f16x3 rate, logits within 2e-2 of the reference's range), "f16"
    (the bf16 kernels with fp16 operands / activations: 8x smaller rounding error at the same speed) or "fp32"
    (CUDA-core arithmetic with fp64 accumulation: 25x slower, the reference point of the parity runs).
    """

    def __init__(self, input_channel, output_channel, hidden_size, num_class, precision: str = "f16x3"):
        super().__init__()
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}, got {precision!r}")
        self.precision = precision
        self.num_class = int(num_class)
        self._cfg = (int(input_channel), int(output_channel), int(hidden_size))
        fe = nn.Module()
        fe.ConvNet = _resnet_holder(input_channel, output_channel)
        self.FeatureExtraction = fe
        self.SequenceModeling = nn.Sequential(_bilstm_holder(output_channel, hidden_size, hidden_size),
                                              _bilstm_holder(hidden_size, hidden_size, hidden_size))
        self.Prediction = nn.Linear(hidden_size, num_class)
        self._handle: Optional[C.c_void_p] = None
        self._handle_device: Optional[int] = None
        self._ws: Optional[torch.Tensor] = None
        self.last_margins = None  # decision margins of the last recognize_host batch (float32 [B])
        for p in self.parameters():
            p.requires_grad_(False)
        self.eval()

    # ------------------------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """Accepts checkpoints with or without the DataParallel `module.` prefix (utils.py:680-693)."""
        if state_dict and all(k.startswith("module.") for k in state_dict):
            state_dict = type(state_dict)((k[7:], v) for k, v in state_dict.items())
        out = super().load_state_dict(state_dict, strict=strict, assign=assign)
        self._release()
        return out

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
        """Creates the native handle on first use and hands it the (BN-unfolded) state dict."""
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if self._handle is not None and self._handle_device == idx:
            return self._handle
        self._release()
        lib = _lib.load()
        cfg = _lib.OcrfaConfig(self._cfg[0], self._cfg[1], self._cfg[2], self.num_class,
                               _lib.PRECISIONS[self.precision], idx)
        handle = C.c_void_p()
        _lib.check(lib.ocrfa_create(C.byref(cfg), C.byref(handle)))
        try:
            for name, t in self.state_dict().items():
                if not t.dtype.is_floating_point:
                    continue  # num_batches_tracked
                host = t.detach().to("cpu", torch.float32).contiguous()
                shape = (C.c_int64 * max(1, host.dim()))(*host.shape)
                _lib.check(lib.ocrfa_load_weight(handle, name.encode(), C.c_void_p(host.data_ptr()), shape, host.dim()))
            _lib.check(lib.ocrfa_finalize_weights(handle, None))
        except Exception:
            lib.ocrfa_destroy(handle)
            raise
        self._handle, self._handle_device = handle, idx
        return handle

    def workspace(self, nbytes: int, device: torch.device) -> torch.Tensor:
        """Grow-only device scratch owned by the model (the library never allocates per call)."""
        if self._ws is None or self._ws.numel() < nbytes + 1024 or self._ws.device != device:
            self._ws = torch.empty(int(nbytes) + 1024, dtype=torch.uint8, device=device)
        off = (-self._ws.data_ptr()) % 1024
        return self._ws[off:off + nbytes]

    # ------------------------------------------------------------------------------------------------------
    def forward(self, inputs: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        """inputs: float32 [B, 1, H, W] on a CUDA device (extra args ignored like the reference, blocks.py:440)."""
        if not inputs.is_cuda:
            raise _lib.OcrfaError("ocrfa_b200.Model runs on CUDA only (no CPU fallback); move the input to the GPU")
        if inputs.dim() != 4 or inputs.shape[1] != 1:
            raise ValueError(f"expected [B, 1, H, W], got {tuple(inputs.shape)}")
        x = inputs.detach().to(torch.float32).contiguous()
        B, _, H, W = x.shape
        lib = _lib.load()
        T = lib.ocrfa_out_timesteps(H, W)
        if T == 0:
            raise ValueError(f"unsupported line geometry H={H} W={W} (H multiple of 16 >= 32, W multiple of 4)")
        with torch.cuda.device(x.device):
            handle = self._ensure_handle(x.device)
            ws_bytes = lib.ocrfa_workspace_bytes(handle, B, H, W)
            ws = self.workspace(ws_bytes, x.device)
            logits = torch.empty((B, T, self.num_class), dtype=torch.float32, device=x.device)
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(lib.ocrfa_forward(handle, C.c_void_p(x.data_ptr()), B, H, W, C.c_void_p(logits.data_ptr()),
                                         C.c_void_p(ws.data_ptr()), ws_bytes, C.c_void_p(stream)))
        return logits

    def features(self, inputs: torch.Tensor) -> torch.Tensor:
        """AdaptiveAvgPool(FeatureExtraction(inputs)): float32 [B, T, 512] (blocks.py:443-446)."""
        x = inputs.detach().to(torch.float32).contiguous()
        B, _, H, W = x.shape
        lib = _lib.load()
        T = lib.ocrfa_out_timesteps(H, W)
        if T == 0 or not x.is_cuda:
            raise ValueError("expected a CUDA tensor [B,1,H,W] with H multiple of 16 >= 32 and W multiple of 4")
        with torch.cuda.device(x.device):
            handle = self._ensure_handle(x.device)
            ws_bytes = lib.ocrfa_workspace_bytes(handle, B, H, W)
            ws = self.workspace(ws_bytes, x.device)
            feat = torch.empty((B, T, self._cfg[1]), dtype=torch.float32, device=x.device)
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(lib.ocrfa_forward_features(handle, C.c_void_p(x.data_ptr()), B, H, W, C.c_void_p(feat.data_ptr()),
                                                  C.c_void_p(ws.data_ptr()), ws_bytes, C.c_void_p(stream)))
        return feat

    def sequence(self, features: torch.Tensor, img_h: int, img_w: int) -> torch.Tensor:
        """Prediction(SequenceModeling(features)) for features of lines that were img_h x img_w (blocks.py:441-442)."""
        f = features.detach().to(torch.float32).contiguous()
        B, T, _ = f.shape
        lib = _lib.load()
        if lib.ocrfa_out_timesteps(img_h, img_w) != T or not f.is_cuda:
            raise ValueError("features do not match the stated line geometry")
        with torch.cuda.device(f.device):
            handle = self._ensure_handle(f.device)
            ws_bytes = lib.ocrfa_workspace_bytes(handle, B, img_h, img_w)
            ws = self.workspace(ws_bytes, f.device)
            logits = torch.empty((B, T, self.num_class), dtype=torch.float32, device=f.device)
            stream = torch.cuda.current_stream(f.device).cuda_stream
            _lib.check(lib.ocrfa_forward_sequence(handle, C.c_void_p(f.data_ptr()), B, img_h, img_w,
                                                  C.c_void_p(logits.data_ptr()), C.c_void_p(ws.data_ptr()), ws_bytes,
                                                  C.c_void_p(stream)))
        return logits

    def forward_u8(self, images_u8: torch.Tensor, widths: Optional[torch.Tensor] = None) -> torch.Tensor:
        """uint8 [B, H, W] lines (already at model height) -> logits; the normalisation is fused into the stem."""
        if not images_u8.is_cuda or images_u8.dtype != torch.uint8 or images_u8.dim() != 3:
            raise ValueError("expected a CUDA uint8 tensor [B, H, W]")
        img = images_u8.contiguous()
        B, H, W = img.shape
        lib = _lib.load()
        T = lib.ocrfa_out_timesteps(H, W)
        if T == 0:
            raise ValueError(f"unsupported line geometry H={H} W={W}")
        with torch.cuda.device(img.device):
            handle = self._ensure_handle(img.device)
            ws_bytes = lib.ocrfa_workspace_bytes(handle, B, H, W)
            ws = self.workspace(ws_bytes, img.device)
            logits = torch.empty((B, T, self.num_class), dtype=torch.float32, device=img.device)
            wptr = None
            if widths is not None:
                widths = widths.to(img.device, torch.int32).contiguous()
                wptr = C.c_void_p(widths.data_ptr())
            stream = torch.cuda.current_stream(img.device).cuda_stream
            _lib.check(lib.ocrfa_forward_u8(handle, C.c_void_p(img.data_ptr()), wptr, B, H, W,
                                            C.c_void_p(logits.data_ptr()), C.c_void_p(ws.data_ptr()), ws_bytes,
                                            C.c_void_p(stream)))
        return logits

    def launch_count(self) -> int:
        return int(_lib.load().ocrfa_launch_count(self._handle)) if self._handle is not None else 0
