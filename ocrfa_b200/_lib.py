"""ctypes binding of libocrfa_b200.so (include/ocrfa_b200.h). Plumbing only: no arithmetic lives here.

The library is built in-tree by `__graft_entry__.build()` / `make -C ocrfa_b200/csrc`. There is no CPU fallback:
if the shared object is missing, importing the compute entry points raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libocrfa_b200.so")

PRECISION_FP32 = 0
PRECISION_BF16 = 1
PRECISION_F16X3 = 2
PRECISION_F16 = 3
FLAG_ASYNC = 1
PRECISIONS = {"fp32": PRECISION_FP32, "bf16": PRECISION_BF16, "f16x3": PRECISION_F16X3, "f16": PRECISION_F16}

# every symbol include/ocrfa_b200.h declares (tests/test_cabi_symbols.py checks the header against this list)
SYMBOLS = (
    "ocrfa_abi_version", "ocrfa_last_error", "ocrfa_create", "ocrfa_destroy", "ocrfa_load_weight",
    "ocrfa_finalize_weights", "ocrfa_out_timesteps", "ocrfa_workspace_bytes", "ocrfa_preprocess_u8", "ocrfa_align_collate_u8", "ocrfa_crop_resize_u8", "ocrfa_warp_perspective_u8", "ocrfa_rot90_u8",
    "ocrfa_craft_create", "ocrfa_craft_workspace_bytes", "ocrfa_craft_forward",
    "ocrfa_det_preprocess_u8", "ocrfa_bgr2gray_u8", "ocrfa_det_workspace_bytes", "ocrfa_det_components",
    "ocrfa_forward", "ocrfa_forward_features", "ocrfa_forward_sequence", "ocrfa_forward_u8", "ocrfa_ctc_greedy", "ocrfa_ctc_collapse", "ocrfa_argmax_rows",
    "ocrfa_beam_workspace_bytes", "ocrfa_ctc_beam", "ocrfa_ctc_beam_spans", "ocrfa_beam_candidates_workspace_bytes",
    "ocrfa_ctc_beam_candidates", "ocrfa_recognize_host",
    "ocrfa_recognize_device", "ocrfa_recognize_host_ex", "ocrfa_recognize_device_ex", "ocrfa_ctc_margin", "ocrfa_recognize_workspace_bytes", "ocrfa_profile_begin", "ocrfa_profile_end",
    "ocrfa_launch_count",
)


class OcrfaConfig(C.Structure):
    _fields_ = [("input_channel", C.c_int32), ("output_channel", C.c_int32), ("hidden_size", C.c_int32),
                ("num_class", C.c_int32), ("precision", C.c_int32), ("device", C.c_int32)]


class OcrfaError(RuntimeError):
    pass


_lib = None


def load():
    """Loads the shared library once; raises OcrfaError (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OcrfaError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                         "or `make -C ocrfa_b200/csrc` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    lib.ocrfa_abi_version.restype = C.c_int
    lib.ocrfa_last_error.restype = C.c_char_p
    lib.ocrfa_create.argtypes = [C.POINTER(OcrfaConfig), C.POINTER(vp)]
    lib.ocrfa_destroy.argtypes = [vp]
    lib.ocrfa_destroy.restype = None
    lib.ocrfa_load_weight.argtypes = [vp, C.c_char_p, vp, C.POINTER(i64), i32]
    lib.ocrfa_finalize_weights.argtypes = [vp, vp]
    lib.ocrfa_out_timesteps.argtypes = [i32, i32]
    lib.ocrfa_out_timesteps.restype = i32
    lib.ocrfa_workspace_bytes.argtypes = [vp, i32, i32, i32]
    lib.ocrfa_workspace_bytes.restype = sz
    lib.ocrfa_preprocess_u8.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    lib.ocrfa_align_collate_u8.argtypes = [vp, vp, vp, i32, i32, i32, C.c_double, vp, vp, vp, vp]
    lib.ocrfa_forward.argtypes = [vp, vp, i32, i32, i32, vp, vp, sz, vp]
    lib.ocrfa_forward_features.argtypes = [vp, vp, i32, i32, i32, vp, vp, sz, vp]
    lib.ocrfa_forward_sequence.argtypes = [vp, vp, i32, i32, i32, vp, vp, sz, vp]
    lib.ocrfa_forward_u8.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, sz, vp]
    lib.ocrfa_ctc_greedy.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.ocrfa_ctc_collapse.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp]
    lib.ocrfa_argmax_rows.argtypes = [vp, i64, i32, vp, vp]
    lib.ocrfa_beam_workspace_bytes.argtypes = [i32, i32, i32, i32]
    lib.ocrfa_beam_workspace_bytes.restype = sz
    lib.ocrfa_ctc_beam.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, vp, sz, vp]
    lib.ocrfa_ctc_beam_spans.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, sz, vp]
    lib.ocrfa_beam_candidates_workspace_bytes.argtypes = [i32, i32, i32, i32, i32]
    lib.ocrfa_beam_candidates_workspace_bytes.restype = sz
    lib.ocrfa_ctc_beam_candidates.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, sz, vp]
    lib.ocrfa_recognize_host.argtypes = [vp, vp, vp, i32, i32, i32, vp, i32, i32, vp, vp, vp, vp, sz, vp]
    lib.ocrfa_recognize_device.argtypes = [vp, vp, vp, i32, i32, i32, vp, i32, i32, vp, vp, vp, vp, sz, vp]
    lib.ocrfa_recognize_host_ex.argtypes = [vp, vp, vp, i32, i32, i32, vp, i32, i32, vp, vp, vp, vp, C.c_uint32, vp, sz, vp]
    lib.ocrfa_recognize_device_ex.argtypes = [vp, vp, vp, i32, i32, i32, vp, i32, i32, vp, vp, vp, vp, vp, sz, vp]
    lib.ocrfa_ctc_margin.argtypes = [vp, i32, i32, i32, vp, vp, vp]
    lib.ocrfa_profile_begin.argtypes = [vp, C.c_uint32]
    lib.ocrfa_profile_end.argtypes = [vp, vp, vp, vp]
    lib.ocrfa_recognize_workspace_bytes.argtypes = [vp, i32, i32, i32, i32]
    lib.ocrfa_recognize_workspace_bytes.restype = sz
    lib.ocrfa_craft_create.argtypes = [i32, i32, C.POINTER(vp)]
    lib.ocrfa_craft_workspace_bytes.argtypes = [vp, i32, i32, i32]
    lib.ocrfa_craft_workspace_bytes.restype = sz
    lib.ocrfa_craft_forward.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, sz, vp]
    lib.ocrfa_det_preprocess_u8.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, i32, C.POINTER(C.c_float),
                                            C.POINTER(C.c_float), vp, vp]
    lib.ocrfa_bgr2gray_u8.argtypes = [vp, i32, i32, i32, vp, i32, vp]
    lib.ocrfa_det_workspace_bytes.argtypes = [i32, i32, i32]
    lib.ocrfa_det_workspace_bytes.restype = sz
    lib.ocrfa_det_components.argtypes = [vp, vp, i32, i32, i32, C.c_float, C.c_float, C.c_float, vp, vp, i32, vp, i32, vp,
                                         vp, sz, vp]
    lib.ocrfa_rot90_u8.argtypes = [vp, i32, i32, i32, vp, vp]
    lib.ocrfa_launch_count.argtypes = [vp]
    lib.ocrfa_launch_count.restype = i64
    for name in SYMBOLS:  # a missing export is a build error, not something to discover at call time
        getattr(lib, name)
    _lib = lib
    return lib


def resolve_device(device="cuda"):
    """The one place a `device` argument becomes a torch.device: a bare "cuda" means the CURRENT CUDA device
    (what `tensor.to("cuda")` means in torch and hence in the reference, utils.py:621), never a hard-coded cuda:0,
    so the page, crops, detector and recognizer of one call always live on the same GPU."""
    import torch

    dev = device if isinstance(device, torch.device) else torch.device(str(device))
    if dev.type != "cuda":
        raise OcrfaError(f"ocrfa_b200 has no CPU path; got device={device!r}")
    if dev.index is None:
        if not torch.cuda.is_available():
            raise OcrfaError("ocrfa_b200 needs a CUDA device (there is no CPU fallback)")
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def check(rc: int) -> None:
    if rc != 0:
        msg = load().ocrfa_last_error()
        raise OcrfaError(f"ocrfa_b200 error {rc}: {msg.decode('utf-8', 'replace') if msg else ''}")
