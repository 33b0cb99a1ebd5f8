"""CPU-only: the C-ABI library loads and exports every symbol include/ocrfa_b200.h declares; compute entry points
refuse to run without a GPU instead of falling back."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "ocrfa_b200.h"), encoding="utf-8").read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ocrfa_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ocrfa_b200 import _lib

    lib = _lib.load()
    declared = _header_symbols()
    assert len(declared) >= 18
    assert sorted(_lib.SYMBOLS) == declared
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.ocrfa_abi_version() == 1


def test_pure_helpers_without_gpu():
    from ocrfa_b200 import _lib

    lib = _lib.load()
    assert lib.ocrfa_out_timesteps(32, 128) == 33
    assert lib.ocrfa_out_timesteps(32, 512) == 129
    assert lib.ocrfa_out_timesteps(64, 2048) == 513
    assert lib.ocrfa_out_timesteps(30, 128) == 0 and lib.ocrfa_out_timesteps(32, 130) == 0
    assert lib.ocrfa_beam_workspace_bytes(128, 129, 185, 8) >= 128 * 2 * 8 * 129 * 2


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import ocrfa_b200 as ob
    from ocrfa_b200 import _lib

    lib = _lib.load()
    cfg = _lib.OcrfaConfig(1, 512, 512, 185, _lib.PRECISION_BF16, 0)
    h = C.c_void_p()
    assert lib.ocrfa_create(C.byref(cfg), C.byref(h)) == -2  # OCRFA_ERR_CUDA
    assert b"CUDA" in lib.ocrfa_last_error()
    m = ob.Model(1, 512, 512, 185)
    with pytest.raises(ob.OcrfaError):
        m(torch.zeros(1, 1, 32, 128), None)
    with pytest.raises(ob.OcrfaError):
        ob.CTCLabelConverter(ob.default_character()).decode_beamsearch(torch.zeros(1, 4, 185).numpy())


def test_model_state_dict_keys_match_reference_layout(synth_sd):
    import ocrfa_b200 as ob

    m = ob.Model(1, 512, 512, 185)
    keys = set(m.state_dict().keys())
    assert keys == set(synth_sd.keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(synth_sd[k].shape), k
    m.load_state_dict({"module." + k: torch.from_numpy(v) for k, v in synth_sd.items()})


def test_charset_tables_match_reference(charset):
    import ocrfa_b200 as ob

    assert ob.default_character() == str(charset["character"])
    assert ob.ARABIC_G1_SYMBOLS == str(charset["symbols"])
    assert ob.default_lang_char() == str(charset["lang_char"])
    assert ob.default_ignore_char() == str(charset["ignore_char"])
    conv = ob.CTCLabelConverter(ob.default_character())
    assert conv.character == charset["converter_character"].tolist()
    assert conv.ignore_idx == charset["converter_ignore_idx"].tolist()
    assert len(conv.character) == int(charset["num_class"]) == 185
