"""Evidence for DESIGN.md: how much of the bf16-vs-fp32 logit gap is inherent to bf16 storage on this network.

CPU only. Evaluates the oracle three ways on a golden batch: fp32; fp32 with ONLY the conv-stack features rounded
to bf16; full bf16 emulation (oracle.recognizer_ref.model_forward_bf16emu_ref). Prints the np.allclose-style scaled
error max|d|/(1+|ref|) and the range-normalised error max|d|/max|ref| of each against fp32.
    python tools/bf16_sensitivity.py
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import recognizer_ref as R  # noqa: E402
from oracle.synth import synth_state_dict  # noqa: E402


def main():
    sd = synth_state_dict(0)
    for name in ("c1_32x128", "c2_32x512"):
        g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", name + ".npz"))
        with torch.no_grad():
            ref, v = R.model_forward_ref(sd, g["x"], return_features=True)
            s = R.bilstm_ref(sd, v.bfloat16().float(), "SequenceModeling.0.")
            s = R.bilstm_ref(sd, s, "SequenceModeling.1.")
            feat_only = F.linear(s, R._t(sd, "Prediction.weight"), R._t(sd, "Prediction.bias"))
            emu = R.model_forward_bf16emu_ref(sd, g["x"])
        for label, y in (("features rounded to bf16, everything else fp32", feat_only), ("full bf16 emulation", emu)):
            d = (y - ref).abs()
            print(f"{name}: {label}: scaled {float((d / (1 + ref.abs())).max()):.3g}, "
                  f"range-normalised {float(d.max() / ref.abs().max()):.3g}, "
                  f"argmax agreement {float((y.argmax(2) == ref.argmax(2)).float().mean()):.4f}")


if __name__ == "__main__":
    main()
