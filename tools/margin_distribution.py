"""Would a two-tier pass pay (single-pass f16 everywhere, f16x3 only for the lines f16 cannot certify)? CPU only.

The oracle's fp32 logits of N synthetic lines of 32x512 give, per timestep, the gap between the two largest unmasked
logits; a line is decided by its SMALLEST gap. A tier-1 result can be kept only if that gap exceeds twice the tier-1
logit error bound (either logit may move by the bound). The f16 bound comes from the measured 10 240-line run
(profiles/r01_parity_10k_32x512_all_modes.json: max |error| = 0.021 of max |logit|); the script prints the fraction of
lines above a range of bounds, i.e. how many lines each tier-1 accuracy could certify.

    python tools/margin_distribution.py [N] > profiles/r02_margin_distribution.json
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import recognizer_ref as R  # noqa: E402
from oracle.synth import synth_lines, synth_state_dict  # noqa: E402
import ocrfa_b200.charset as cs  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth_state_dict(0)
    chars = cs.default_character()
    ign = R.ignore_idx_ref(chars, cs.default_ignore_char())
    gaps = []
    scale = 0.0
    for i in range(0, n, 32):
        imgs = synth_lines(min(32, n - i), 32, 512, seed=777 + i)
        logits = R.model_forward_ref(sd, R.align_collate_ref(list(imgs), 32, 512)).numpy().astype(np.float64)
        scale = max(scale, float(np.abs(logits).max()))
        logits[:, :, ign] = -np.inf
        top2 = np.sort(logits, axis=2)[:, :, -2:]
        gaps.append(top2[:, :, 1] - top2[:, :, 0])
    gap = np.concatenate(gaps)            # [n, T] absolute logit units
    line_gap = gap.min(axis=1)
    out = {"lines": int(n), "timesteps": int(gap.shape[1]), "max_abs_logit": scale,
           "timestep_gap_quantiles": {q: float(np.quantile(gap, float(q))) for q in ("0.01", "0.05", "0.25", "0.5")},
           "line_min_gap_quantiles": {q: float(np.quantile(line_gap, float(q))) for q in ("0.05", "0.5", "0.95")},
           "tiers": {}}
    # error bound of a logit, as a fraction of max |logit|: f16x3's certificate, f16 / bf16 maxima of the 10 240-line runs,
    # and what a tier 1 would need for the two-tier pass to reach 10 k lines/s (> 60 % of the lines certified)
    for name, bound in (("f16x3 (measured max 2.2e-5)", 2.2e-5), ("1e-4", 1e-4), ("1e-3", 1e-3),
                        ("f16 (measured max 2.1e-2)", 2.1e-2), ("bf16 (measured max 0.16)", 0.16)):
        need = 2.0 * bound * scale
        out["tiers"][name] = {"required_gap_abs": need,
                              "timesteps_above": float((gap > need).mean()),
                              "lines_certifiable": float((line_gap > need).mean())}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
