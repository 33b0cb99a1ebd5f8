"""Generates tests/golden/craft_64x96.npz with the UNMODIFIED reference detector network
(/root/reference/src/ocrfa/utils.py:48-98 class OCRFA) loaded with the synthetic weights of
ocrfa_b200.craft.synth_craft_state_dict(0). TEST INFRASTRUCTURE, dev container only (needs /root/reference and
torchvision):  python oracle/gen_golden_craft.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle._refimport import import_reference  # noqa: E402
from ocrfa_b200.craft import synth_craft_state_dict  # noqa: E402


def main():
    utils, _ = import_reference()
    torch.manual_seed(0)
    net = utils.OCRFA(pretrained=False)
    sd = {k: torch.from_numpy(v) for k, v in synth_craft_state_dict(0).items()}
    missing, unexpected = net.load_state_dict(sd, strict=False)
    missing = [m for m in missing if "num_batches_tracked" not in m]
    assert not missing and not unexpected, (missing, unexpected)
    net.eval()
    rng = np.random.default_rng(3)
    out = {}
    for name, (n, h, w) in {"a": (2, 64, 96), "b": (1, 96, 64)}.items():
        x = rng.standard_normal((n, 3, h, w)).astype(np.float32)
        with torch.no_grad():
            y, feat = net(torch.from_numpy(x))
        out[f"x_{name}"], out[f"y_{name}"], out[f"feature_{name}"] = x, y.numpy(), feat.numpy()
        print(name, x.shape, y.shape, feat.shape, float(np.abs(y.numpy()).max()), float(np.abs(feat.numpy()).max()))
    path = os.path.join(ROOT, "tests", "golden", "craft_64x96.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
