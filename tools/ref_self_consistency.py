"""How reproducible is the reference's own fp32 CPU arithmetic?  (CPU only; test infrastructure: imports oracle/)

    python tools/ref_self_consistency.py [--lines 2048] [--configs 8:32,4:32,1:32,8:8] [--out profiles/r02_ref_self_consistency.json]

Each config is threads:batch; the first one is the base every other run is compared with.

Runs the oracle (the same torch CPU fp32 ops the reference executes: F.conv2d / batch_norm / matmul, oneDNN + MKL)
on the same lines and weights with different `torch.set_num_threads` settings and counts how many greedy strings and
per-timestep argmax decisions differ between the runs.  oneDNN / MKL choose their blocking (and hence the fp32
summation order) from the thread count, so the reference's strings are only defined up to that noise: this is the
floor any "identical strings" comparison against a reference run on a different host lives with.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ocrfa_b200.charset import default_character, default_ignore_char  # noqa: E402
from ocrfa_b200.synth import synth_lines, synth_state_dict  # noqa: E402
from oracle import recognizer_ref as R  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lines", type=int, default=2048)
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--configs", default="8:32,4:32")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_ref_self_consistency.json"))
    args = ap.parse_args()
    threads = [tuple(int(v) for v in t.split(":")) for t in args.configs.split(",")]
    sd = {k: torch.from_numpy(v) for k, v in synth_state_dict(0).items()}
    chars = default_character()
    character = ["[blank]"] + list(chars)
    ign = R.ignore_idx_ref(chars, default_ignore_char())
    runs = {}
    t_all = {}
    for cfg in threads:
        nt, batch = cfg
        torch.set_num_threads(nt)
        strs, idxs, lgs = [], [], []
        t0 = time.perf_counter()
        for b0 in range(0, args.lines, batch):
            n = min(batch, args.lines - b0)
            imgs = np.concatenate([synth_lines(1, 32, args.width, seed=900000 + b0 + i) for i in range(n)])
            x = R.align_collate_ref(list(imgs), 32, args.width)
            lg = R.model_forward_ref(sd, x).numpy()
            pr = R.probs_ref(lg, ign)
            strs += R.greedy_decode_ref(pr, character)
            idxs.append(pr.argmax(2))
            lgs.append(lg)
        t_all[cfg] = time.perf_counter() - t0
        runs[cfg] = (strs, np.concatenate(idxs), np.concatenate(lgs))
        print(f"threads={nt} batch={batch}: {args.lines / t_all[cfg]:.1f} lines/s", flush=True)
    base = threads[0]
    out = {"lines": args.lines, "shape": [32, args.width], "torch": torch.__version__,
           "cpu": open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t") if os.path.exists("/proc/cpuinfo") else "?",
           "lines_per_s": {f"{k[0]}:{k[1]}": args.lines / v for k, v in t_all.items()}, "pairs": {}}
    for nt in threads[1:]:
        a, b = runs[base], runs[nt]
        out["pairs"][f"{base[0]}:{base[1]}_vs_{nt[0]}:{nt[1]}"] = {
            "different_strings": int(sum(x != y for x, y in zip(a[0], b[0]))),
            "different_timesteps": int((a[1] != b[1]).sum()), "timesteps": int(a[1].size),
            "max_abs_logit_diff": float(np.abs(a[2] - b[2]).max()),
            "rms_logit_diff": float(np.sqrt(((a[2] - b[2]).astype(np.float64) ** 2).mean())),
        }
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
