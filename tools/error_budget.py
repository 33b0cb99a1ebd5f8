"""Where does each precision mode's logit error come from?  (GPU box; test infrastructure: imports oracle/)

    python tools/error_budget.py [--lines 32] [--width 512] [--precisions fp32,f16x3] [--out gpurun_out/error_budget.json]

Truth = the oracle's forward evaluated in float64 on the CPU.  Reported per mode, for the visual features
(ResNet + mean over H') and for the logits: rms and max |x - truth| relative to rms(truth).  The fp32 oracle itself
("ref32": what the reference computes, torch CPU fp32) is listed beside the CUDA modes, and every CUDA mode is
also compared with ref32 directly (that distance is what flips an argmax against the reference).  `head:*` rows feed
the SAME fp32 features (ref32's) to each mode's sequence head, isolating the LSTM / Linear / Prediction error.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import ocrfa_b200 as ob  # noqa: E402
from ocrfa_b200.synth import synth_lines, synth_state_dict  # noqa: E402
from oracle import recognizer_ref as R  # noqa: E402


def forward64(sd64, x64):
    with torch.no_grad():
        f = R.resnet_ref(sd64, x64)
        v = f.permute(0, 3, 1, 2).mean(dim=3)
        s = R.bilstm_ref(sd64, v, "SequenceModeling.0.")
        s = R.bilstm_ref(sd64, s, "SequenceModeling.1.")
        return F.linear(s, sd64["Prediction.weight"], sd64["Prediction.bias"]), v


def head64(sd64, v):
    with torch.no_grad():
        s = R.bilstm_ref(sd64, v, "SequenceModeling.0.")
        s = R.bilstm_ref(sd64, s, "SequenceModeling.1.")
        return F.linear(s, sd64["Prediction.weight"], sd64["Prediction.bias"])


def stats(x, truth):
    d = np.asarray(x, np.float64) - truth
    scale = float(np.sqrt((truth ** 2).mean()))
    # shrink: the least-squares delta of x = (1 + delta) * truth; rms_rel_unshrunk: what is left without that component
    delta = float((d * truth).sum() / (truth ** 2).sum())
    rest = d - delta * truth
    return {"rms_rel": float(np.sqrt((d ** 2).mean())) / scale, "max_rel": float(np.abs(d).max()) / scale,
            "mean_rel": float(d.mean()) / scale, "shrink": delta,
            "rms_rel_unshrunk": float(np.sqrt((rest ** 2).mean())) / scale}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lines", type=int, default=32)
    ap.add_argument("--height", type=int, default=32)
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--precisions", default="fp32,f16x3")
    ap.add_argument("--comps", default="", help="comma list of OCRFA_X3_COMP values to scan in the f16x3 mode")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "error_budget.json"))
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    sd_np = synth_state_dict(0)
    sd32 = {k: torch.from_numpy(v) for k, v in sd_np.items()}
    sd64 = {k: v.double() for k, v in sd32.items()}
    imgs = synth_lines(args.lines, args.height, args.width, seed=424242)
    x = R.align_collate_ref(list(imgs), args.height, args.width)
    lg64, ft64 = forward64(sd64, torch.from_numpy(x).double())
    lg64, ft64 = lg64.numpy(), ft64.numpy()
    lg32, ft32 = R.model_forward_ref(sd32, x, return_features=True)
    lg32, ft32 = lg32.numpy(), ft32.numpy()
    hd64 = head64(sd64, torch.from_numpy(ft32).double()).numpy()  # exact head on ref32's features
    hd32 = head64(sd32, torch.from_numpy(ft32)).numpy()
    out = {"lines": args.lines, "shape": [args.height, args.width],
           "rms_truth": {"features": float(np.sqrt((ft64 ** 2).mean())), "logits": float(np.sqrt((lg64 ** 2).mean()))},
           "ref32": {"features": stats(ft32, ft64), "logits": stats(lg32, lg64), "head": stats(hd32, hd64)}}
    dev = torch.device("cuda", 0)
    xd = torch.from_numpy(x).to(dev)
    fd = torch.from_numpy(ft32).to(dev)
    runs = [(p, None) for p in args.precisions.split(",")] + [("f16x3", c) for c in args.comps.split(",") if c]
    for prec, comp in runs:
        if comp is not None:
            cx, _, cl = comp.partition("/")  # "8e-8" or "8e-8/6e-8": igemm_x3 / LSTM recurrence
            os.environ["OCRFA_X3_COMP"] = cx
            os.environ["OCRFA_LSTM_COMP"] = cl or cx
        m = ob.Model(1, 512, 512, 185, precision=prec)
        m.load_state_dict(sd32)
        ft = m.features(xd).cpu().numpy()
        lg = m(xd, None).cpu().numpy()
        hd = m.sequence(fd, args.height, args.width).cpu().numpy()
        out[prec if comp is None else "f16x3@" + comp] = {"features": stats(ft, ft64), "logits": stats(lg, lg64), "head": stats(hd, hd64),
                     "features_vs_ref32": stats(ft, ft32.astype(np.float64)),
                     "logits_vs_ref32": stats(lg, lg32.astype(np.float64)),
                     "head_vs_ref32": stats(hd, hd32.astype(np.float64))}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
