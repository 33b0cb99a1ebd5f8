"""One recognize step per iteration, eager launches, for profilers (GPU box):

    ncu --set full --clock-control none --import-source on --launch-skip <launches of the warm-up step> -c <launches of one step> \
        -o gpurun_out/step python tools/profile_step.py --precision f16x3 --config c2 --steps 2

The first step is the warm-up (module load, function attributes); pass its launch count (printed at the end of a plain
run) as --launch-skip to capture exactly the second one. Profiling infrastructure, not part of the product.
"""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OCRFA_GRAPHS", "0")

import ocrfa_b200 as ob  # noqa: E402
from ocrfa_b200 import _lib  # noqa: E402
from ocrfa_b200.recognition import _ignore_mask  # noqa: E402
from ocrfa_b200.synth import synth_lines_global, synth_state_dict  # noqa: E402

CONFIGS = {"c1": (1, 32, 128, 0, 1), "c2": (256, 32, 512, 0, 1), "c3": (64, 64, 2048, 0, 1), "c4": (128, 32, 512, 1, 8)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="f16x3")
    ap.add_argument("--config", default="c2")
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    B, H, W, dec, bw = CONFIGS[args.config]
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    chars = ob.default_character()
    conv = ob.CTCLabelConverter(chars)
    Cn = len(conv.character)
    m = ob.Model(1, 512, 512, Cn, precision=args.precision)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth_state_dict(0).items()})
    handle = m._ensure_handle(dev)
    T = lib.ocrfa_out_timesteps(H, W)
    d_img = torch.from_numpy(synth_lines_global(0, B, H, W, 1234)).to(dev)
    mask = torch.from_numpy(_ignore_mask(Cn, [chars.index(ch) + 1 for ch in ob.default_ignore_char()])).to(dev)
    wsb = lib.ocrfa_recognize_workspace_bytes(handle, B, H, W, bw)
    ws = m.workspace(wsb, dev)
    lab = torch.empty((B, T), dtype=torch.int32, device=dev)
    ln = torch.empty(B, dtype=torch.int32, device=dev)
    cf = torch.empty(B, dtype=torch.float64, device=dev)
    st = torch.cuda.Stream(dev)
    counts = []
    for _ in range(args.steps):
        n0 = lib.ocrfa_launch_count(handle)
        _lib.check(lib.ocrfa_recognize_device(handle, C.c_void_p(d_img.data_ptr()), None, B, H, W, C.c_void_p(mask.data_ptr()),
                                              dec, bw, C.c_void_p(lab.data_ptr()), C.c_void_p(ln.data_ptr()),
                                              C.c_void_p(cf.data_ptr()), C.c_void_p(ws.data_ptr()), wsb, C.c_void_p(st.cuda_stream)))
        torch.cuda.synchronize(dev)
        counts.append(int(lib.ocrfa_launch_count(handle) - n0))
    print("launches per step:", counts)


if __name__ == "__main__":
    main()
