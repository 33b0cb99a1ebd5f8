#!/usr/bin/env python
"""BASELINE config 5: N synthetic lines of 32x512 (default 1,000,000) sharded contiguously over the ranks.

    python tools/sweep_1m.py [--lines 1000000]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep_1m.py

Every rank generates ITS shard on its own GPU. The set is defined in global chunks of 8192 lines - chunk c comes from
torch.Generator(device).manual_seed(1234 + c): light background, dark vertical strokes, noise -
so a line is a function of its GLOBAL index and any sharding reads the same lines: the label checksum of a 1-GPU and an
8-GPU run of the same set must be equal. Each rank recognises its shard in inner batches of 256 through
ocrfa_recognize_device (device-resident inputs, CUDA events, no collective on the path), and only then the decoded
labels are all-gathered over NCCL (ocrfa_b200.distributed.gather_decoded) and the slowest rank's device time is taken.
Rank 0 prints one JSON line. Not a unit of the product: a driver for the throughput sweep.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import ocrfa_b200 as ob  # noqa: E402
from ocrfa_b200 import _lib  # noqa: E402
from ocrfa_b200.distributed import gather_decoded, labels_to_strings, max_over_ranks, shard_range  # noqa: E402
from ocrfa_b200.recognition import _ignore_mask  # noqa: E402
from ocrfa_b200.synth import synth_state_dict  # noqa: E402


def synth_lines_device(n: int, h: int, w: int, gen: torch.Generator, dev: torch.device, chunk: int = 8192) -> torch.Tensor:
    """uint8 [n, h, w] on the device: background U(200,255), ~24 dark full-height strokes of width 2..12 per line at
    random columns, N(0,6) noise."""
    out = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
    cols = torch.arange(w, device=dev).view(1, 1, w)
    for a in range(0, n, chunk):
        m = min(chunk, n - a)
        img = torch.empty((m, 1, w), device=dev).uniform_(200.0, 255.0, generator=gen)[:, :, :1].expand(m, 1, w).clone()
        x0 = torch.randint(0, w - 12, (m, 24, 1), device=dev, generator=gen)
        sw = torch.randint(2, 13, (m, 24, 1), device=dev, generator=gen)
        ink = torch.empty((m, 24, 1), device=dev).uniform_(0.0, 80.0, generator=gen)
        hit = (cols >= x0) & (cols < x0 + sw)                       # [m, 24, w]
        dark = torch.where(hit, ink, torch.full_like(ink, 1e9)).amin(dim=1, keepdim=True)  # [m, 1, w]
        img = torch.minimum(img, dark)
        full = img.expand(m, h, w) + torch.randn((m, h, w), device=dev, generator=gen) * 6.0
        out[a:a + m] = full.round().clamp_(0, 255).to(torch.uint8)
    return out


GLOBAL_CHUNK = 8192


def synth_lines_device_global(lo: int, hi: int, h: int, w: int, dev: torch.device) -> torch.Tensor:
    """Lines lo..hi-1 of the global set (see the module docstring)."""
    out = torch.empty((max(hi - lo, 0), h, w), dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev)
    for c in range(lo // GLOBAL_CHUNK, (max(hi, lo + 1) - 1) // GLOBAL_CHUNK + 1):
        gen.manual_seed(1234 + c)
        chunk = synth_lines_device(GLOBAL_CHUNK, h, w, gen, dev, chunk=GLOBAL_CHUNK)
        a, b = max(lo, c * GLOBAL_CHUNK), min(hi, (c + 1) * GLOBAL_CHUNK)
        if b > a:
            out[a - lo:b - lo] = chunk[a - c * GLOBAL_CHUNK:b - c * GLOBAL_CHUNK]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lines", type=int, default=1_000_000)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--precision", default="f16x3", choices=["f16x3", "bf16", "f16", "fp32"])
    args = ap.parse_args()
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/ocrfa_nccl_%h_%p.log")
        dist.init_process_group(backend="nccl", device_id=dev)
    H, W, B = 32, 512, args.batch
    lib = _lib.load()
    chars = ob.default_character()
    conv = ob.CTCLabelConverter(chars)
    Cn = len(conv.character)
    model = ob.Model(1, 512, 512, Cn, precision=args.precision)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in synth_state_dict(0).items()})
    handle = model._ensure_handle(dev)
    T = lib.ocrfa_out_timesteps(H, W)
    lo, hi = shard_range(args.lines, rank, world)
    n_local = hi - lo
    lines = synth_lines_device_global(lo, max(hi, lo + 1), H, W, dev)
    d_mask = torch.from_numpy(_ignore_mask(Cn, [chars.index(ch) + 1 for ch in ob.default_ignore_char()])).to(dev)
    ws_bytes = lib.ocrfa_recognize_workspace_bytes(handle, B, H, W, 1)
    ws = model.workspace(ws_bytes, dev)
    # fixed staging buffers: one captured CUDA graph serves every full batch (the graph bakes pointers in)
    stage = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    d_lab = torch.empty((B, T), dtype=torch.int32, device=dev)
    d_len = torch.empty(B, dtype=torch.int32, device=dev)
    d_conf = torch.empty(B, dtype=torch.float64, device=dev)
    labs = torch.zeros((max(n_local, 1), T), dtype=torch.int32, device=dev)
    lens = torch.zeros(max(n_local, 1), dtype=torch.int32, device=dev)
    confs = torch.zeros(max(n_local, 1), dtype=torch.float64, device=dev)
    stream = torch.cuda.Stream(dev)  # a capturable stream: the step replays as a CUDA graph
    torch.cuda.set_stream(stream)

    def run(a: int, b: int):
        nb = b - a
        stage[:nb].copy_(lines[a:b])
        _lib.check(lib.ocrfa_recognize_device(handle, C.c_void_p(stage.data_ptr()), None, nb, H, W,
                                              C.c_void_p(d_mask.data_ptr()), 0, 1, C.c_void_p(d_lab.data_ptr()),
                                              C.c_void_p(d_len.data_ptr()), C.c_void_p(d_conf.data_ptr()),
                                              C.c_void_p(ws.data_ptr()), ws_bytes, C.c_void_p(stream.cuda_stream)))
        labs[a:b].copy_(d_lab[:nb])
        lens[a:b].copy_(d_len[:nb])
        confs[a:b].copy_(d_conf[:nb])

    for _ in range(3):
        run(0, min(B, n_local))
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for a in range(0, n_local, B):
        run(a, min(a + B, n_local))
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = max_over_ranks(e0.elapsed_time(e1), dev)
    g_lab, g_len, g_conf = gather_decoded(labs[:n_local], lens[:n_local], confs[:n_local], args.lines)
    if rank == 0:
        sample = labels_to_strings(g_lab[:3], g_len[:3], conv.character)
        checksum = int((g_lab.to(torch.int64) * (torch.arange(T, device=g_lab.device) < g_len[:, None])).sum().item())
        print(json.dumps({"metric": "text lines/sec (device-timed, max over ranks)", "value": args.lines / (ms * 1e-3),
                          "unit": "lines/s", "n_gpus": world, "lines": args.lines, "ms_total": ms, "scaling": "strong",
                          "dtype": args.precision, "data": "synthetic (generated on device; line g is a function of its global index)",
                          "config": {"workload": f"c5: {args.lines} lines of {H}x{W}, contiguous shards, inner batch {B}, greedy"},
                          "gathered_lines": int(g_lab.shape[0]), "label_checksum": checksum, "sample": sample}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
