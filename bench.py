#!/usr/bin/env python
"""bench.py - text lines/sec of the OCR-FA recognition hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config c2|c1|c3|c4] [--precision f16x3|bf16|f16|fp32]

A step = one pass of the hot path (uint8 lines -> fused normalise+stem -> ResNet-29 -> 2xBiLSTM -> Linear -> softmax/
mask/renorm/argmax/CTC collapse/confidence) over one batch of synthetic lines per rank. Default workload is BASELINE
config 2 (256 lines of 32x512, greedy) in the DEFAULT precision of the library, f16x3: fp32-grade tensor-core
arithmetic, the mode whose decoded strings are held to the reference's (the `parity` object of the line, measured in
this run against the oracle). The BASELINE config's bf16 arithmetic is timed in the same run and reported under
`other_precisions` (and is the headline of `--precision bf16`). Ranks are independent (weak scaling): rank r reads lines
[r*B, (r+1)*B) of a synthetic set in which a line is a function of its global index; NCCL is used only after the timed
regions, to gather timings and decoded labels (rank 0 then recomputes every rank's lines itself and compares).

JSON line keys: value = whole-job lines/s with inputs resident in HBM (CUDA events on the launching stream, max over
ranks); e2e = the same through ocrfa_recognize_host_ex with pinned HOST buffers, H2D + D2H inside the timed region,
two batches in flight on two streams; roofline = the tcgen05 implicit-GEMM kernels, algorithmic FLOPs / CUDA-event
kernel time in the timed region vs MEASURED_PEAKS.json; cpu_baseline = the oracle (port of the reference's CPU path)
on a bounded sample, rank 0, N=1 only. `--impl reference` times that CPU path alone: the reference's OWN
recognizer_predict when a copy of its package is present (/root/reference in the dev container; oracle/_ref/, git-ignored
build output of __graft_entry__.build(), on the GPU box), else the oracle port pinned against it in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "text lines/sec (device-timed, max over ranks)"
UNIT = "lines/s"
CONFIGS = {
    # name: (batch, H, W, decoder, beam)
    "c1": (1, 32, 128, "greedy", 1),
    "c2": (256, 32, 512, "greedy", 1),
    "c3": (64, 64, 2048, "greedy", 1),
    "c4": (128, 32, 512, "beamsearch", 8),
}
# algorithmic FLOPs per line (2*MAC of every conv/LSTM/Linear), SURVEY.md section 8d / BASELINE.md section 3
FLOPS_PER_LINE = {(32, 128): 13.453e9, (32, 512): 52.788e9, (64, 2048): 411.546e9}
DTYPE_TEXT = {
    "f16x3": "f16x3 (fp16 hi/lo split operands on tcgen05, 3 products per MAC, hi.hi summed per K chunk in fp32 registers: fp32-grade)",
    "bf16": "bf16", "f16": "f16 (fp16 operands/activations, fp32 accumulate: the bf16 kernels, 3 more mantissa bits)", "fp32": "f32",
}
LINE_SEED = 1234  # line g of the synthetic set is synth_lines(seed = LINE_SEED + g)


def workload_text(cfg_name: str) -> str:
    B, H, W, decoder, beam = CONFIGS[cfg_name]
    return (f"{cfg_name}: {B} lines of {H}x{W} per rank per step, {decoder}" + (f" beam={beam}" if decoder != "greedy" else "")
            + ", reference default model (ResNet-29 + 2xBiLSTM + Linear185), random-init weights")


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_sustained": float(p.get("bf16_tflops_sustained", 1354.4)), "bf16_burst": float(p.get("bf16_tflops", 1654.9)),
                "hbm": float(p.get("hbm_gbs", 6550.4)), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def load_traffic(config: str, precision: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of the igemm launches of ONE step, per launch (mean), from the
    committed `ncu --set full` capture of this bench command (profiles/*_traffic.json, written by tools/ncu_summary.py);
    None when no capture exists for this config/precision."""
    path = os.path.join(ROOT, "profiles", f"igemm_traffic_{config}_{precision}.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f).get("dram_bytes_per_launch")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local_rank, world


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's CPU path (recognizer_predict on torch CPU fp32)
# ------------------------------------------------------------------------------------------------------------------
def cpu_lines_per_s(cfg_name: str, sample_lines: int, repeats: int, warmup: int, quantize: bool = False):
    """lines/s of the reference's CPU path on a bounded sample. When a copy of the reference package is present
    (/root/reference here, oracle/_ref/ on the GPU box: see __graft_entry__.build) the reference's OWN get_recognizer +
    recognizer_predict run (kind "reference"); otherwise the oracle port of the same calls (kind "port")."""
    from oracle import recognizer_ref as R
    from oracle import _refimport
    from ocrfa_b200.synth import synth_lines, synth_state_dict
    import ocrfa_b200.charset as cs

    B, H, W, decoder, beam = CONFIGS[cfg_name]
    n = min(sample_lines, B)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd_np = synth_state_dict(0)
    character = cs.default_character()
    ign = R.ignore_idx_ref(character, cs.default_ignore_char())
    imgs = synth_lines(n, H, W, seed=777)
    bw = beam if decoder != "greedy" else 5
    kind = "port"
    one_pass = None
    if _refimport.reference_available():
        try:
            import tempfile
            from collections import OrderedDict
            from PIL import Image

            utils, _ = _refimport.import_reference()
            with tempfile.TemporaryDirectory() as td:
                path = os.path.join(td, "synthetic.pth")
                torch.save(OrderedDict(("module." + k, torch.from_numpy(v)) for k, v in sd_np.items()), path)
                model, converter = utils.get_recognizer("generation1", {"input_channel": 1, "output_channel": 512, "hidden_size": 512},
                                                        character, {}, {}, path, device="cpu", quantize=quantize)
            model.eval()
            collate = utils.AlignCollate(imgH=H, imgW=W, keep_ratio_with_pad=True)
            pil = [Image.fromarray(im, "L") for im in imgs]

            def one_pass():  # AlignCollate + recognizer_predict, utils.py:567-662
                x = collate(pil)
                return utils.recognizer_predict(model, converter, [x], int(W / 10), ign, {}, decoder, bw, "cpu")

            kind = "reference"
        except Exception as exc:  # a broken copy must not take the bench down: time the port instead
            print(f"[bench] reference import failed ({exc!r}); timing the oracle port", file=sys.stderr)
            one_pass = None
    if one_pass is None:
        sd = {k: torch.from_numpy(v) for k, v in sd_np.items()}
        chars = ["[blank]"] + list(character)

        def one_pass():
            x = R.align_collate_ref(list(imgs), H, W)
            return R.recognizer_predict_ref(sd, x, chars, ign, decoder, bw, quantize=quantize)

    for _ in range(warmup):
        one_pass()
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        one_pass()
        times.append(time.perf_counter() - t0)
    return n / (sum(times) / len(times)), threads, n, sum(times) / len(times), kind


def run_reference_arm(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    B, H, W, decoder, beam = CONFIGS[args.config]
    sample = min(B, 32 if (H, W) == (32, 512) else (4 if H == 64 else B))
    lps, threads, n, sec, kind = cpu_lines_per_s(args.config, sample, max(1, args.steps), max(1, min(args.warmup, 1)))
    what = ("the reference's own get_recognizer + AlignCollate + recognizer_predict (src/ocrfa/utils.py:567-695), copied to oracle/_ref by build()"
            if kind == "reference" else "oracle port of recognizer_predict; no copy of the Python reference on this box")
    line = {
        "impl": "reference", "metric": METRIC, "value": lps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args.config), "sample": f"{n} lines per step on the host CPU"},
        "cpu_baseline": {"value": lps, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{n} lines of {H}x{W} per step, torch CPU fp32, {threads} threads ({what})"},
        "e2e": {"value": lps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
class Slot:
    """One in-flight batch of the end-to-end path: its own stream, workspace and pinned result buffers."""

    def __init__(self, lib, handle, B, H, W, T, bw, dev):
        self.stream = torch.cuda.Stream(dev)
        self.ws_bytes = lib.ocrfa_recognize_workspace_bytes(handle, B, H, W, bw)
        buf = torch.empty(int(self.ws_bytes) + 1024, dtype=torch.uint8, device=dev)
        off = (-buf.data_ptr()) % 1024
        self.ws = buf[off:off + self.ws_bytes]
        self.labels = torch.empty((B, T), dtype=torch.int32).pin_memory()
        self.lens = torch.empty(B, dtype=torch.int32).pin_memory()
        self.conf = torch.empty(B, dtype=torch.float64).pin_memory()
        self.margin = torch.empty(B, dtype=torch.float32).pin_memory()


def measure_device(lib, _lib, handle, geom, d_img, d_mask, dec, bw, ws, ws_bytes, outs, stream, steps, warmup, dev, world,
                   profile_mask, flush_l2):
    """K steps on device-resident inputs, CUDA events on the launching stream. Returns (ms, per-class ms, flops, launches)."""
    B, H, W, T = geom
    d_labels, d_len, d_conf = outs

    def step():
        _lib.check(lib.ocrfa_recognize_device(handle, C.c_void_p(d_img.data_ptr()), None, B, H, W,
                                              C.c_void_p(d_mask.data_ptr()), dec, bw, C.c_void_p(d_labels.data_ptr()),
                                              C.c_void_p(d_len.data_ptr()), C.c_void_p(d_conf.data_ptr()),
                                              C.c_void_p(ws.data_ptr()), ws_bytes, C.c_void_p(stream.cuda_stream)))

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(warmup):
        step()
    barrier()
    launches0 = lib.ocrfa_launch_count(handle)
    _lib.check(lib.ocrfa_profile_begin(handle, profile_mask))
    barrier()
    if flush_l2:
        flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        with torch.cuda.stream(stream):
            for i, (a, b) in enumerate(pairs):
                flush_buf.fill_(i & 0xff)
                a.record(stream)
                step()
                b.record(stream)
        barrier()
        dev_ms = sum(a.elapsed_time(b) for a, b in pairs)
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        barrier()
        dev_ms = e0.elapsed_time(e1)
    ms_cls = (C.c_double * 8)()
    fl_cls = (C.c_double * 8)()
    n_cls = (C.c_int64 * 8)()
    _lib.check(lib.ocrfa_profile_end(handle, ms_cls, fl_cls, n_cls))
    launches = int(lib.ocrfa_launch_count(handle) - launches0)
    return dev_ms, list(ms_cls), list(fl_cls), list(n_cls), launches, step


def other_configs(args, lib, _lib, handle, d_mask, dev, synth_lines_global):
    """Device-timed lines/s of the configurations other than args.config on the same handle. A failure is recorded in
    the row instead of costing the run its headline line."""
    rows = {}
    for name in sorted(CONFIGS):
        if name == args.config:
            continue
        B, H, W, decoder, beam = CONFIGS[name]
        try:
            T = lib.ocrfa_out_timesteps(H, W)
            bw = beam if decoder == "beamsearch" else 1
            dec = 1 if decoder == "beamsearch" else 0
            d_img = torch.from_numpy(synth_lines_global(0, B, H, W, LINE_SEED)).to(dev)
            slot = Slot(lib, handle, B, H, W, T, bw, dev)
            outs = (torch.empty((B, T), dtype=torch.int32, device=dev), torch.empty(B, dtype=torch.int32, device=dev),
                    torch.empty(B, dtype=torch.float64, device=dev))
            flush_l2 = int(slot.ws_bytes) + 220 * (1 << 20) < 512 * (1 << 20)
            k = min(args.steps, 10)
            ms, _, _, _, launches, _ = measure_device(lib, _lib, handle, (B, H, W, T), d_img, d_mask, dec, bw, slot.ws,
                                                      slot.ws_bytes, outs, slot.stream, k, 3, dev, 1, 0, flush_l2)
            lens = outs[1].cpu().numpy().astype(np.int64)
            labels = outs[0].cpu().numpy().astype(np.int64)
            valid = np.arange(T)[None, :] < lens[:, None]
            fl = FLOPS_PER_LINE.get((H, W))
            rows[name] = {"workload": workload_text(name), "value": B * k / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / k,
                          "steps": k, "warmup": 3, "gpu_launches": launches,
                          "l2": "256 MB fill before every timed step" if flush_l2 else "working set exceeds L2",
                          "whole_step_tflops": (fl * B / (ms / k * 1e-3) / 1e12) if fl else None,
                          "label_checksum": int((labels * valid).sum() % (1 << 31))}
            del d_img, slot, outs
            torch.cuda.empty_cache()
        except Exception as exc:  # noqa: BLE001
            rows[name] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    return rows


def run_gpu_arm(args):
    import ocrfa_b200 as ob
    from ocrfa_b200 import _lib
    from ocrfa_b200.recognition import CERTIFY_MARGIN, _ignore_mask
    from ocrfa_b200.synth import synth_lines_global, synth_state_dict  # synthetic data generator (not timed)

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        # stdout carries exactly one JSON line: NCCL's own banner / debug output goes to a file
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/ocrfa_nccl_%h_%p.log")
        dist.init_process_group(backend="nccl", device_id=dev)

    B, H, W, decoder, beam = CONFIGS[args.config]
    lib = _lib.load()
    chars = ob.default_character()
    conv = ob.CTCLabelConverter(chars)
    Cn = len(conv.character)
    sd = {k: torch.from_numpy(v) for k, v in synth_state_dict(0).items()}
    model = ob.Model(1, 512, 512, Cn, precision=args.precision)
    model.load_state_dict(sd)
    handle = model._ensure_handle(dev)
    T = lib.ocrfa_out_timesteps(H, W)

    # this rank's lines: [rank * B, (rank + 1) * B) of the global synthetic set
    imgs_np = synth_lines_global(rank * B, (rank + 1) * B, H, W, LINE_SEED)
    h_img = torch.from_numpy(imgs_np).pin_memory()
    d_img = h_img.to(dev)
    ign = [chars.index(ch) + 1 for ch in ob.default_ignore_char()]
    mask_np = _ignore_mask(Cn, ign)
    d_mask = torch.from_numpy(mask_np).to(dev)
    bw = beam if decoder == "beamsearch" else 1
    dec = 1 if decoder == "beamsearch" else 0
    slots = [Slot(lib, handle, B, H, W, T, bw, dev) for _ in range(2)]
    ws, ws_bytes, stream = slots[0].ws, slots[0].ws_bytes, slots[0].stream  # a non-default stream: CUDA-graph replay is possible
    d_labels = torch.empty((B, T), dtype=torch.int32, device=dev)
    d_len = torch.empty(B, dtype=torch.int32, device=dev)
    d_conf = torch.empty(B, dtype=torch.float64, device=dev)
    # A step whose working set (activation workspace + weights) fits the 126 MB L2 would be timed on a warm cache: such
    # configurations (c1) get an explicit flush - a 256 MB fill - before every timed step, and the step is bracketed by
    # its own event pair so that the flush is not in the sum. The big configurations stream > 1 GB per step.
    flush_l2 = int(ws_bytes) + 220 * (1 << 20) < 512 * (1 << 20)
    warmup = max(3, args.warmup)

    # ---- timed region 1: device-resident inputs; the tcgen05 implicit-GEMM class is bracketed by events (eager launches)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms, ms_cls, fl_cls, n_cls, launches, step_device = measure_device(
        lib, _lib, handle, (B, H, W, T), d_img, d_mask, dec, bw, ws, ws_bytes, (d_labels, d_len, d_conf), stream, args.steps,
        warmup, dev, world, 1 << 1, flush_l2)

    # ---- timed region 2: end to end through the host-buffer C-ABI call, two batches in flight (two streams, two
    # workspaces, pinned host buffers): H2D of batch i+1 and D2H of batch i-1 overlap the kernels of batch i
    def submit(slot):
        _lib.check(lib.ocrfa_recognize_host_ex(
            handle, C.c_void_p(h_img.data_ptr()), None, B, H, W, mask_np.ctypes.data_as(C.c_void_p), dec, bw,
            C.c_void_p(slot.labels.data_ptr()), C.c_void_p(slot.lens.data_ptr()), C.c_void_p(slot.conf.data_ptr()),
            C.c_void_p(slot.margin.data_ptr()), _lib.FLAG_ASYNC, C.c_void_p(slot.ws.data_ptr()), slot.ws_bytes,
            C.c_void_p(slot.stream.cuda_stream)))

    for s in slots:  # warm-up: first sight of each slot's pointers captures its CUDA graph
        submit(s)
        submit(s)
    torch.cuda.synchronize(dev)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    if flush_l2:
        flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        host_s = 0.0
        for i in range(args.steps):
            flush_buf.fill_(i & 0xff)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            submit(slots[0])
            slots[0].stream.synchronize()
            host_s += time.perf_counter() - t0
    else:
        t0 = time.perf_counter()
        for i in range(args.steps):
            s = slots[i & 1]
            s.stream.synchronize()  # the batch that used this slot two submissions ago has been read back
            submit(s)
        torch.cuda.synchronize(dev)
        host_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    e2e_labels = slots[(args.steps - 1) & 1].labels.clone() if not flush_l2 else slots[0].labels.clone()
    e2e_lens = slots[(args.steps - 1) & 1].lens.clone() if not flush_l2 else slots[0].lens.clone()
    margins = (slots[(args.steps - 1) & 1].margin if not flush_l2 else slots[0].margin).numpy().copy()

    # ---- extra pass (not timed): per-class kernel time shares for the record
    _lib.check(lib.ocrfa_profile_begin(handle, 0xff))
    step_device()
    ms_all = (C.c_double * 8)()
    fl_all = (C.c_double * 8)()
    n_all = (C.c_int64 * 8)()
    _lib.check(lib.ocrfa_profile_end(handle, ms_all, fl_all, n_all))
    torch.cuda.synchronize(dev)
    same_paths = bool(torch.equal(d_labels.cpu() * (torch.arange(T) < d_len.cpu()[:, None]),
                                  e2e_labels * (torch.arange(T) < e2e_lens[:, None])))

    # ---- max over ranks; gather the decoded labels (NCCL only after the timed regions)
    times = torch.tensor([dev_ms, host_s * 1e3], dtype=torch.float64, device=dev)
    valid = torch.arange(T, device=dev) < d_len[:, None]
    lab_masked = torch.where(valid, d_labels, torch.zeros_like(d_labels))
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros_like(lab_masked) for _ in range(world)]
        dist.all_gather(gathered, lab_masked)
        all_labels = torch.cat(gathered).cpu()
    else:
        all_labels = lab_masked.cpu()
    dev_ms_max, host_ms_max = float(times[0].item()), float(times[1].item())

    if rank == 0:
        peaks = load_peaks()
        total_lines = B * args.steps * world
        value = total_lines / (dev_ms_max * 1e-3)
        e2e_value = total_lines / (host_ms_max * 1e-3)
        tc_ms, tc_flops, tc_n = float(ms_cls[1]), float(fl_cls[1]), int(n_cls[1])
        achieved = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
        names = ["stem", "igemm_tc", "lstm_tc", "pool_mean_cvt", "conv_simt", "lstm_cell_f32", "ctc_greedy", "ctc_beam"]
        tot_all = sum(float(v) for v in ms_all) or 1.0
        shares = {n: {"ms_per_step": round(float(ms_all[i]), 4), "launches": int(n_all[i]),
                      "share": round(float(ms_all[i]) / tot_all, 4)} for i, n in enumerate(names) if n_all[i]}
        flops_line = FLOPS_PER_LINE.get((H, W))
        products = 3 if args.precision == "f16x3" else 1

        # multi-GPU equivalence (the reference's analogue is DataParallel's gather,
        # utils.py:692): rank 0 recomputes every rank's lines itself and compares with the gathered labels
        equivalence = None
        if world > 1:
            mism = 0
            for r in range(1, world):
                other = torch.from_numpy(synth_lines_global(r * B, (r + 1) * B, H, W, LINE_SEED)).to(dev)
                _lib.check(lib.ocrfa_recognize_device(handle, C.c_void_p(other.data_ptr()), None, B, H, W,
                                                      C.c_void_p(d_mask.data_ptr()), dec, bw, C.c_void_p(d_labels.data_ptr()),
                                                      C.c_void_p(d_len.data_ptr()), C.c_void_p(d_conf.data_ptr()),
                                                      C.c_void_p(ws.data_ptr()), ws_bytes, C.c_void_p(stream.cuda_stream)))
                torch.cuda.synchronize(dev)
                v2 = torch.arange(T, device=dev) < d_len[:, None]
                mine = torch.where(v2, d_labels, torch.zeros_like(d_labels)).cpu()
                mism += int((mine != all_labels[r * B:(r + 1) * B]).any(dim=1).sum())
            equivalence = {"lines": world * B, "ranks": world, "lines_differing_from_single_gpu": mism,
                           "how": "rank 0 recomputed the lines of every other rank and compared them with the all-gathered labels"}
        label_checksum = int((all_labels.to(torch.int64) * (1 + torch.arange(T))).sum().item())

        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE_TEXT[args.precision], "data": "synthetic",
            "config": {"workload": workload_text(args.config),
                       "lines": f"rank r reads lines [r*{B}, (r+1)*{B}) of a synthetic set whose line g is synth_lines(seed={LINE_SEED}+g)",
                       "l2": ("explicit flush: a 256 MB fill before every timed step (the working set of this configuration "
                              "would fit the 126 MB L2), each step bracketed by its own CUDA event pair") if flush_l2 else
                             (f"per-step working set ({int(ws_bytes) / 2**30:.1f} GB activation workspace streamed every step + "
                              "the weights) exceeds the 126 MB L2; no explicit flush"),
                       "timing": "value: eager launches with CUDA events around every tcgen05 GEMM launch (the roofline's kernel "
                                 "time comes from the same region); e2e: CUDA-graph replay, two batches in flight",
                       "parallelism": f"dp{world} (independent ranks, no collective on the path)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(B * H * W + Cn),
                    "d2h_bytes_per_step": int(B * T * 4 + B * 4 + B * 8 + B * 4), "ms_per_step": host_ms_max / args.steps,
                    "in_flight": 1 if flush_l2 else 2, "same_labels_as_device_path": same_paths},
            "gpu_launches": launches,
            "roofline": {"kernel": ("igemm_x3_kernel (tcgen05 cta_group::2 implicit GEMM on fp16 hi/lo split operands: 28 convs, 4 of them "
                                    "halo-staged and 2 of those with the following max pool fused, + the LSTM input projections and the "
                                    "3 linear GEMMs)" if args.precision == "f16x3" else
                                    "igemm_tc_kernel<Pair> + igemm_halo_kernel (tcgen05 implicit GEMM, cta_group::2: 28 convs, 2 of them with the "
                                    "following max pool fused, + the 3 linear GEMMs; the LSTM input projection is fused into lstm_tc)"),
                         "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["bf16_sustained"], "traffic": load_traffic(args.config, args.precision),
                         "launches_per_step": tc_n / max(1, args.steps), "kernel_ms_per_step": tc_ms / max(1, args.steps),
                         "peak_source": peaks["source"] + ", sustained bf16 (cuBLAS 8192^3 back to back); burst "
                                        f"{peaks['bf16_burst']:.0f} TFLOP/s -> frac {achieved / peaks['bf16_burst']:.3f}",
                         "products_per_mac": products,
                         "executed_tflops": achieved * products,
                         "frac_of_mode_ceiling": achieved * products / peaks["bf16_sustained"],
                         "note": ("achieved counts ALGORITHMIC FLOPs (2 per MAC of the reference's layers). The f16x3 arithmetic spends three "
                                  "fp16 tensor-core products per MAC (hi.hi + lo.hi + hi.lo), so its ceiling is one third of the peak: "
                                  "frac_of_mode_ceiling = executed_tflops / peak. " if products == 3 else "") +
                                 "A fraction can read above 1.0 of the sustained peak: that denominator is power-capped at ~1.2 GHz, "
                                 "these launches alternate with latency-bound kernels and run at the SM clock in `clocks`",
                         "whole_step_tflops": (flops_line * B / (dev_ms_max / args.steps * 1e-3) / 1e12) if flops_line else None},
            "kernel_shares": shares,
            "clocks": clocks,
            "label_checksum": label_checksum,
            "certificates": {"margin_bound": CERTIFY_MARGIN, "certified_lines": int((margins >= CERTIFY_MARGIN).sum()), "lines": int(B),
                             "meaning": "lines whose smallest top-2 logit gap exceeds the bound: every fp32-grade evaluation "
                                        "decodes them to the same greedy string (ocrfa_ctc_margin)"} if dec == 0 else None,
        }
        if equivalence is not None:
            line["multi_gpu_equivalence"] = equivalence
        if world == 1 and not args.no_cpu_baseline:
            # ---- parity of THIS precision, measured in this run: the oracle's greedy strings on a bounded sample
            from oracle import recognizer_ref as R
            n_par = min(B, args.parity_lines)
            torch.set_num_threads(os.cpu_count() or 1)
            x = R.align_collate_ref(list(imgs_np[:n_par]), H, W)
            ref_logits = R.model_forward_ref(sd, x)
            ign_ref = R.ignore_idx_ref(chars, ob.default_ignore_char())
            ref_probs = R.probs_ref(ref_logits, ign_ref)
            if decoder == "greedy":
                ref_strs = R.greedy_decode_ref(ref_probs, conv.character)
            else:
                ref_strs = R.beam_decode_ref(ref_probs, conv.character, beam)
            got = conv.labels_to_text(e2e_labels.numpy()[:n_par], e2e_lens.numpy()[:n_par])
            same = sum(a == b for a, b in zip(got, ref_strs))
            cert_ok = [a == b for a, b, m in zip(got, ref_strs, margins[:n_par]) if m >= CERTIFY_MARGIN]
            line["parity"] = {"mode": args.precision, "decoder": decoder, "lines": n_par, "identical_strings": same,
                              "certified_lines": len(cert_ok), "identical_among_certified": int(sum(cert_ok)),
                              "against": "oracle (CPU port of the reference, torch fp32) on the first lines of this run's batch; "
                                         "the >= 10k-line runs are profiles/r02_parity_10k_*.json"}
            sample = 32 if (H, W) == (32, 512) else (4 if H == 64 else min(B, 32))
            lps, threads, n, sec, kind = cpu_lines_per_s(args.config, sample, 2, 1)
            line["cpu_baseline"] = {"value": lps, "unit": UNIT, "cores": threads, "kind": kind,
                                    "sample": f"{n} lines of {H}x{W}, best-effort mean of 2 passes after 1 warm-up, torch CPU fp32 "
                                              f"{threads} threads ({'the reference itself' if kind == 'reference' else 'oracle port'}: "
                                              f"recognizer_predict, {decoder})"}
            # the reference's literal CPU default, OCR(gpu=False, quantize=True): dynamic int8 LSTM / Linear layers
            q_lps, _, _, _, _ = cpu_lines_per_s(args.config, sample, 2, 1, quantize=True)
            line["cpu_baseline"]["quantize_true"] = {"value": q_lps, "unit": UNIT,
                                                     "note": "same sample with torch dynamic int8 LSTM/Linear (utils.py:686-690)"}
        if world == 1 and not args.no_other_precisions:
            # ---- the other tensor-core precisions on the same workload, same run (a few steps each, device-resident inputs)
            others = {}
            for prec in ("bf16", "f16", "f16x3"):
                if prec == args.precision:
                    continue
                m2 = ob.Model(1, 512, 512, Cn, precision=prec)
                m2.load_state_dict(sd)
                h2 = m2._ensure_handle(dev)
                wsb2 = lib.ocrfa_recognize_workspace_bytes(h2, B, H, W, bw)
                if wsb2 > ws_bytes:
                    buf = torch.empty(int(wsb2) + 1024, dtype=torch.uint8, device=dev)
                    ws2 = buf[(-buf.data_ptr()) % 1024:][:wsb2]
                else:
                    ws2 = ws[:wsb2]
                ms2, _, _, _, _, _ = measure_device(lib, _lib, h2, (B, H, W, T), d_img, d_mask, dec, bw, ws2, wsb2,
                                                    (d_labels, d_len, d_conf), stream, min(args.steps, 10), 3, dev, 1, 0, flush_l2)
                k2 = min(args.steps, 10)
                others[prec] = {"value": B * k2 / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / k2, "steps": k2,
                                "dtype": DTYPE_TEXT[prec]}
                del m2
            line["other_precisions"] = others
        if world == 1 and not args.no_other_configs and args.precision != "fp32":
            # ---- BASELINE.json's other single-GPU configurations in the same run and precision (a few steps each,
            # device-resident inputs, the same timing rules): informative rows, the headline stays `value`
            line["other_configs"] = other_configs(args, lib, _lib, handle, d_mask, dev, synth_lines_global)
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--precision", default="f16x3", choices=["f16x3", "bf16", "f16", "fp32"],
                    help="f16x3 (default): tcgen05 fp16 hi/lo split, fp32-grade logits, the mode held to the reference's strings; "
                         "bf16: the BASELINE config's dtype; f16: the bf16 kernels with fp16 operands; fp32: CUDA-core parity mode")
    ap.add_argument("--parity-lines", type=int, default=128, help="lines of the in-run parity check against the oracle")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU legs (parity sample and cpu_baseline)")
    ap.add_argument("--no-other-precisions", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the c1 / c3 / c4 rows of the default run")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
