"""GPU tests added in round 2 (run with `pytest -m gpu` on the B200 box, everything through the C ABI):

  * the >= 1 024-line string gate of the default (f16x3) precision against the oracle, with the decision-margin certificate
  * ocrfa_ctc_margin against numpy
  * the async / extended host entry point against the synchronous one
  * multi-rank equivalence on hardware: two processes shard the lines, gather over gloo, equal the single-rank strings
  * the stress run of the persistent LSTM's cross-CTA publish (tools/test_igemm.cu --lstm-stress)
  * the halo-staged / per-tap f16x3 kernel harness (tools/test_igemm.cu --split)
"""
import ctypes as C
import os
import socket
import subprocess

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def env(synth_sd, charset):
    import ocrfa_b200 as ob
    from ocrfa_b200.recognition import _ignore_mask

    assert torch.cuda.is_available()
    chars = str(charset["character"])
    conv = ob.CTCLabelConverter(chars)
    ign = charset["ignore_idx"].tolist()
    return dict(ob=ob, conv=conv, chars=chars, ign=ign, mask=_ignore_mask(len(conv.character), ign),
                sd={k: torch.from_numpy(v) for k, v in synth_sd.items()}, dev=torch.device("cuda", 0))


def _model(env, prec):
    m = env["ob"].Model(1, 512, 512, len(env["conv"].character), precision=prec)
    m.load_state_dict(env["sd"])
    return m


def test_default_precision_is_the_parity_mode(env):
    ob = env["ob"]
    assert ob.Model(1, 512, 512, 185).precision == "f16x3"
    import inspect
    assert inspect.signature(ob.get_recognizer).parameters["precision"].default == "f16x3"


def test_f16x3_strings_against_the_oracle_on_1024_config2_lines(env, synth_sd):
    """north_star's correctness target on the default precision, inside the driver-run suite: 1 024 lines of the BASELINE
    config-2 shape (32x512, 129 timesteps each = 132 096 argmax decisions) decoded by the f16x3 tensor-core path
    through ocrfa_recognize_host_ex against the oracle's greedy strings.

    Asserted: (1) every CERTIFIED line - decision margin >= CERTIFY_MARGIN, i.e. no timestep whose top-2 logits are
    closer than 3e-5 of the row's logit rms - decodes identically (a hard property: the f16x3 logit error is 1.2e-6 of
    the rms, tools/error_budget.py); (2) at least 90 % of the lines are certified; (3) at most ONE string differs in
    total, and only on an uncertified line: below the bound the reference's own fp32 result moves when nothing but its
    batch size changes (profiles/r02_ref_self_consistency.json), so 'identical to the reference' is not defined there
    beyond that noise. Measured on B200: 2 048 / 2 048 identical, 95.8 % certified."""
    from ocrfa_b200.recognition import CERTIFY_MARGIN
    from ocrfa_b200.synth import synth_lines
    from oracle import recognizer_ref as R

    ob = env["ob"]
    m = _model(env, "f16x3")
    torch.set_num_threads(os.cpu_count() or 1)
    n_same = n_cert = n_bad_cert = n_total = 0
    flips = steps = 0
    for b0 in range(0, 1024, 256):
        imgs = synth_lines(256, 32, 512, seed=900000 + b0)
        x = R.align_collate_ref(list(imgs), 32, 512)
        ref_probs = R.probs_ref(R.model_forward_ref(synth_sd, x), env["ign"])
        ref_strs = R.greedy_decode_ref(ref_probs, env["conv"].character)
        batch = ob.AlignCollate(32, 512)(list(imgs))
        strs, conf, margin = ob.recognize_host(m, env["conv"], batch.images_u8, batch.widths, env["mask"], "greedy", 5,
                                               env["dev"], return_margin=True)
        logits = m.forward_u8(batch.images_u8.to(env["dev"])).cpu().numpy()
        idx = R.probs_ref(logits, env["ign"]).argmax(2)
        flips += int((idx != ref_probs.argmax(2)).sum())
        steps += idx.size
        cert = margin >= CERTIFY_MARGIN
        n_cert += int(cert.sum())
        for a, b, c in zip(strs, ref_strs, cert):
            n_total += 1
            n_same += a == b
            n_bad_cert += (a != b) and bool(c)
    print(f"f16x3 vs oracle: {n_same}/{n_total} identical strings, {flips}/{steps} argmax decisions differ, "
          f"{n_cert} certified lines, {n_bad_cert} mismatches on certified lines")
    assert n_bad_cert == 0
    assert n_cert >= 0.90 * n_total
    assert n_same >= n_total - 1


def test_f16x3_error_against_float64_is_at_the_reference_fp32_level(env, synth_sd):
    """The truncation compensation of igemm_x3.cuh / lstm_tc.cuh, pinned: on 16 lines of 32x256 the oracle's graph is
    evaluated in float64 (truth) and in fp32 (what the reference computes); the f16x3 logits must be no further from
    the truth than 1.25 x the reference's own fp32 run, and the visual features must not be scaled - the least-squares
    factor of features = (1 + shrink) * truth stays within +-3e-7 (measured on B200: 0.75 x, +0.3e-7; without the
    compensation: 1.8 x, -8.9e-7, tools/error_budget.py)."""
    import torch.nn.functional as F
    from ocrfa_b200.synth import synth_lines
    from oracle import recognizer_ref as R

    torch.set_num_threads(os.cpu_count() or 1)
    sd32 = {k: torch.from_numpy(np.asarray(v)) if not torch.is_tensor(v) else v for k, v in synth_sd.items()}
    sd64 = {k: v.double() for k, v in sd32.items()}
    imgs = synth_lines(16, 32, 256, seed=777001)
    x = R.align_collate_ref(list(imgs), 32, 256)
    with torch.no_grad():
        f64 = R.resnet_ref(sd64, torch.from_numpy(x).double()).permute(0, 3, 1, 2).mean(dim=3)
        s64 = R.bilstm_ref(sd64, R.bilstm_ref(sd64, f64, "SequenceModeling.0."), "SequenceModeling.1.")
        lg64 = F.linear(s64, sd64["Prediction.weight"], sd64["Prediction.bias"]).numpy()
    ft64 = f64.numpy()
    lg32 = np.asarray(R.model_forward_ref(sd32, x), np.float64)
    m = _model(env, "f16x3")
    xd = torch.from_numpy(x).to(env["dev"])
    ft = m.features(xd).cpu().numpy().astype(np.float64)
    lg = m(xd, None).cpu().numpy().astype(np.float64)
    rms = lambda d: float(np.sqrt((d ** 2).mean()))
    e_ref, e_x3 = rms(lg32 - lg64), rms(lg - lg64)
    shrink = float(((ft - ft64) * ft64).sum() / (ft64 ** 2).sum())
    print(f"logit error vs float64 (of rms): reference fp32 {e_ref / rms(lg64):.3g}, f16x3 {e_x3 / rms(lg64):.3g}; "
          f"feature shrink {shrink:+.3g}")
    assert e_x3 <= 1.25 * e_ref
    assert abs(shrink) <= 3e-7


def test_single_pass_precisions_have_a_floor_at_config2_size(env, synth_sd):
    """bf16 / f16 (single-pass tensor-core modes) against the oracle at the config-2 shape: their agreement is bounded
    from below, not just printed. Measured on 10 240 lines (profiles/r01_parity_10k_32x512_all_modes.json): per-timestep
    argmax agreement 96.35 % (bf16) / 99.34 % (f16), identical strings 19.7 % / 71.5 %."""
    from ocrfa_b200.synth import synth_lines
    from oracle import recognizer_ref as R

    ob = env["ob"]
    imgs = synth_lines(128, 32, 512, seed=900000)
    x = R.align_collate_ref(list(imgs), 32, 512)
    ref_logits = R.model_forward_ref(synth_sd, x).numpy()
    ref_probs = R.probs_ref(ref_logits, env["ign"])
    ref_strs = R.greedy_decode_ref(ref_probs, env["conv"].character)
    batch = ob.AlignCollate(32, 512)(list(imgs))
    floors = {"bf16": (0.94, 0.08, 2e-2), "f16": (0.985, 0.55, 4e-3)}
    for prec, (step_floor, str_floor, range_tol) in floors.items():
        m = _model(env, prec)
        strs, _ = ob.recognize_host(m, env["conv"], batch.images_u8, batch.widths, env["mask"], "greedy", 5, env["dev"])
        logits = m.forward_u8(batch.images_u8.to(env["dev"])).cpu().numpy()
        agree = float((R.probs_ref(logits, env["ign"]).argmax(2) == ref_probs.argmax(2)).mean())
        same = sum(a == b for a, b in zip(strs, ref_strs)) / len(ref_strs)
        d = np.abs(logits - ref_logits)
        frac = float((d <= 2e-2 + 2e-2 * np.abs(ref_logits)).mean())
        print(f"{prec}: per-timestep agreement {agree:.4f}, identical strings {same:.3f}, max|d|/max|ref| "
              f"{d.max() / np.abs(ref_logits).max():.3g}, elementwise allclose(2e-2) fraction {frac:.5f}")
        assert agree >= step_floor and same >= str_floor
        assert d.max() <= range_tol * np.abs(ref_logits).max()
        if prec == "f16":
            assert frac >= 0.9999  # north_star's 16-bit tolerance, elementwise: |d| <= 2e-2 + 2e-2 |ref|


def test_margin_kernel_matches_numpy(env):
    from ocrfa_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(5)
    B, T, Cn = 7, 33, 185
    logits = rng.standard_normal((B, T, Cn)).astype(np.float32) * 4
    logits[1, 5, 10] = logits[1, 5, 20] = logits[1, 5].max() + 1.0   # exact tie of two unmasked classes -> margin 0
    logits[2, :, 3] = 100.0                                           # a masked class dominates: must be ignored
    logits[3, 7, 50] = logits[3, 7].max() + 1e-4                      # a tiny gap
    mask = np.zeros(Cn, np.uint8)
    mask[[3, 40, 41, 90]] = 1
    keep = mask == 0
    lk = np.sort(logits[:, :, keep].astype(np.float64), axis=2)
    exp = ((lk[:, :, -1] - lk[:, :, -2]) / np.sqrt((lk ** 2).mean(axis=2))).min(axis=1)
    d_l = torch.from_numpy(logits).cuda()
    d_m = torch.from_numpy(mask).cuda()
    out = torch.empty(B, dtype=torch.float32, device="cuda")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.ocrfa_ctc_margin(C.c_void_p(d_l.data_ptr()), B, T, Cn, C.c_void_p(d_m.data_ptr()), C.c_void_p(out.data_ptr()), stream))
    got = out.cpu().numpy()
    assert got[1] == 0.0
    np.testing.assert_allclose(got, exp, rtol=2e-5, atol=1e-7)
    _lib.check(lib.ocrfa_ctc_margin(C.c_void_p(d_l.data_ptr()), B, T, Cn, None, C.c_void_p(out.data_ptr()), stream))
    lk = np.sort(logits.astype(np.float64), axis=2)
    exp = ((lk[:, :, -1] - lk[:, :, -2]) / np.sqrt((lk ** 2).mean(axis=2))).min(axis=1)
    np.testing.assert_allclose(out.cpu().numpy(), exp, rtol=2e-5, atol=1e-7)


def test_async_host_entry_point_matches_the_synchronous_one(env):
    """ocrfa_recognize_host_ex with OCRFA_FLAG_ASYNC on two streams / two workspaces (the end-to-end bench path) returns
    what the synchronous ocrfa_recognize_host returns, for every batch, whatever the interleaving."""
    from ocrfa_b200 import _lib
    from ocrfa_b200.synth import synth_lines

    ob, lib, dev = env["ob"], _lib.load(), env["dev"]
    m = _model(env, "f16x3")
    handle = m._ensure_handle(dev)
    B, H, W = 24, 32, 256
    T = lib.ocrfa_out_timesteps(H, W)
    batches = [torch.from_numpy(synth_lines(B, H, W, seed=50 + i)).pin_memory() for i in range(5)]
    want = [ob.recognize_host(m, env["conv"], b, None, env["mask"], "greedy", 5, dev, return_margin=True) for b in batches]
    wsb = lib.ocrfa_recognize_workspace_bytes(handle, B, H, W, 1)
    slots = []
    for _ in range(2):
        buf = torch.empty(int(wsb) + 1024, dtype=torch.uint8, device=dev)
        slots.append(dict(ws=buf[(-buf.data_ptr()) % 1024:][:wsb], st=torch.cuda.Stream(dev),
                          lab=torch.empty((B, T), dtype=torch.int32).pin_memory(), ln=torch.empty(B, dtype=torch.int32).pin_memory(),
                          cf=torch.empty(B, dtype=torch.float64).pin_memory(), mg=torch.empty(B, dtype=torch.float32).pin_memory()))
    got = [None] * len(batches)
    pending = [None, None]
    for rep in range(2):  # the second round replays the captured CUDA graphs
        for i, b in enumerate(batches):
            s = slots[i & 1]
            s["st"].synchronize()
            if pending[i & 1] is not None:
                j = pending[i & 1]
                got[j] = (env["conv"].labels_to_text(s["lab"].numpy(), s["ln"].numpy()), s["cf"].numpy().copy(), s["mg"].numpy().copy())
            _lib.check(lib.ocrfa_recognize_host_ex(
                handle, C.c_void_p(b.data_ptr()), None, B, H, W, env["mask"].ctypes.data_as(C.c_void_p), 0, 1,
                C.c_void_p(s["lab"].data_ptr()), C.c_void_p(s["ln"].data_ptr()), C.c_void_p(s["cf"].data_ptr()),
                C.c_void_p(s["mg"].data_ptr()), _lib.FLAG_ASYNC, C.c_void_p(s["ws"].data_ptr()), wsb, C.c_void_p(s["st"].cuda_stream)))
            pending[i & 1] = i
        for k in range(2):
            slots[k]["st"].synchronize()
            j = pending[k]
            s = slots[k]
            got[j] = (env["conv"].labels_to_text(s["lab"].numpy(), s["ln"].numpy()), s["cf"].numpy().copy(), s["mg"].numpy().copy())
            pending[k] = None
        for (gs, gc, gm), (ws_, wc, wm) in zip(got, want):
            assert gs == ws_ and np.array_equal(gc, wc) and np.array_equal(gm, wm)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


_RANK_SCRIPT = r"""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
import ocrfa_b200 as ob
from ocrfa_b200.distributed import recognize_sharded, labels_to_strings
from ocrfa_b200.recognition import _ignore_mask
from ocrfa_b200.synth import synth_lines_global, synth_state_dict
from ocrfa_b200 import _lib
import ctypes as C

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)   # results travel over gloo: both ranks may share one GPU
dev = torch.device("cuda", rank % torch.cuda.device_count())
torch.cuda.set_device(dev)
chars = ob.default_character()
conv = ob.CTCLabelConverter(chars)
m = ob.Model(1, 512, 512, len(conv.character), precision="f16x3")
m.load_state_dict({{k: torch.from_numpy(v) for k, v in synth_state_dict(0).items()}})
ign = [chars.index(ch) + 1 for ch in ob.default_ignore_char()]
mask = _ignore_mask(len(conv.character), ign)
N, H, W = {n_lines}, 32, 256
lib = _lib.load()
T = lib.ocrfa_out_timesteps(H, W)

def rec(lo, hi):
    imgs = torch.from_numpy(synth_lines_global(lo, hi, H, W, 777))
    strs, conf = ob.recognize_host(m, conv, imgs, None, mask, "greedy", 5, dev)
    lab = torch.zeros((hi - lo, T), dtype=torch.int32)
    ln = torch.zeros(hi - lo, dtype=torch.int32)
    index = {{c: i for i, c in reversed(list(enumerate(conv.character)))}}
    for i, s in enumerate(strs):
        ln[i] = len(s)
        for j, ch in enumerate(s):
            lab[i, j] = index[ch]
    return lab, ln, torch.from_numpy(conf)

labels, lengths, conf = recognize_sharded(N, rec, T, batch_size=48)
if rank == 0:
    json.dump({{"strings": labels_to_strings(labels, lengths, conv.character), "conf": conf.tolist(),
               "launches": m.launch_count()}}, open({out!r}, "w"))
dist.barrier()
dist.destroy_process_group()
"""


def test_two_ranks_equal_one_rank_on_hardware(env, tmp_path):
    """SURVEY section 4 item 4 / the reference's DataParallel gather (utils.py:692): the same 150 lines decoded by two
    processes (contiguous shards, all_gather of the label tensors afterwards) and by one process give the same strings
    and confidences, in global line order. The collective runs over gloo so that the test also works when the box
    exposes a single GPU to both ranks; bench.py --gpus N repeats the check over NCCL at N = 2, 4, 8
    (multi_gpu_equivalence in its JSON line)."""
    import json
    import sys

    results = {}
    for world in (1, 2):
        out = str(tmp_path / f"w{world}.json")
        script = tmp_path / f"rank_w{world}.py"
        script.write_text(_RANK_SCRIPT.format(root=ROOT, n_lines=150, out=out))
        port = _free_port()
        procs = []
        for r in range(world):
            e = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
            procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        for p in procs:
            o, _ = p.communicate(timeout=600)
            assert p.returncode == 0, o[-3000:]
        results[world] = json.load(open(out))
    assert results[1]["launches"] > 0 and results[2]["launches"] > 0
    assert results[2]["strings"] == results[1]["strings"] and len(results[1]["strings"]) == 150
    assert results[2]["conf"] == results[1]["conf"]
    assert len(set(results[1]["strings"])) > 100  # the lines really differ from one another


def _run_harness(args, timeout, env=None):
    exe = os.path.join(ROOT, "build", "test_igemm")
    assert os.path.exists(exe), "build/test_igemm missing: run __graft_entry__.build()"
    out = subprocess.run([exe] + args, capture_output=True, text=True, timeout=timeout,
                         env=dict(os.environ, **env) if env else None)
    print(out.stdout[-4000:])
    assert out.returncode == 0 and "ALL OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
    return out.stdout


def test_lstm_publish_stress():
    """The persistent LSTM's only cross-CTA dependency (lstm_tc.cuh: TMA store of the h slice, completion wait,
    red.release.gpu / acquire poll + fence.proxy.async) under stress: 100 launches per geometry (c1 / c2 / c3: 1, 256, 64
    lines) and mode, each against a NaN-poisoned exchange buffer, every one bit-identical to the first over all batch
    rows, and the first within tolerance of a float64 host recurrence. Round 1's relaxed publish fails this run."""
    out = _run_harness(["--lstm-stress", "100"], 900)
    assert out.count("launches differ from the first (0 values) ok") == 6


@pytest.mark.parametrize("chain", [1, 2])
def test_f16x3_kernel_harness(chain):
    """Every conv / GEMM geometry of the path through igemm_x3_kernel (per-tap and halo-staged forms, fused 1x1 source,
    residual, fused pool) against a float64 CUDA reference of the same operand values (tools/test_igemm.cu --split), with
    one K chunk per hi.hi accumulation (the linear GEMMs and the detector) and with two (the recognizer's convolutions:
    alternating buffers, and on 256-column tiles the staggered column halves)."""
    _run_harness(["--split"], 600, env={"OCRFA_X3_CHAIN": str(chain)})
