"""Decoded-string parity on >= 10k synthetic lines (north_star's correctness target), run on the GPU box:

    python tools/parity_10k.py [--lines 10240] [--height 32] [--width 128] [--out gpurun_out/parity_10k.json]

For every line: the oracle (CPU restatement of the reference, fp32) decodes greedy strings; the CUDA path decodes
the same uint8 lines in fp32 parity mode, in split-precision (f16x3) tensor-core mode and in bf16 tensor-core mode
through ocrfa_recognize_host. Reports how
many strings are identical, and the logit error statistics on a subsample. Test infrastructure (imports oracle/).

Two more columns put the numbers in context:
  * "reference@batch32": the oracle (the reference's torch CPU fp32 ops) run again on the SAME lines in batches of 32
    instead of 256 - oneDNN / MKL pick their blocking from the problem size, so the reference's logits move by ~5e-7 of
    their rms and a few near-tie timesteps flip: the reference is only reproducible against itself up to that noise.
  * decision margins (ocrfa_ctc_margin): a line is CERTIFIED when its smallest top-2 logit gap exceeds
    recognition.CERTIFY_MARGIN of the row's logit rms; certified lines must decode identically in every fp32-grade
    evaluation, and every mismatch must sit on an uncertified line.
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

import ocrfa_b200 as ob  # noqa: E402
from ocrfa_b200.recognition import CERTIFY_MARGIN, _ignore_mask  # noqa: E402
from ocrfa_b200.synth import synth_lines, synth_state_dict  # noqa: E402
from oracle import recognizer_ref as R  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lines", type=int, default=10240)
    ap.add_argument("--height", type=int, default=32)
    ap.add_argument("--width", type=int, default=128)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--precisions", default="fp32,f16x3,bf16")
    ap.add_argument("--ref-batch", type=int, default=32, help="second oracle pass in batches of this size (0 = skip)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_10k.json"))
    args = ap.parse_args()
    N, H, W = args.lines, args.height, args.width
    torch.set_num_threads(os.cpu_count() or 1)
    sd_np = synth_state_dict(0)
    sd = {k: torch.from_numpy(v) for k, v in sd_np.items()}
    chars = ob.default_character()
    conv = ob.CTCLabelConverter(chars)
    ign = R.ignore_idx_ref(chars, ob.default_ignore_char())
    mask = _ignore_mask(len(conv.character), ign)
    dev = torch.device("cuda", 0)
    models = {}
    precs = tuple(args.precisions.split(","))
    for prec in precs:
        m = ob.Model(1, 512, 512, len(conv.character), precision=prec)
        m.load_state_dict(sd)
        models[prec] = m

    cols = precs + ((f"reference@batch{args.ref_batch}",) if args.ref_batch else ())
    same = {k: 0 for k in cols}
    step_agree = {k: [0, 0] for k in cols}
    cert = {k: {"certified_lines": 0, "mismatches_on_certified_lines": 0, "mismatches_on_uncertified_lines": 0} for k in precs}
    ref_uncertified = 0
    conf_err = {k: 0.0 for k in precs}
    logit_err = {k: 0.0 for k in precs}
    mismatches = {k: [] for k in cols}
    t_cpu = t_gpu = 0.0
    for b0 in range(0, N, args.batch):
        n = min(args.batch, N - b0)
        imgs = synth_lines(n, H, W, seed=900000 + b0)  # disjoint from the timing set (seed 1234 + rank)
        t0 = time.perf_counter()
        x = R.align_collate_ref(list(imgs), H, W)
        logits = R.model_forward_ref(sd, x)
        probs = R.probs_ref(logits, ign)
        ref_strs = R.greedy_decode_ref(probs, conv.character)
        ref_conf = np.array(R.confidence_ref(probs), dtype=np.float64)
        ref_idx = probs.argmax(2)
        t_cpu += time.perf_counter() - t0
        batch = ob.AlignCollate(H, W)(list(imgs))
        ref_logits = logits.numpy() if hasattr(logits, "numpy") else np.asarray(logits)
        # the reference's own decision margins (same definition as ocrfa_ctc_margin, numpy)
        keep = np.ones(ref_logits.shape[2], bool)
        keep[list(ign)] = False
        lk = np.sort(ref_logits[:, :, keep], axis=2)
        ref_margin = ((lk[:, :, -1] - lk[:, :, -2]) / np.sqrt((lk.astype(np.float64) ** 2).mean(axis=2))).min(axis=1)
        ref_uncertified += int((ref_margin < CERTIFY_MARGIN).sum())
        if args.ref_batch:
            col = cols[-1]
            strs2, idx2 = [], []
            for s0 in range(0, n, args.ref_batch):
                lg2 = R.model_forward_ref(sd, x[s0:s0 + args.ref_batch])
                pr2 = R.probs_ref(lg2, ign)
                strs2 += R.greedy_decode_ref(pr2, conv.character)
                idx2.append(pr2.argmax(2))
            idx2 = np.concatenate(idx2)
            for i, (a, b) in enumerate(zip(strs2, ref_strs)):
                if a == b:
                    same[col] += 1
                elif len(mismatches[col]) < 5:
                    mismatches[col].append({"line": b0 + i, "got": a, "ref": b, "ref_margin": float(ref_margin[i])})
            step_agree[col][0] += int((idx2 == ref_idx).sum())
            step_agree[col][1] += idx2.size
        for prec in precs:
            t0 = time.perf_counter()
            strs, conf, margin = ob.recognize_host(models[prec], conv, batch.images_u8, batch.widths, mask, "greedy", 5, dev,
                                                   return_margin=True)
            t_gpu += time.perf_counter() - t0
            certified = margin >= CERTIFY_MARGIN
            cert[prec]["certified_lines"] += int(certified.sum())
            for i, (a, b) in enumerate(zip(strs, ref_strs)):
                if a == b:
                    same[prec] += 1
                else:
                    cert[prec]["mismatches_on_certified_lines" if certified[i] else "mismatches_on_uncertified_lines"] += 1
                    if len(mismatches[prec]) < 5:
                        mismatches[prec].append({"line": b0 + i, "got": a, "ref": b, "margin": float(margin[i]),
                                                 "ref_margin": float(ref_margin[i])})
            ok = np.isfinite(ref_conf) & (ref_conf > 0)
            if ok.any():
                conf_err[prec] = max(conf_err[prec], float(np.max(np.abs(conf[ok] - ref_conf[ok]) / ref_conf[ok])))
            lg = models[prec].forward_u8(batch.images_u8.to(dev), None).cpu().numpy()
            logit_err[prec] = max(logit_err[prec], float((np.abs(lg - ref_logits) / (1.0 + np.abs(ref_logits))).max()))
            idx = R.probs_ref(lg, ign).argmax(2)
            step_agree[prec][0] += int((idx == ref_idx).sum())
            step_agree[prec][1] += idx.size
        print(f"{b0 + n}/{N}: identical " + " ".join(f"{k} {v}" for k, v in same.items()), flush=True)
    out = {
        "lines": N, "shape": [H, W], "decoder": "greedy", "weights": "synthetic seed 0 (ocrfa_b200/synth.py)",
        "identical_strings": same,
        "identical_fraction": {k: v / N for k, v in same.items()},
        "per_timestep_argmax_agreement": {k: v[0] / max(1, v[1]) for k, v in step_agree.items()},
        "max_rel_confidence_error": conf_err,
        "max_scaled_logit_error": logit_err,  # max |got - ref| / (1 + |ref|): compare with rtol = atol
        "first_mismatches": mismatches,
        "certify_margin": CERTIFY_MARGIN, "certificates": cert,
        "reference_lines_below_certify_margin": ref_uncertified,
        "cpu_oracle_lines_per_s": N / t_cpu, "cpu_threads": torch.get_num_threads(),
    }
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False, indent=1)
    print(json.dumps(out, ensure_ascii=False))


if __name__ == "__main__":
    main()
