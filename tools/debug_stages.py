"""GPU bring-up aid: stage-wise comparison of the bf16 / fp32 CUDA paths with the oracle (fp32 and bf16-emulated)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocrfa_b200 as ob
from oracle import recognizer_ref as R
from oracle.synth import synth_state_dict

def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b)), float(np.abs(a - b).max())

def sc(a, b):
    return float((np.abs(a - b) / (1 + np.abs(b))).max())

sd = synth_state_dict(0)
name = sys.argv[1] if len(sys.argv) > 1 else "c1_32x128"
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", name + ".npz"))
x = g["x"]; H, W = x.shape[2], x.shape[3]
models = {}
for prec in ("fp32", "bf16"):
    m = ob.Model(1, 512, 512, 185, precision=prec)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    models[prec] = m
ref_logits, ref_feat = R.model_forward_ref(sd, x, return_features=True)
ref_logits, ref_feat = ref_logits.numpy(), ref_feat.numpy()
emu_feat = R.features_bf16emu_ref(sd, x).numpy()
emu_logits = R.sequence_bf16emu_ref(sd, emu_feat).numpy()
xc = torch.from_numpy(x).cuda()
f32 = models["fp32"].features(xc).cpu().numpy()
f16 = models["bf16"].features(xc).cpu().numpy()
print("features: gpu fp32 vs oracle fp32   (relL2, maxabs):", rel(f32, ref_feat))
print("features: emu bf16 vs oracle fp32                  :", rel(emu_feat, ref_feat))
print("features: gpu bf16 vs oracle fp32                  :", rel(f16, ref_feat))
print("features: gpu bf16 vs emu bf16                     :", rel(f16, emu_feat))
# head on identical inputs
h16 = models["bf16"].sequence(torch.from_numpy(emu_feat).cuda(), H, W).cpu().numpy()
h32 = models["fp32"].sequence(torch.from_numpy(emu_feat).cuda(), H, W).cpu().numpy()
ref_head_on_emu = R.sequence_bf16emu_ref(sd, emu_feat).numpy()
import torch.nn.functional as F
with torch.no_grad():
    s = R.bilstm_ref(sd, torch.from_numpy(emu_feat), "SequenceModeling.0.")
    s = R.bilstm_ref(sd, s, "SequenceModeling.1.")
    fp32_head_on_emu = F.linear(s, R._t(sd, "Prediction.weight"), R._t(sd, "Prediction.bias")).numpy()
print("head(emu feat): gpu fp32 vs oracle fp32 head  scaled:", sc(h32, fp32_head_on_emu))
print("head(emu feat): emu bf16 head vs fp32 head    scaled:", sc(ref_head_on_emu, fp32_head_on_emu))
print("head(emu feat): gpu bf16 vs fp32 head         scaled:", sc(h16, fp32_head_on_emu))
print("head(emu feat): gpu bf16 vs emu bf16 head     scaled:", sc(h16, ref_head_on_emu), "maxabs", float(np.abs(h16 - ref_head_on_emu).max()))
full16 = models["bf16"](xc, None).cpu().numpy()
print("full: gpu bf16 vs emu bf16 scaled:", sc(full16, emu_logits), " vs fp32 ref scaled:", sc(full16, ref_logits),
      " range-norm:", float(np.abs(full16 - ref_logits).max() / np.abs(ref_logits).max()))
print("full: argmax agreement gpu bf16 vs ref:", float((full16.argmax(2) == ref_logits.argmax(2)).mean()),
      " emu vs ref:", float((emu_logits.argmax(2) == ref_logits.argmax(2)).mean()))
