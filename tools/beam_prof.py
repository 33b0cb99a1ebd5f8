"""One-off section profile of the beam kernel (build/libocrfa_prof.so = engine.cu compiled with -DOCRFA_BEAM_PROF):
runs config 4 once through the profiling build; block 0 prints cycles per step per section. Measurement tooling."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ocrfa_b200._lib as L
L.LIB_PATH = os.path.join(ROOT, "build", "libocrfa_prof.so")
import ocrfa_b200 as ob
from ocrfa_b200.synth import synth_lines, synth_state_dict
from ocrfa_b200.recognition import _ignore_mask
chars = ob.default_character(); conv = ob.CTCLabelConverter(chars)
m = ob.Model(1, 512, 512, len(conv.character), precision="bf16")
m.load_state_dict({k: torch.from_numpy(v) for k, v in synth_state_dict(0).items()})
imgs = synth_lines(128, 32, 512, seed=1234)
batch = ob.AlignCollate(32, 512)(list(imgs))
mask = _ignore_mask(len(conv.character), [chars.index(ch) + 1 for ch in ob.default_ignore_char()])
os.environ["OCRFA_GRAPHS"] = "0"
s, c = ob.recognize_host(m, conv, batch.images_u8, batch.widths, mask, "beamsearch", 8, torch.device("cuda", 0))
torch.cuda.synchronize()
print(s[0])
