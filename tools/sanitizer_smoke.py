import numpy as np, torch, sys
sys.path.insert(0, '/root/repo')
import ocrfa_b200 as ob
from ocrfa_b200.synth import synth_lines, synth_state_dict
from ocrfa_b200.recognition import _ignore_mask
chars = ob.default_character(); conv = ob.CTCLabelConverter(chars)
for prec in ("bf16", "f16"):
    m = ob.Model(1, 512, 512, len(conv.character), precision=prec)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth_state_dict(0).items()})
    for (n, H, W) in ((5, 32, 128), (3, 64, 256)):
        imgs = synth_lines(n, H, W, seed=5)
        batch = ob.AlignCollate(H, W)(list(imgs))
        ign = [chars.index(ch) + 1 for ch in ob.default_ignore_char()]
        mask = _ignore_mask(len(conv.character), ign)
        for dec, bw in (("greedy", 5), ("beamsearch", 8)):
            s, c = ob.recognize_host(m, conv, batch.images_u8, batch.widths, mask, dec, bw, torch.device("cuda", 0))
            print(prec, H, W, dec, s[:2])
print("sanity done")
