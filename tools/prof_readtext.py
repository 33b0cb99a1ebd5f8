import cProfile, pstats, os, sys, tempfile, io
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import ocrfa_b200 as ob
from ocrfa_b200.detection import detector_input
from ocrfa_b200.synth import synth_state_dict
rng = np.random.default_rng(0)
with tempfile.TemporaryDirectory() as d:
    torch.save({k: torch.from_numpy(v) for k, v in synth_state_dict(0).items()}, os.path.join(d, "arabic.pth"))
    torch.save({k: torch.from_numpy(v) for k, v in ob.synth_craft_state_dict(0).items()}, os.path.join(d, "craft_mlt_25k.pth"))
    reader = ob.OCR(model_storage_directory=d, verify_md5=False)
rgb = torch.from_numpy(rng.integers(0, 256, (736, 1280, 3), dtype=np.uint8)).cuda()
x, _ = detector_input(rgb, 2560, 1.0)
yy, _ = reader.detector(x)
tmap, lmap = yy[0, :, :, 0], yy[0, :, :, 1]
low, tt = float(torch.quantile(tmap.flatten()[:1 << 20], 0.80)), float(torch.quantile(tmap.flatten()[:1 << 20], 0.93))
lt = float(torch.quantile(lmap.flatten()[:1 << 20], 0.97))
kw = dict(batch_size=64, text_threshold=tt, low_text=low, link_threshold=lt, min_size=10)
for _ in range(2): res = reader.readtext(rgb, **kw)
pr = cProfile.Profile(); pr.enable()
for _ in range(5): res = reader.readtext(rgb, **kw)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
print(len(res))
