"""One pass of the detection glue kernels for an `ncu` capture (profiling infrastructure, not the product):
detector input (shrink mode) + grey conversion on a 2200 x 3400 page, component labeling on a 368 x 640 map pair."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ocrfa_b200.detection import detector_input, get_det_boxes  # noqa: E402
from ocrfa_b200.reader import bgr2gray_device  # noqa: E402
from ocrfa_b200.synth import synth_score_maps  # noqa: E402

rng = np.random.default_rng(0)
page = torch.from_numpy(rng.integers(0, 256, (2200, 3400, 3), dtype=np.uint8)).cuda()
text, link = synth_score_maps(368, 640, 31, 160)
y = torch.from_numpy(np.stack([text, link], axis=-1)).cuda()
for _ in range(2):
    detector_input(page, 2560, 1.0)
    bgr2gray_device(page)
    boxes = get_det_boxes(y[:, :, 0], y[:, :, 1], 0.7, 0.4, 0.4)[0]
torch.cuda.synchronize()
print("boxes", len(boxes))
