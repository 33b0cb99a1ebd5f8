"""CPU restatement of the reference's CRAFT detector forward (TEST INFRASTRUCTURE - only tests/, smoke() and the
baseline legs may import this).

Follows `OCRFA.forward` (/root/reference/src/ocrfa/utils.py:75-98) over `Vgg16BN.forward`
(/root/reference/src/ocrfa/modules/blocks.py:69-82; slice boundaries :46-59) and `DoubleConv`
(utils.py:31-46), functional torch fp32 on a state dict with the reference's keys. Pinned by
tests/golden/craft_64x96.npz, which oracle/gen_golden_craft.py produced with the reference's own OCRFA module.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def _t(sd, k):
    v = sd[k]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))


def _conv(sd, x, key, pad=0, dil=1):
    return F.conv2d(x, _t(sd, key + ".weight"), _t(sd, key + ".bias"), padding=pad, dilation=dil)


def _bn(sd, x, key):
    return F.batch_norm(x, _t(sd, key + ".running_mean"), _t(sd, key + ".running_var"), _t(sd, key + ".weight"),
                        _t(sd, key + ".bias"), False, 0.0, BN_EPS)


def _cbr(sd, x, conv, bn, relu=True):
    y = _bn(sd, _conv(sd, x, conv, pad=1), bn)
    return F.relu(y) if relu else y


def vgg16bn_ref(sd: Dict[str, np.ndarray], x: torch.Tensor):
    """Vgg16BN.forward. The slices cut torchvision's vgg16_bn.features between a BatchNorm and its ReLU, and that ReLU
    is `nn.ReLU(inplace=True)`: the tensors returned as relu2_2 / relu3_2 / relu4_3 are the very objects the next
    slice's first module rectifies in place, so the U network receives them AFTER the ReLU. relu5_3 is followed by
    slice5's MaxPool (not in place) and stays pre-ReLU. Reproduced, not fixed."""
    p = "basenet."
    h = _cbr(sd, x, p + "slice1.0", p + "slice1.1")
    h = _cbr(sd, h, p + "slice1.3", p + "slice1.4")
    h = F.max_pool2d(h, 2, 2)
    h = _cbr(sd, h, p + "slice1.7", p + "slice1.8")
    relu2_2 = _cbr(sd, h, p + "slice1.10", p + "slice1.11")               # rectified in place by slice2[12]
    h = F.max_pool2d(relu2_2, 2, 2)
    h = _cbr(sd, h, p + "slice2.14", p + "slice2.15")
    relu3_2 = _cbr(sd, h, p + "slice2.17", p + "slice2.18")               # rectified in place by slice3[19]
    h = _cbr(sd, relu3_2, p + "slice3.20", p + "slice3.21")
    h = F.max_pool2d(h, 2, 2)
    h = _cbr(sd, h, p + "slice3.24", p + "slice3.25")
    relu4_3 = _cbr(sd, h, p + "slice3.27", p + "slice3.28")               # rectified in place by slice4[29]
    h = _cbr(sd, relu4_3, p + "slice4.30", p + "slice4.31")
    h = F.max_pool2d(h, 2, 2)
    h = _cbr(sd, h, p + "slice4.34", p + "slice4.35")
    relu5_3 = _cbr(sd, h, p + "slice4.37", p + "slice4.38", relu=False)   # slice4 ends on the BatchNorm; no ReLU follows
    h = F.max_pool2d(relu5_3, 3, 1, 1)
    h = _conv(sd, h, p + "slice5.1", pad=6, dil=6)
    fc7 = _conv(sd, h, p + "slice5.2")
    return fc7, relu5_3, relu4_3, relu3_2, relu2_2


def _double_conv(sd, x, pre):
    y = F.relu(_bn(sd, _conv(sd, x, pre + ".conv.0"), pre + ".conv.1"))
    return F.relu(_bn(sd, _conv(sd, y, pre + ".conv.3", pad=1), pre + ".conv.4"))


def craft_forward_ref(sd: Dict[str, np.ndarray], x):
    """x float32 [N,3,H,W] -> (y [N,H/2,W/2,2], feature [N,32,H/2,W/2])."""
    with torch.no_grad():
        x = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
        s = vgg16bn_ref(sd, x.float())
        y = _double_conv(sd, torch.cat([s[0], s[1]], dim=1), "upconv1")
        y = F.interpolate(y, size=s[2].shape[2:], mode="bilinear", align_corners=False)
        y = _double_conv(sd, torch.cat([y, s[2]], dim=1), "upconv2")
        y = F.interpolate(y, size=s[3].shape[2:], mode="bilinear", align_corners=False)
        y = _double_conv(sd, torch.cat([y, s[3]], dim=1), "upconv3")
        y = F.interpolate(y, size=s[4].shape[2:], mode="bilinear", align_corners=False)
        feature = _double_conv(sd, torch.cat([y, s[4]], dim=1), "upconv4")
        y = F.relu(_conv(sd, feature, "conv_cls.0", pad=1))
        y = F.relu(_conv(sd, y, "conv_cls.2", pad=1))
        y = F.relu(_conv(sd, y, "conv_cls.4", pad=1))
        y = F.relu(_conv(sd, y, "conv_cls.6"))
        y = _conv(sd, y, "conv_cls.8")
    return y.permute(0, 2, 3, 1).contiguous(), feature
