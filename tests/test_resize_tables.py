"""CPU-only: the Pillow BICUBIC restatement in the oracle is bit-exact with PIL itself, and the product's
vectorised coefficient tables (ocrfa_b200/resize.py) equal the oracle's scalar ones."""
import numpy as np
import pytest
from PIL import Image

from ocrfa_b200.resize import bicubic_coeffs, resized_width
from oracle import recognizer_ref as R

SIZES = [(40, 333, 32, 267), (64, 512, 32, 256), (32, 100, 32, 100), (17, 90, 32, 170), (100, 37, 32, 12),
         (64, 2048, 64, 2000), (33, 77, 64, 150), (5, 9, 32, 58), (128, 64, 32, 16), (31, 500, 32, 517), (1, 7, 32, 224)]


@pytest.mark.parametrize("h,w,oh,ow", SIZES)
def test_oracle_resize_equals_pillow(h, w, oh, ow):
    rng = np.random.default_rng(h * 1000 + w)
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img, "L").resize((ow, oh), Image.BICUBIC))
    assert np.array_equal(R.pillow_bicubic_resize_ref(img, ow, oh), ref)


def test_product_tables_equal_oracle_tables():
    rng = np.random.default_rng(7)
    pairs = [(100, 32), (32, 100), (64, 32), (333, 267), (5, 32), (2048, 2000), (1, 32), (32, 1), (500, 517)]
    pairs += [(int(a), int(b)) for a, b in rng.integers(1, 900, (80, 2))]
    for a, b in pairs:
        b1, k1, s1 = bicubic_coeffs(a, b)
        b2, k2, s2 = R.pillow_coeffs_ref(a, b)
        assert s1 == s2 and np.array_equal(b1, b2) and np.array_equal(k1, k2), (a, b)


def test_resized_width_rule():
    assert resized_width(100, 32, 32, 320) == 100
    assert resized_width(333, 40, 32, 320) == R.resized_width_ref(333, 40, 32, 320) == 267
    assert resized_width(4000, 32, 32, 320) == 320
