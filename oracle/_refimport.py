"""Imports the UNMODIFIED reference package from /root/reference (TEST INFRASTRUCTURE, dev container only).

Four third-party modules the reference imports at module scope are absent here (skimage, bidi, shapely,
pyclipper). None is touched by the recognition path, so empty stand-ins are registered
before `src.ocrfa` is imported. /root/reference does not exist on the GPU box: only oracle/gen_golden.py
(which writes the committed fixtures under tests/golden/) may call this.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("OCRFA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "ocrfa"))


def import_reference():
    """Returns (utils_module, modules_module) of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in ("skimage", "skimage.io", "bidi", "bidi.algorithm", "shapely", "shapely.geometry", "pyclipper"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["skimage"].io = sys.modules["skimage.io"]
    sys.modules["bidi"].algorithm = sys.modules["bidi.algorithm"]
    sys.modules["bidi.algorithm"].get_display = lambda s: s
    sys.modules["shapely"].geometry = sys.modules["shapely.geometry"]
    sys.modules["shapely.geometry"].Polygon = object
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    utils = importlib.import_module("src.ocrfa.utils")
    modules = importlib.import_module("src.ocrfa.modules")
    return utils, modules
