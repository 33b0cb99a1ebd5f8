"""Imports the UNMODIFIED reference package from /root/reference (TEST INFRASTRUCTURE, dev container only).

Four third-party modules the reference imports at module scope are absent here (skimage, bidi, shapely,
pyclipper). None is touched by the recognition path, so empty stand-ins are registered
before `src.ocrfa` is imported. Callers: the oracle/gen_golden*.py scripts (which write the committed fixtures under
tests/golden/) and bench.py's CPU legs (`--impl reference`, `cpu_baseline`), which time the reference's own
recognizer_predict when a copy of the tree is present and fall back to the oracle port otherwise.
"""
from __future__ import annotations

import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _resolve_root() -> str:
    """/root/reference in the dev container; on the GPU box the copy `__graft_entry__.build()` leaves under
    oracle/_ref/ (git-ignored build output: it travels with the snapshot, it is never committed)."""
    env = os.environ.get("OCRFA_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isdir(os.path.join(cand, "src", "ocrfa")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _resolve_root()


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
