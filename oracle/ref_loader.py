"""Load the UNMODIFIED reference scripts from /root/reference for golden generation.

TEST INFRASTRUCTURE ONLY.  This module works only where /root/reference exists
(the build container); nothing under tests -m gpu, smoke() or bench.py imports it.

The reference scripts run a whole video at import time (Project.py:354-361,
Project_harder_challenge.py:430-437) and import matplotlib / moviepy / skimage,
none of which are installed.  We therefore install inert stub modules, exec the
source verbatim under a private module name, and chdir into the reference
directory only while the module body runs (the pickle path is relative,
Project.py:346).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REF_DIR = os.environ.get("LANE_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "Project.py"))


def read_rgb(path: str) -> np.ndarray:
    """JPEG -> uint8 RGB, identical to matplotlib.image.imread for JPEG (PIL decode)."""
    from PIL import Image

    with Image.open(path) as im:
        return np.asarray(im.convert("RGB")).copy()


class _Clip:
    def __init__(self, *a, **k):
        pass

    def fl_image(self, fn):
        return self

    def subclip(self, *a, **k):
        return self

    def write_videofile(self, *a, **k):
        return None


def _install_stubs() -> None:
    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    if "matplotlib" not in sys.modules:
        mpl = mod("matplotlib")
        mpl.pyplot = mod("matplotlib.pyplot")
        mpl.image = mod("matplotlib.image", imread=read_rgb)
    if "moviepy" not in sys.modules:
        mp = mod("moviepy")
        mp.editor = mod("moviepy.editor", VideoFileClip=_Clip)
    if "skimage" not in sys.modules:
        sk = mod("skimage")
        sk.exposure = mod("skimage.exposure")
    if not hasattr(np, "int"):
        np.int = int  # removed in numpy>=1.24; used at Project_harder_challenge.py:108,133, helper.py:238..289


def load(script: str, alias: str):
    """exec `script` (e.g. 'Project.py') verbatim; returns the module object."""
    if not available():
        raise RuntimeError(f"reference not present at {REF_DIR}")
    _install_stubs()
    path = os.path.join(REF_DIR, script)
    spec = importlib.util.spec_from_file_location(alias, path)
    m = importlib.util.module_from_spec(spec)
    cwd = os.getcwd()
    os.chdir(REF_DIR)
    try:
        spec.loader.exec_module(m)
    finally:
        os.chdir(cwd)
    return m


def fresh_lines(m) -> None:
    """Reset the module-global tracking state (Project.py:351-352)."""
    m.left = m.Line()
    m.right = m.Line()


def load_helper():
    import pickle

    h = load("helper.py", "_ref_helper")
    with open(os.path.join(REF_DIR, "wide_dist_pickle.p"), "rb") as f:
        cal = pickle.load(f)
    h.mtx, h.dist = cal["mtx"], cal["dist"]
    h.src = np.float32([[490, 482], [810, 482], [1250, 720], [0, 720]])  # helper.py:880-883
    h.dst = np.float32([[0, 0], [1280, 0], [1250, 720], [40, 720]])
    return h
