import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
STILLS = os.path.join(GOLDEN, "stills")
PKG_NAME = "advanced-lane-detection_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """tests marked `gpu` are skipped -- not failed -- on a machine without a CUDA device"""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (run on the B200 box: python -m pytest tests -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_pkg():
    """The package directory has a hyphen in its name; import it by string."""
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "reference_golden.json")) as f:
        return json.load(f)


def read_still(name):
    from PIL import Image
    with Image.open(os.path.join(STILLS, name + ".jpg")) as im:
        return np.asarray(im.convert("RGB")).copy()


@pytest.fixture(scope="session")
def stills():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = read_still(name)
        return cache[name]
    return get


def unhex(v):
    return float.fromhex(v)
