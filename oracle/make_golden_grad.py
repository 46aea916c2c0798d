"""Generate tests/golden/reference_golden_grad.json by EXECUTING THE UNMODIFIED reference helper.py
(build container only):  python -m oracle.make_golden_grad

helper.img2binary / sobelx_filter / color_filter / combine_bin (helper.py:50-122, 493-511) on undistorted stills.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader as R  # noqa: E402

STILLS = os.path.join(ROOT, "tests", "golden", "stills")
NAMES = ["test1", "test4", "straight_lines2", "test_ch3", "test_ch11", "test_ch12"]


def rec(a):
    return {"sum": int(a.sum()), "sha1": hashlib.sha1(np.packbits(a.astype(bool), axis=1, bitorder="little").tobytes()).hexdigest()}


def main():
    import cv2
    h = R.load_helper()
    gold = {"versions": {"cv2": cv2.__version__, "numpy": np.__version__}, "stills": {}}
    for n in NAMES:
        und = cv2.undistort(R.read_rgb(os.path.join(STILLS, n + ".jpg")), h.mtx, h.dist, None, h.mtx)
        gold["stills"][n] = {
            "img2binary": rec(h.img2binary(und)),                                   # defaults (100,255), (20,100)
            "img2binary_90_30_120": rec(h.img2binary(und, (90, 255), (30, 120))),
            "sobelx_filter": rec(h.sobelx_filter(und)),
            "color_filter": rec(h.color_filter(und)),
            "combine_bin": rec(h.combine_bin(und)),
        }
        print(n, gold["stills"][n]["img2binary"]["sum"], flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "reference_golden_grad.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print("written")


if __name__ == "__main__":
    main()
