"""Generate tests/golden/reference_golden_tracked.json by EXECUTING THE UNMODIFIED reference helper.py
(build container only):  python -m oracle.make_golden_tracked

helper.skip_sliding_window / helper.sliding_window (helper.py:224-378) driven frame by frame over synthetic sequences of
bird's-eye binaries, with the driver loop of oracle.lane_port.helper_tracked_search (the loop helper.refine, helper.py:381-403,
is written for).  The binaries are produced by the oracle (binary_birdseye of the project variant); what is pinned here is
what the REFERENCE's searches select and fit on them, frame after frame.
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import lane_port as O  # noqa: E402
from oracle import ref_loader as R  # noqa: E402

STILLS = os.path.join(ROOT, "tests", "golden", "stills")
PROJECT = ["straight_lines1", "straight_lines2", "test1", "test2", "test3", "test4", "test5", "test6"]
# (name, stills, frames, seed, events): events = {frame: op}; "blank" empties the mask (both searches return None), "left"
# empties its left half (band AND window search lose the left lane), ("shift", dx) moves the mask sideways (the bands lose the
# lanes, the window search finds them again)
# init = the fits the walk starts from (None: cold start)
CASES = [
    ("plain_24", PROJECT, 24, 2, {}, None),
    ("events_40", PROJECT, 40, 7, {5: "blank", 6: "blank", 12: ("shift", 160), 13: ("shift", 160), 20: "left", 30: ("shift", -140)}, None),
    ("cold_start_blank_16", PROJECT, 16, 9, {0: "blank", 1: "blank", 9: "left"}, None),
    # start from fits in the middle of the road: the bands of frame 0 hold no lane pixels -> None -> window search
    ("lost_bands_12", PROJECT, 12, 4, {6: ("shift", 40)}, [[0.0, 0.0, 640.0], [0.0, 0.0, 700.0]]),
    # start from fits that are good: frame 0 is a band search already
    ("warm_start_12", PROJECT, 12, 5, {}, [[1.0e-4, -0.2, 330.0], [1.0e-4, -0.15, 1060.0]]),
]


def tracked_masks(stills, n, seed, events):
    """the bird's-eye binaries of a tracked-search case (shared with tests/test_tracked_search.py)"""
    orc = O.LaneOracle("project")
    masks = []
    for i, f in enumerate(O.synth_sequence(stills, n, seed=seed)):
        m = O.binary_birdseye(O.undistort(f, orc.mtx, orc.dist), orc.v)
        ev = events.get(i)
        if ev == "blank":
            m = np.zeros_like(m)
        elif ev == "left":
            m = m.copy()
            m[:, :640] = 0
        elif isinstance(ev, (tuple, list)):
            m = np.roll(m, int(ev[1]), axis=1)
        masks.append(np.ascontiguousarray(m))
    return masks


def hexes(v):
    return None if v is None else [float(x).hex() for x in v]


def main():
    import cv2
    h = R.load_helper()
    gold = {"versions": {"cv2": cv2.__version__, "numpy": np.__version__}, "cases": {}}
    for name, names, n, seed, events, init in CASES:
        stills = [R.read_rgb(os.path.join(STILLS, s + ".jpg")) for s in names]
        masks = tracked_masks(stills, n, seed, events)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            got = O.helper_tracked_search(masks, *(init or (None, None)), margin_search=h.skip_sliding_window, window_search=h.sliding_window)
        gold["cases"][name] = {"stills": names, "frames": n, "seed": seed, "events": {str(k): v for k, v in events.items()}, "init": init,
                               "frames_out": [{"left_fit": hexes(lf), "right_fit": hexes(rf), "n_left": nl, "n_right": nr, "source": src}
                                              for lf, rf, nl, nr, src in got]}
        print(name, [g[4] for g in got], flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "reference_golden_tracked.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print("written")


if __name__ == "__main__":
    main()
