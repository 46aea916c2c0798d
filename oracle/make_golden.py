"""Generate tests/golden/*.json by EXECUTING THE UNMODIFIED REFERENCE (build container only).

    python -m oracle.make_golden

Every number written here comes out of /root/reference/Project.py or
/root/reference/Project_harder_challenge.py run through oracle/ref_loader.py; nothing
comes from this repository's own restatement.  The JPEG stills under tests/golden/stills/
are byte copies of /root/reference/test_images/*.jpg (data fixtures, not source).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import warnings

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader as R  # noqa: E402
from oracle.lane_port import synth_sequence, integer_moments  # noqa: E402  (generator + exact sums only)

STILLS = os.path.join(ROOT, "tests", "golden", "stills")
PROJECT_STILLS = ["straight_lines1", "straight_lines2", "test1", "test2", "test3", "test4", "test5", "test6"]
CH_STILLS = ["test_ch%d" % i for i in range(15)]


def fhex(v):
    return float(v).hex()


def packed_sha1(mask):
    return hashlib.sha1(np.packbits(mask.astype(bool), axis=1, bitorder="little").tobytes()).hexdigest()


def still_record(mod, img, l_thresh, b_thresh):
    import cv2
    R.fresh_lines(mod)
    und = cv2.undistort(img, mod.mtx, mod.dist, None, mod.mtx)
    mask = mod.luv_lab_filter(und, l_thresh=l_thresh, b_thresh=b_thresh)
    rec = {"mask_popcount": int(mask.sum()), "mask_sha1": packed_sha1(mask),
           "undist_md5": hashlib.md5(und.tobytes()).hexdigest()}
    nzx, nzy = np.nonzero(np.transpose(mask))
    lanes = []
    for name in ("left", "right"):
        ret = getattr(mod, name).blind_search(nzx, nzy, mask)
        x, y = ret[0], ret[1]
        if x is None:
            lanes.append(None)
        else:
            lanes.append({"n": len(x), "moments": [str(m) for m in integer_moments(x, y)]})
    rec["lanes"] = lanes
    R.fresh_lines(mod)
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        try:
            out = mod.process_image(img)
        except TypeError as e:
            rec["status"] = "TypeError: %s" % e
            return rec
    rec["status"] = "ok"
    rec["rank_warnings"] = sum(1 for w in wlist if "RankWarning" in type(w.message).__name__ or "Polyfit" in str(w.message))
    rec["left_fit"] = [fhex(c) for c in mod.left.fit]
    rec["right_fit"] = [fhex(c) for c in mod.right.fit]
    rec["n_fit_points"] = [int(len(mod.left.fity)), int(len(mod.right.fity))]
    if mod.__name__.endswith("project"):
        off, curv = mod.car_pos(mod.left.fit, mod.right.fit)
    else:
        off = mod.car_pos(mod.left.fit, mod.right.fit)
        curv = np.mean([mod.curvature(mod.left.fit), mod.curvature(mod.right.fit)])
    rec["offset"], rec["mean_curv"] = fhex(off), fhex(curv)
    rec["out_md5"] = hashlib.md5(out.tobytes()).hexdigest()
    return rec


def sequence_records(mod, frames):
    R.fresh_lines(mod)
    recs = []
    for f in frames:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = mod.process_image(f)
        if mod.__name__.endswith("project"):
            off, curv = mod.car_pos(mod.left.fit, mod.right.fit)
        else:
            off = mod.car_pos(mod.left.fit, mod.right.fit)
            curv = np.mean([mod.curvature(mod.left.fit), mod.curvature(mod.right.fit)])
        recs.append({"left_fit": [fhex(c) for c in mod.left.fit], "right_fit": [fhex(c) for c in mod.right.fit],
                     "detected": [bool(mod.left.detected), bool(mod.right.detected)],
                     "n_fit_points": [int(len(mod.left.fity)), int(len(mod.right.fity))],
                     "offset": fhex(off), "mean_curv": fhex(curv),
                     "out_md5": hashlib.md5(out.tobytes()).hexdigest()})
    return recs


def main():
    import cv2
    P = R.load("Project.py", "_ref_project")
    Hd = R.load("Project_harder_challenge.py", "_ref_harder")
    imgs = {n: R.read_rgb(os.path.join(STILLS, n + ".jpg")) for n in PROJECT_STILLS + CH_STILLS}
    gold = {"versions": {"cv2": cv2.__version__, "numpy": np.__version__}, "stills": {}, "sequences": {}}
    for n, img in imgs.items():
        gold["stills"][n] = {
            "project": still_record(P, img, (215, 255), (145, 200)),       # Project.py:295-296
            "harder": still_record(Hd, img, (210, 255), (143, 200)),       # harder:330-331
        }
        print(n, gold["stills"][n]["project"]["status"], gold["stills"][n]["harder"]["status"], flush=True)
    # config-2 style sequence: the 8 project stills, seed 0, 40 frames (> 2*frame_num so the deques wrap)
    seq = list(synth_sequence([imgs[n] for n in PROJECT_STILLS], 40, seed=0))
    gold["sequences"]["project_seed0_40"] = {"stills": PROJECT_STILLS, "seed": 0, "records": sequence_records(P, seq)}
    # harder variant, includes test_ch6/test_ch8 (empty lane) AFTER good frames: exercises the
    # previous-points fallback of blind_search (harder:138-140)
    ch = ["test_ch0", "test_ch1", "test_ch6", "test_ch2", "test_ch8", "test_ch8", "test_ch3", "test_ch4",
          "test_ch5", "test_ch7", "test_ch9", "test_ch10", "test_ch11", "test_ch12", "test_ch13", "test_ch14"]
    seq = list(synth_sequence([imgs[n] for n in ch], 32, seed=3))
    gold["sequences"]["harder_seed3_32"] = {"stills": ch, "seed": 3, "records": sequence_records(Hd, seq)}
    seq = list(synth_sequence([imgs[n] for n in ch], 32, seed=3))
    gold["sequences"]["project_on_ch_seed3_32"] = {"stills": ch, "seed": 3, "records": sequence_records(P, seq)}
    with open(os.path.join(ROOT, "tests", "golden", "reference_golden.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print("written")


if __name__ == "__main__":
    main()
