"""helper.py's gradient / HLS / RGB thresholds (img2binary, sobelx_filter, color_filter, combine_bin).

CPU: oracle/lane_port.py against goldens produced by executing the unmodified reference (oracle/make_golden_grad.py),
and the integer arithmetic the kernels use (gray formula, Sobel with BORDER_REFLECT_101, the HLS-S colour table) against cv2.
GPU: the kernels, through the drop-in module and the C ABI, bit-exact against the goldens and the oracle."""
import hashlib
import importlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, PKG_NAME
from oracle import lane_port as O

NAMES = ["test1", "test4", "straight_lines2", "test_ch3", "test_ch11", "test_ch12"]


def sha(a):
    return hashlib.sha1(np.packbits(a.astype(bool), axis=1, bitorder="little").tobytes()).hexdigest()


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLDEN, "reference_golden_grad.json")) as f:
        return json.load(f)["stills"]


def undistorted(stills, name):
    orc = O.LaneOracle("project")
    return O.undistort(stills(name), orc.mtx, orc.dist)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_goldens(name, gold, stills):
    und, g = undistorted(stills, name), gold[name]
    for key, got in (("img2binary", O.helper_img2binary(und)), ("img2binary_90_30_120", O.helper_img2binary(und, (90, 255), (30, 120))),
                     ("sobelx_filter", O.helper_sobelx_filter(und)), ("color_filter", O.helper_color_filter(und)),
                     ("combine_bin", O.helper_combine_bin(und))):
        assert int(got.sum()) == g[key]["sum"], key
        assert sha(got) == g[key]["sha1"], key


def test_kernel_arithmetic_against_cv2(stills):
    import cv2
    # gray: (9798 R + 19235 G + 3735 B + 2^14) >> 15 over a 2^21-colour lattice plus the extremes
    rng = np.random.default_rng(0)
    c = rng.integers(0, 256, (1024, 2048, 3), dtype=np.uint8)
    c[0, :256, :] = np.arange(256, dtype=np.uint8)[:, None]
    want = cv2.cvtColor(c, cv2.COLOR_RGB2GRAY)
    ci = c.astype(np.int64)
    assert ((9798 * ci[..., 0] + 19235 * ci[..., 1] + 3735 * ci[..., 2] + 16384) >> 15 == want).all()
    # Sobel: integer 3x3 with BORDER_REFLECT_101
    gray = cv2.cvtColor(undistorted(stills, "test1"), cv2.COLOR_RGB2GRAY)
    p = np.pad(gray.astype(np.int64), 1, mode="reflect")
    sx = (p[:-2, 2:] + 2 * p[1:-1, 2:] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[1:-1, :-2] + p[2:, :-2])
    sy = (p[2:, :-2] + 2 * p[2:, 1:-1] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[:-2, 1:-1] + p[:-2, 2:])
    assert (sx == cv2.Sobel(gray, cv2.CV_64F, 1, 0)).all() and (sy == cv2.Sobel(gray, cv2.CV_64F, 0, 1)).all()
    # the S channel is a function of the colour alone
    plan = importlib.import_module(PKG_NAME + ".plan")
    und = undistorted(stills, "test4")
    tab = plan.hls_s_table()
    idx = (und[..., 0].astype(np.int64) << 16) | (und[..., 1].astype(np.int64) << 8) | und[..., 2]
    assert (tab[idx] == cv2.cvtColor(und, cv2.COLOR_RGB2HLS)[..., 2]).all()
    # ... and of its largest and smallest channel alone: the 64 KB table the kernel reads reproduces all 2^24 entries
    mm = plan.hls_s_minmax_table()
    c = np.arange(1 << 24, dtype=np.uint32)
    r, g, b = (c >> 16) & 255, (c >> 8) & 255, c & 255
    key = (np.maximum(np.maximum(r, g), b) << 8) | np.minimum(np.minimum(r, g), b)
    assert (mm[key] == tab).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_dropin_matches_goldens(name, gold, stills):
    Hp = importlib.import_module(PKG_NAME + ".helper")
    und, g = undistorted(stills, name), gold[name]
    for key, got in (("img2binary", Hp.img2binary(und)), ("img2binary_90_30_120", Hp.img2binary(und, (90, 255), (30, 120))),
                     ("sobelx_filter", Hp.sobelx_filter(und)), ("color_filter", Hp.color_filter(und)), ("combine_bin", Hp.combine_bin(und))):
        assert got.dtype == np.uint8 and got.shape == (720, 1280)
        assert int(got.sum()) == g[key]["sum"], key
        assert sha(got) == g[key]["sha1"], key


@pytest.mark.gpu
def test_gpu_batch_and_edge_cases(stills):
    import torch
    pkg = importlib.import_module(PKG_NAME)
    eng = pkg.LaneEngine("project", device=0, max_frames=4)
    rng = np.random.default_rng(5)
    frames = np.stack([undistorted(stills, "test2"), rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8),
                       np.full((720, 1280, 3), 77, np.uint8)])                       # a still, noise, a flat image (max gradient 0)
    t = eng.to_device(frames)
    got = eng.grad_binary_dev(t, eng.gray_dev(t), s_thresh=(100, 255), sx_thresh=(20, 100), mag_thresh=(50, 255)).cpu().numpy()
    for i in range(2):                                                               # per-IMAGE maxima: batching must not mix them
        assert (got[i] == O.helper_img2binary(frames[i])).all(), i
    import cv2
    s_flat = cv2.cvtColor(frames[2], cv2.COLOR_RGB2HLS)[..., 2]
    assert (got[2] == ((s_flat >= 100) & (s_flat <= 255))).all()                     # 0/0 gradients cast to 0 -> only the S term remains
    torch.cuda.synchronize()
    eng.close()
