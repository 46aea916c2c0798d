"""Drop-in for the hot path of the reference's Project.py: same module surface, GPU underneath.

    from moviepy.editor import VideoFileClip
    Project = importlib.import_module("advanced-lane-detection_b200.Project")
    VideoFileClip(path).fl_image(Project.process_image).write_videofile(out, audio=False)   # Project.py:354-361

Unlike the reference, importing this module does not process a video (the reference runs its
driver at import time, Project.py:357-361).  `process_batch` is the throughput entry: the same
function over many frames at once, trackers advanced in frame order.
"""
import numpy as np

from . import _dropin
from .plan import PROJECT, CAL_MTX, CAL_DIST

img_shape = (720, 1280)                                          # Project.py:336
img_size = [1280, 720]                                           # Project.py:337
src = np.float32(PROJECT.src)                                    # Project.py:339-340
dst = np.float32(PROJECT.dst)                                    # Project.py:341-342
ploty = np.linspace(0, img_shape[0] - 1, img_shape[0])           # Project.py:343
mtx, dist = CAL_MTX.copy(), CAL_DIST.copy()                      # wide_dist_pickle.p, Project.py:346-349
frame_num = PROJECT.frame_num                                    # Project.py:350

_pipeline = _dropin.Pipeline(PROJECT, mtx, dist)
Line = _pipeline.Line
left, right = _pipeline.left, _pipeline.right                    # Project.py:351-352


def reset():
    """`left = Line(); right = Line()` -- fresh trackers (host mirrors and the GPU state)."""
    global left, right
    left, right = _pipeline.reset()
    return left, right


def process_image(img):
    """Project.py:289-334."""
    return _pipeline.process_image(img)


def process_batch(frames):
    """process_image over uint8[n,720,1280,3] in frame order -> (uint8[n,720,1280,3], per-frame fit records)."""
    eng = _pipeline.engine()
    return eng.process_host(frames)


def warp(img):
    """Project.py:222-233."""
    return _pipeline.warp(img)


def luv_lab_filter(img, l_thresh=(195, 255), b_thresh=(140, 200), plot=False):
    """Project.py:236-271 (the plot branch needs matplotlib and is not part of the hot path)."""
    return _pipeline.luv_lab_filter(img, l_thresh, b_thresh, plot, do_warp=True)


def undistort(img, mtx=None, dist=None):
    """Project.py:274-286."""
    return _pipeline.undistort(img, mtx, dist)


def draw_area(undist, left_fitx, lefty, right_fitx, righty):
    """Project.py:144-170."""
    return _pipeline.draw_area(undist, left_fitx, lefty, right_fitx, righty)


def curvature(fit):
    """Project.py:173-189."""
    return _pipeline.curvature(fit)


def car_pos(left_fit, right_fit):
    """Project.py:192-219 -> (offset, mean_curv)."""
    return _pipeline.car_pos(left_fit, right_fit)
