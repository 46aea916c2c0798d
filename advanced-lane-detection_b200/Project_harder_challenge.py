"""Drop-in for the hot path of the reference's Project_harder_challenge.py (same surface, GPU underneath).
Differences from Project.py that this module carries: geometry (harder:412-415), thresholds (:330-331),
frame_num 5 (:426), the re-centring window search (:73-142), bottom anchor only (:161-171), the clipped
polygon (:192-197), per-lane curvature with 30/720 (:221-237) and car_pos returning the offset only (:240-254).
"""
import numpy as np

from . import _dropin
from .plan import HARDER, CAL_MTX, CAL_DIST

img_shape = (720, 1280)
img_size = [1280, 720]
src = np.float32(HARDER.src)
dst = np.float32(HARDER.dst)
ploty = np.linspace(0, img_shape[0] - 1, img_shape[0])
mtx, dist = CAL_MTX.copy(), CAL_DIST.copy()
frame_num = HARDER.frame_num

_pipeline = _dropin.Pipeline(HARDER, mtx, dist)
Line = _pipeline.Line
left, right = _pipeline.left, _pipeline.right


def reset():
    global left, right
    left, right = _pipeline.reset()
    return left, right


def process_image(img, plot=False):
    """Project_harder_challenge.py:324-407 (plot=True needs matplotlib; not part of the hot path)."""
    return _pipeline.process_image(img)


def process_batch(frames):
    return _pipeline.engine().process_host(frames)


def warp(img):
    return _pipeline.warp(img)


def luv_lab_filter(img, l_thresh=(195, 255), b_thresh=(140, 200), plot=False):
    return _pipeline.luv_lab_filter(img, l_thresh, b_thresh, plot, do_warp=True)


def undistort(img, mtx=None, dist=None):
    return _pipeline.undistort(img, mtx, dist)


def draw_area(undist, left_fitx, lefty, right_fitx, righty):
    return _pipeline.draw_area(undist, left_fitx, lefty, right_fitx, righty)


def curvature(fit):
    return _pipeline.curvature(fit)


def car_pos(left_fit, right_fit):
    return _pipeline.car_pos(left_fit, right_fit)
