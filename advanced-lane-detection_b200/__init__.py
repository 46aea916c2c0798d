"""B200-native implementation of the per-frame lane-finding hot path of
uranus4ever/Advanced-Lane-Detection (Project.py: process_image and everything under it).

    engine   -- LaneEngine: plan + context over the C ABI of csrc/liblanedet.so (sm_100a kernels)
    plan     -- host-side table builder (calibration-like one-time step)
    Project, Project_harder_challenge, helper -- drop-in module surfaces of the reference scripts
                (imported lazily: they create a GPU context)
    sharding -- frame-range partition + tracker hand-off between GPUs

Importing the package loads the shared library and fails if it is missing: there is no CPU path.
"""
from . import _abi

_lib = _abi.load()

from .plan import PROJECT, HARDER, HELPER, VARIANTS, build_plan, LanePlan, Variant  # noqa: E402
from .engine import LaneEngine  # noqa: E402

__all__ = ["LaneEngine", "build_plan", "LanePlan", "Variant", "PROJECT", "HARDER", "HELPER", "VARIANTS"]
