"""ctypes binding of csrc/liblanedet.so (include/lanedet.h).  There is no fallback: if the library is
missing, cannot be loaded, or lacks a symbol, importing this module raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.environ.get("LANEDET_LIB") or os.path.join(CSRC, "liblanedet.so")   # LANEDET_LIB: instrumented debug builds only
HEADER = os.path.join(os.path.dirname(HERE), "include", "lanedet.h")

LD_OK, LD_ERR_ARG, LD_ERR_CUDA, LD_ERR_NO_DEVICE, LD_ERR_EMPTY_LANE, LD_ERR_POLYGON = range(6)
LD_MAX_DEPTH, LD_MAX_WINDOWS, LD_ROWMAP_WORDS, LD_MAX_CROSSINGS = 16, 9, 24, 8
LD_TEXT_X0, LD_TEXT_Y0, LD_TEXT_WORDS, LD_TEXT_ROWS = 32, 70, 24, 70
LD_GLYPH_ROWS, LD_GLYPH_WORDS, LD_TEXT_MAX_GLYPHS = 40, 12, 24

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-extended-lambda", "-Xcompiler", "-fPIC", "-shared"]


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -> csrc/liblanedet.so (in-tree, so it travels with the repo snapshot)."""
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(HEADER)
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, os.path.join(CSRC, "lanedet.cu")]
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB_PATH


class PlanDesc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("search_mode", C.c_int32), ("margin", C.c_int32),
                ("minpix", C.c_int32), ("frame_num", C.c_int32), ("top_anchor", C.c_int32),
                ("clip_polygon", C.c_int32), ("curv_fixed_xm", C.c_int32), ("n_tiles", C.c_int32),
                ("ym_per_pix", C.c_double),
                ("und_map", C.c_void_p), ("wtab", C.c_void_p), ("near_map", C.c_void_p), ("tiles", C.c_void_p),
                ("bmap", C.c_void_p), ("ctab", C.c_void_p), ("cbits", C.c_void_p),
                ("inv_y0", C.c_int32), ("inv_y1", C.c_int32),
                ("inv_idx", C.c_void_p), ("inv_fi", C.c_void_p), ("inv_pack", C.c_void_p), ("glyph_bits", C.c_void_p), ("glyph_adv", C.c_void_p),
                ("n_glyphs", C.c_int32), ("n_frame_tiles", C.c_int32),
                ("frame_tiles", C.c_void_p), ("mask_chunks", C.c_void_p), ("mask_chunk_start", C.c_void_p), ("frame_amap", C.c_void_p), ("frame_amap_start", C.c_void_p),
                ("mask_amap", C.c_void_p), ("mask_amap_start", C.c_void_p), ("bdesc", C.c_void_p),
                ("n_mask_chunks", C.c_int32), ("reserved", C.c_int32), ("cpre", C.c_void_p)]


class LaneRecord(C.Structure):
    _fields_ = [("mom", C.c_uint64 * 8), ("rowmap", C.c_uint32 * LD_ROWMAP_WORDS), ("ok", C.c_int32),
                ("n_windows", C.c_int32), ("win_base", C.c_int32 * LD_MAX_WINDOWS),
                ("win_n", C.c_int32 * LD_MAX_WINDOWS), ("win_kept", C.c_int32 * LD_MAX_WINDOWS), ("pad", C.c_int32)]


class FrameFit(C.Structure):
    _fields_ = [("fit", (C.c_double * 3) * 2), ("fit1", (C.c_double * 3) * 2), ("fit2", (C.c_double * 3) * 2),
                ("bottom", C.c_double * 2), ("top", C.c_double * 2), ("curv", C.c_double * 2),
                ("offset", C.c_double), ("mean_curv", C.c_double),
                ("rank1", C.c_int32 * 2), ("rank2", C.c_int32 * 2), ("detected", C.c_int32 * 2),
                ("n_points", C.c_int32 * 2), ("status", C.c_int32), ("n_text", C.c_int32 * 2),
                ("text", (C.c_uint8 * LD_TEXT_MAX_GLYPHS) * 2), ("rowmap", (C.c_uint32 * LD_ROWMAP_WORDS) * 2),
                ("near_integer", C.c_int32)]


RECORD_DTYPE = np.dtype(LaneRecord)
FIT_DTYPE = np.dtype(FrameFit)
assert RECORD_DTYPE.itemsize == 280 and FIT_DTYPE.itemsize == 496, (RECORD_DTYPE.itemsize, FIT_DTYPE.itemsize)

_P = C.c_void_p
_SIGNATURES = {
    "ld_abi_version": (C.c_int, []),
    "ld_error_string": (C.c_char_p, [C.c_int]),
    "ld_last_cuda_error": (C.c_char_p, []),
    "ld_plan_create": (C.c_int, [C.POINTER(PlanDesc), C.c_int, C.POINTER(_P)]),
    "ld_plan_destroy": (C.c_int, [_P]),
    "ld_ctx_create": (C.c_int, [_P, C.c_int, C.POINTER(_P)]),
    "ld_ctx_destroy": (C.c_int, [_P]),
    "ld_mask": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "ld_mask_fused": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "ld_hist": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "ld_search": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "ld_fit": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "ld_overlay": (C.c_int, [_P, _P, _P, C.c_int, _P, _P]),
    "ld_raster": (C.c_int, [_P, _P, C.c_int, _P]),
    "ld_frame_pass": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "ld_stage": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P]),
    "ld_stage_tiles": (C.c_int, [_P, C.c_int]),
    "ld_debug_knob": (C.c_int, [C.c_int, C.c_int]),
    "ld_ctx_mask": (_P, [_P]),
    "ld_process_batch": (C.c_int, [_P, _P, C.c_int, _P, _P, _P]),
    "ld_process_batch_host": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "ld_prepass": (C.c_int, [_P, _P, C.c_int, _P]),
    "ld_prepass_out": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "ld_postpass": (C.c_int, [_P, _P, C.c_int, _P, _P, _P, _P]),
    "ld_postpass_flags": (C.c_int, [_P, _P, C.c_int, _P, _P, _P, _P, C.c_int]),
    "ld_wait_state": (C.c_int, [_P, _P]),
    "ld_halo_set": (C.c_int, [_P, _P, C.c_int, _P]),
    "ld_ctx_records": (_P, [_P]),
    "ld_ctx_halo_slots": (_P, [_P, C.c_int]),
    "ld_halo_arrives": (C.c_int, [_P, C.c_int, _P]),
    "ld_tail_check": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "ld_halo_status": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ld_warmup": (C.c_int, [_P, _P, C.c_int, C.c_int, _P]),
    "ld_process_range": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "ld_process_range_host": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, C.c_int, _P, _P, C.POINTER(C.c_int)]),
    "ld_check": (C.c_int, [_P, _P, C.POINTER(C.c_int)]),
    "ld_state_reset": (C.c_int, [_P, _P]),
    "ld_state_size": (C.c_size_t, []),
    "ld_state_export": (C.c_int, [_P, _P, _P]),
    "ld_state_import": (C.c_int, [_P, _P, _P]),
    "ld_undistort": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "ld_warp": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    "ld_colour_mask": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "ld_pack_mask": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "ld_unpack_mask": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "ld_margin_search": (C.c_int, [_P, _P, _P, C.c_int, _P, _P]),
    "ld_polyfit": (C.c_int, [_P, _P, C.c_int, _P, _P, _P]),
    "ld_polyfit_points": (C.c_int, [_P, _P, _P, C.c_int, C.c_double, _P, _P, _P]),
    "ld_tracked_search": (C.c_int, [_P, _P, C.c_int, _P, _P, _P, _P, _P]),
    "ld_draw": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "ld_gray": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    "ld_sobel_max": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "ld_grad_binary": (C.c_int, [_P, _P, _P, _P, C.c_int, _P] + [C.c_int] * 7 + [_P, _P]),
    "ld_color_filter": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
}


def declared_symbols():
    """Every function include/lanedet.h declares (parsed from the header, so the CPU-only test can check
    that the built library exports exactly that surface)."""
    import re
    with open(HEADER) as f:
        text = f.read()
    return sorted(set(re.findall(r"\b(ld_[a-z_0-9]+)\s*\(", text)))


def load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)           # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.ld_abi_version() != 1:
        raise ImportError("liblanedet.so ABI version mismatch")
    return lib


class LaneDetError(RuntimeError):
    pass


def check(lib, status: int):
    if status == LD_OK:
        return
    if status == LD_ERR_EMPTY_LANE:
        # the reference's failure mode for an empty lane on a fresh tracker (Project.py:98-100 -> :118)
        raise TypeError("expected 1D vector for x")
    msg = lib.ld_error_string(status).decode()
    if status == LD_ERR_CUDA:
        msg += ": " + lib.ld_last_cuda_error().decode()
    raise LaneDetError(msg)
