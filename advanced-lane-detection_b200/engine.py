"""LaneEngine: one plan + one context (= one video stream on one GPU) behind a small Python API.

torch is used for what it is good at here -- device/pinned memory and streams; every computation
goes through the C ABI of csrc/liblanedet.so (include/lanedet.h).  All *_dev methods take and
return CUDA tensors and are asynchronous on torch's current stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _abi
from .plan import LanePlan, Variant, VARIANTS, build_plan


def _ptr(t):
    return C.c_void_p(t.data_ptr())


class LaneEngine:
    def __init__(self, variant="project", mtx=None, dist=None, size=(1280, 720), device: Optional[int] = None,
                 max_frames: int = 64, l_thresh=None, b_thresh=None, plan: Optional[LanePlan] = None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("advanced-lane-detection_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.torch = torch
        self.lib = _abi.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        if plan is None:
            v = VARIANTS[variant] if isinstance(variant, str) else variant
            plan = build_plan(v, mtx, dist, size, l_thresh, b_thresh)
        self.plan = plan
        self.variant: Variant = plan.variant
        self.W, self.H = plan.size
        self.WW = self.W // 32
        self.max_frames = int(max_frames)
        self._keep = []                     # host arrays referenced by the descriptor
        self._plan_h = C.c_void_p()
        self._ctx_h = C.c_void_p()
        _abi.check(self.lib, self.lib.ld_plan_create(C.byref(self._descriptor(plan)), self.device, C.byref(self._plan_h)))
        _abi.check(self.lib, self.lib.ld_ctx_create(self._plan_h, self.max_frames, C.byref(self._ctx_h)))
        self._pin_in = self._pin_out = self._pin_fit = None

    # ------------------------------------------------------------------ plumbing
    def _descriptor(self, p: LanePlan) -> _abi.PlanDesc:
        v = p.variant

        def host(a, dtype):
            a = np.ascontiguousarray(a, dtype=dtype)
            self._keep.append(a)
            return a.ctypes.data

        d = _abi.PlanDesc()
        d.width, d.height = p.size
        d.search_mode, d.margin, d.minpix, d.frame_num = v.search, v.margin, v.minpix, v.frame_num
        d.top_anchor, d.clip_polygon, d.curv_fixed_xm = int(v.top_anchor), int(v.clip_polygon), int(v.curv_fixed_xm)
        d.n_tiles = len(p.tiles)
        d.ym_per_pix = v.ym_per_pix
        d.und_map, d.wtab = host(p.und_map, np.uint32), host(p.wtab, np.int16)
        d.near_map, d.tiles, d.bmap = host(p.near, np.int32), host(p.tiles, np.int32), host(p.bmap, np.uint16)
        d.ctab, d.cbits = host(p.ctab, np.uint32), host(p.cbits, np.uint32)
        d.cpre = host(p.cpre, np.uint32)
        d.inv_y0, d.inv_y1 = p.inv_y0, p.inv_y1
        d.inv_idx, d.inv_fi = host(p.inv_idx, np.uint32), host(p.inv_fi, np.uint16)
        d.inv_pack = host(p.inv_pack, np.uint32)
        d.glyph_bits, d.glyph_adv = host(p.glyph_bits, np.uint32), host(p.glyph_adv, np.int32)
        d.n_glyphs = len(p.glyph_adv)
        d.n_frame_tiles, d.frame_tiles = len(p.frame_tiles), host(p.frame_tiles, np.int32)
        d.n_mask_chunks, d.mask_chunks = len(p.mask_chunks), host(p.mask_chunks, np.int32)
        d.mask_chunk_start = host(p.mask_chunk_start, np.int32)
        d.bdesc = host(p.bdesc, np.uint32)
        d.frame_amap, d.frame_amap_start = host(p.frame_amap, np.uint32), host(p.frame_amap_start, np.int32)
        d.mask_amap, d.mask_amap_start = host(p.mask_amap, np.uint32), host(p.mask_amap_start, np.int32)
        return d

    def close(self):
        if getattr(self, "_ctx_h", None) and self._ctx_h.value:
            self.lib.ld_ctx_destroy(self._ctx_h)
            self._ctx_h = C.c_void_p()
        if getattr(self, "_plan_h", None) and self._plan_h.value:
            self.lib.ld_plan_destroy(self._plan_h)
            self._plan_h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, shape, dtype):
        return self.torch.empty(shape, dtype=dtype, device="cuda:%d" % self.device)

    def _frames(self, t, channels=3):
        torch = self.torch
        if t.dtype != torch.uint8 or not t.is_cuda or not t.is_contiguous():
            raise ValueError("expected a contiguous CUDA uint8 tensor")
        if t.dim() == (3 if channels > 1 else 2):
            t = t.unsqueeze(0)
        want = (self.H, self.W, channels) if channels > 1 else (self.H, self.W)
        if tuple(t.shape[1:]) != want:
            raise ValueError("expected frames of shape %s, got %s" % (want, tuple(t.shape[1:])))
        return t

    def to_device(self, a):
        return self.torch.as_tensor(np.ascontiguousarray(a)).to("cuda:%d" % self.device)

    def check(self):
        """Wait for the current stream; raise TypeError like the reference if a tracker had no points."""
        bad = C.c_int(-1)
        _abi.check(self.lib, self.lib.ld_check(self._ctx_h, self._stream(), C.byref(bad)))

    # ------------------------------------------------------------------ hot path, stage by stage
    def mask_dev(self, frames, fused=False):
        """undistort + warp + luv_lab_filter -> packed mask int32[n,H,W/32] (bit i of word k = pixel 32k+i).
        fused=True: the one-kernel route (ld_mask_fused) instead of the plan's default."""
        frames = self._frames(frames)
        out = self._dev((frames.shape[0], self.H, self.WW), self.torch.int32)
        fn = self.lib.ld_mask_fused if fused else self.lib.ld_mask
        _abi.check(self.lib, fn(self._ctx_h, _ptr(frames), frames.shape[0], _ptr(out), self._stream()))
        return out

    def hist_dev(self, mask):
        out = self._dev((mask.shape[0], 8, self.W), self.torch.int16)
        _abi.check(self.lib, self.lib.ld_hist(self._ctx_h, _ptr(mask), mask.shape[0], _ptr(out), self._stream()))
        return out

    def search_dev(self, mask):
        out = self._dev((mask.shape[0], 2, _abi.RECORD_DTYPE.itemsize), self.torch.uint8)
        _abi.check(self.lib, self.lib.ld_search(self._ctx_h, _ptr(mask), mask.shape[0], _ptr(out), self._stream()))
        return out

    def fit_dev(self, records):
        n = records.shape[0]
        out = self._dev((n, _abi.FIT_DTYPE.itemsize), self.torch.uint8)
        for a in range(0, n, self.max_frames):
            b = min(n, a + self.max_frames)
            _abi.check(self.lib, self.lib.ld_fit(self._ctx_h, _ptr(records[a:b]), b - a, _ptr(out[a:b]), self._stream()))
        return out

    def overlay_dev(self, frames, fits):
        frames = self._frames(frames)
        out = self.torch.empty_like(frames)
        _abi.check(self.lib, self.lib.ld_overlay(self._ctx_h, _ptr(frames), _ptr(fits), frames.shape[0], _ptr(out),
                                                 self._stream()))
        return out

    def process_dev(self, frames, out=None, fits=None):
        """process_image for a batch of device-resident frames, in frame order."""
        frames = self._frames(frames)
        n = frames.shape[0]
        if out is None:
            out = self.torch.empty_like(frames)
        if fits is None:
            fits = self._dev((n, _abi.FIT_DTYPE.itemsize), self.torch.uint8)
        _abi.check(self.lib, self.lib.ld_process_batch(self._ctx_h, _ptr(frames), n, _ptr(out), _ptr(fits), self._stream()))
        return out, fits

    def prepass_dev(self, frames, out=None):
        """mask + search of a frame range into the context workspace (no tracker access).  out = the buffer the following
        postpass_dev will fill, if it is already known: the pixels under the horizon that the inverse warp cannot touch are
        final after the undistortion and go straight there (ld_prepass_out)."""
        frames = self._frames(frames)
        if out is None:
            _abi.check(self.lib, self.lib.ld_prepass(self._ctx_h, _ptr(frames), frames.shape[0], self._stream()))
        else:
            _abi.check(self.lib, self.lib.ld_prepass_out(self._ctx_h, _ptr(frames), frames.shape[0], _ptr(out), self._stream()))

    def context_masks(self, n):
        """copy of the packed masks the last prepass / process call left in the context: int32[n, H, W/32] (device)"""
        out = self._dev((n, self.H, self.WW), self.torch.int32)
        out.copy_(self._raw_view(self.lib.ld_ctx_mask(self._ctx_h), (n, self.H, self.WW), "<i4"))
        return out

    def plain_pass_dev(self, frames, out):
        """The tiles of the range given to prepass_dev that depend on neither the fit nor the canvas (pure cv2.undistort),
        written into `out` ahead of postpass_dev(..., skip_plain=True).  Returns False when the plan has no such split."""
        frames = self._frames(frames)
        return self.lib.ld_stage(self._ctx_h, 5, _ptr(frames), frames.shape[0], _ptr(out), self._stream()) == _abi.LD_OK

    def _raw_view(self, ptr, shape, typestr):
        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}
        return self.torch.as_tensor(raw, device="cuda:%d" % self.device)

    def context_records(self, n):
        """VIEW (no copy) of the lane records the last prepass left in the context: uint8[n, 2, 280] on the device"""
        return self._raw_view(self.lib.ld_ctx_records(self._ctx_h), (n, 2, _abi.RECORD_DTYPE.itemsize), "|u1")

    def halo_set(self, records):
        """lane records (uint8[h, 2, 280], device) of the h frames just before the range given to prepass_dev; the next
        postpass_dev(use_halo=True) rebuilds the trackers from them instead of importing a predecessor's state"""
        records = records.contiguous()
        _abi.check(self.lib, self.lib.ld_halo_set(self._ctx_h, _ptr(records), records.shape[0], self._stream()))

    is_cuda = True

    def new_like(self, frames):
        return self.torch.empty_like(frames)

    def new_fits(self, n):
        return self._dev((n, _abi.FIT_DTYPE.itemsize), self.torch.uint8)

    def new_state_blob(self):
        """device buffer for one tracker blob (export_state_dev / import_state_dev)"""
        return self._dev((self.lib.ld_state_size(),), self.torch.uint8)

    def records_ok(self, n):
        """host int32[n, 2]: the `ok` flags of the records the last prepass left in the context (synchronises)"""
        off = _abi.RECORD_DTYPE.fields["ok"][1]
        r = self.context_records(n)[:, :, off:off + 4].contiguous().cpu().numpy()
        return r.view(np.int32).reshape(n, 2)

    def halo_slots(self, h):
        """VIEW of the h record slots in front of the context's records (uint8[h, 2, 280], device): a transfer may deliver the
        look-back records there directly; announce it with halo_arrives(h, stream)"""
        ptr = self.lib.ld_ctx_halo_slots(self._ctx_h, int(h))
        if not ptr:
            raise ValueError("halo of %d frames is not supported" % h)
        return self._raw_view(ptr, (h, 2, _abi.RECORD_DTYPE.itemsize), "|u1")

    def halo_arrives(self, h, stream):
        """the h look-back records are being written into halo_slots(h) by work queued on `stream` (torch stream): the next
        postpass_dev(use_halo=True) uses them; only its fit stage waits for that work"""
        _abi.check(self.lib, self.lib.ld_halo_arrives(self._ctx_h, int(h), C.c_void_p(stream.cuda_stream)))

    def tail_check(self, n, h):
        """decide on the device whether the last h of the n records of the last prepass_dev are a usable halo for the successor"""
        _abi.check(self.lib, self.lib.ld_tail_check(self._ctx_h, int(n), int(h), self._stream()))

    def halo_status(self):
        """(halo_invalid, tail_invalid) of the last step; waits for those two device decisions only"""
        a, b = C.c_int(0), C.c_int(0)
        _abi.check(self.lib, self.lib.ld_halo_status(self._ctx_h, C.byref(a), C.byref(b)))
        return bool(a.value), bool(b.value)

    def warmup_dev(self, lookback, starts_video=False):
        """Zero-communication sharding: rebuild the trackers from the frames just before this rank's range (mask + search + fit
        from empty trackers, nothing drawn).  The device starts the trackers at the latest look-back frame from which the
        result is exact (frame_num frames without an empty lane, 2 (frame_num - 1) frames before the range) -- or at frame 0 when
        the look-back starts the video.  Verdict (an exact start exists): halo_status()[0] is False."""
        lookback = self._frames(lookback)
        _abi.check(self.lib, self.lib.ld_warmup(self._ctx_h, _ptr(lookback), lookback.shape[0], 0 if starts_video else 1, self._stream()))

    def process_range_dev(self, lookback, frames, out=None, fits=None, starts_video=False):
        """warmup_dev(lookback) + process_dev(frames).  When `lookback` ends exactly where `frames` begins in memory (two views of
        one buffer, lookback + range <= max_frames) both run as ONE pass without a single extra launch (ld_process_range)."""
        frames = self._frames(frames)
        n = frames.shape[0]
        if out is None:
            out = self.torch.empty_like(frames)
        if fits is None:
            fits = self.new_fits(n)
        if lookback is None or lookback.shape[0] == 0:
            self.reset()
            return self.process_dev(frames, out=out, fits=fits)
        lookback = self._frames(lookback)
        h = lookback.shape[0]
        if lookback.data_ptr() + lookback.numel() == frames.data_ptr() and h + n <= self.max_frames:
            _abi.check(self.lib, self.lib.ld_process_range(self._ctx_h, _ptr(lookback), h, 0 if starts_video else 1, n, _ptr(out), _ptr(fits),
                                                           self._stream()))
            return out, fits
        self.warmup_dev(lookback, starts_video=starts_video)
        return self.process_dev(frames, out=out, fits=fits)

    def postpass_dev(self, frames, out=None, fits=None, state_out=None, skip_plain=False, fit_first=False, use_halo=False):
        """fit (tracker order!) + overlay of the range given to prepass_dev; state_out = CUDA uint8 blob or None"""
        frames = self._frames(frames)
        n = frames.shape[0]
        if out is None:
            out = self.torch.empty_like(frames)
        if fits is None:
            fits = self._dev((n, _abi.FIT_DTYPE.itemsize), self.torch.uint8)
        _abi.check(self.lib, self.lib.ld_postpass_flags(self._ctx_h, _ptr(frames), n, _ptr(out), _ptr(fits),
                                                        _ptr(state_out) if state_out is not None else None, self._stream(),
                                                        (1 if skip_plain else 0) | (2 if fit_first else 0) | (4 if use_halo else 0)))
        return out, fits

    def wait_state(self, stream):
        _abi.check(self.lib, self.lib.ld_wait_state(self._ctx_h, C.c_void_p(stream.cuda_stream)))

    def process_host(self, frames: np.ndarray, want_fits: bool = True):
        """process_image for a batch of host frames (uint8 [n,H,W,3]); copies in/out are pipelined inside."""
        torch = self.torch
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        if frames.ndim == 3:
            frames = frames[None]
        n = frames.shape[0]
        if frames.shape[1:] != (self.H, self.W, 3):
            raise ValueError("expected frames of shape (n, %d, %d, 3)" % (self.H, self.W))
        if self._pin_in is None or self._pin_in.shape[0] < n:
            self._pin_in = torch.empty((n, self.H, self.W, 3), dtype=torch.uint8, pin_memory=True)
            self._pin_out = torch.empty((n, self.H, self.W, 3), dtype=torch.uint8, pin_memory=True)
            self._pin_fit = torch.empty((n, _abi.FIT_DTYPE.itemsize), dtype=torch.uint8, pin_memory=True)
        self._pin_in[:n].numpy()[...] = frames
        self.process_pinned(self._pin_in[:n], self._pin_out[:n], self._pin_fit[:n] if want_fits else None)
        out = self._pin_out[:n].numpy().copy()
        fits = self._pin_fit[:n].numpy().copy().view(_abi.FIT_DTYPE).reshape(n) if want_fits else None
        return out, fits

    def _host_frames(self, t, what):
        """host frame buffers go to the C ABI as raw pointers: refuse anything that is not a dense uint8 [n, H, W, 3] block"""
        if t.is_cuda or t.dtype != self.torch.uint8 or not t.is_contiguous() or tuple(t.shape[1:]) != (self.H, self.W, 3):
            raise ValueError("%s: expected a contiguous host uint8 tensor [n, %d, %d, 3]" % (what, self.H, self.W))
        return t

    def process_pinned(self, pin_in, pin_out, pin_fit=None):
        """The timed end-to-end call of bench.py: pinned host in -> pinned host out, synchronous."""
        self._host_frames(pin_in, "pin_in")
        self._host_frames(pin_out, "pin_out")
        n = pin_in.shape[0]
        if pin_out.shape[0] < n or (pin_fit is not None and (not pin_fit.is_contiguous() or pin_fit.numel() < n * _abi.FIT_DTYPE.itemsize)):
            raise ValueError("output buffers are smaller than the input")
        st = self.lib.ld_process_batch_host(self._ctx_h, _ptr(pin_in), n, _ptr(pin_out),
                                            _ptr(pin_fit) if pin_fit is not None else None)
        _abi.check(self.lib, st)

    def process_range_pinned(self, pin_lookback, pin_in, pin_out, pin_fit=None, starts_video=False):
        """process_pinned for one rank's range of a sharded video: pin_lookback = the frames before the range (None: the range
        starts the video), warmed up through the same pipeline.  Returns True if no exact tracker start exists in the
        look-back (see warmup_dev): pass a longer one."""
        h = 0 if pin_lookback is None else pin_lookback.shape[0]
        self._host_frames(pin_in, "pin_in")
        self._host_frames(pin_out, "pin_out")
        if h:
            self._host_frames(pin_lookback, "pin_lookback")
        if pin_out.shape[0] < pin_in.shape[0] or (pin_fit is not None and (not pin_fit.is_contiguous() or
                                                                            pin_fit.numel() < pin_in.shape[0] * _abi.FIT_DTYPE.itemsize)):
            raise ValueError("output buffers are smaller than the input")
        bad = C.c_int(0)
        st = self.lib.ld_process_range_host(self._ctx_h, _ptr(pin_lookback) if h else None, h, 0 if (starts_video or not h) else 1,
                                            _ptr(pin_in), pin_in.shape[0], _ptr(pin_out), _ptr(pin_fit) if pin_fit is not None else None,
                                            C.byref(bad))
        _abi.check(self.lib, st)
        return bool(bad.value)

    # ------------------------------------------------------------------ tracker state
    def reset(self):
        _abi.check(self.lib, self.lib.ld_state_reset(self._ctx_h, self._stream()))

    def export_state(self) -> bytes:
        buf = C.create_string_buffer(self.lib.ld_state_size())
        _abi.check(self.lib, self.lib.ld_state_export(self._ctx_h, buf, self._stream()))
        return buf.raw

    def import_state(self, blob: bytes):
        if len(blob) != self.lib.ld_state_size():
            raise ValueError("tracker state blob has the wrong size")
        _abi.check(self.lib, self.lib.ld_state_import(self._ctx_h, C.create_string_buffer(blob, len(blob)), self._stream()))

    def export_state_dev(self, blob_dev):
        """tracker blob into a CUDA uint8 tensor (what ranks exchange over NVLink)"""
        _abi.check(self.lib, self.lib.ld_state_export(self._ctx_h, _ptr(blob_dev), self._stream()))
        return blob_dev

    def import_state_dev(self, blob_dev):
        _abi.check(self.lib, self.lib.ld_state_import(self._ctx_h, _ptr(blob_dev), self._stream()))

    def overlay_into(self, frames, fits, out):
        frames = self._frames(frames)
        _abi.check(self.lib, self.lib.ld_overlay(self._ctx_h, _ptr(frames), _ptr(fits), frames.shape[0], _ptr(out),
                                                 self._stream()))
        return out

    # ------------------------------------------------------------------ helper.py surface
    def undistort_dev(self, frames):
        frames = self._frames(frames)
        out = self.torch.empty_like(frames)
        _abi.check(self.lib, self.lib.ld_undistort(self._ctx_h, _ptr(frames), frames.shape[0], _ptr(out), self._stream()))
        return out

    def warp_dev(self, img):
        channels = 3 if (img.dim() == 4 or (img.dim() == 3 and img.shape[-1] == 3)) else 1
        img = self._frames(img, channels)
        out = self.torch.empty_like(img)
        _abi.check(self.lib, self.lib.ld_warp(self._ctx_h, _ptr(img), img.shape[0], channels, _ptr(out), self._stream()))
        return out

    def colour_mask_dev(self, img):
        img = self._frames(img)
        out = self._dev((img.shape[0], self.H, self.W), self.torch.uint8)
        _abi.check(self.lib, self.lib.ld_colour_mask(self._ctx_h, _ptr(img), img.shape[0], _ptr(out), self._stream()))
        return out

    def pack_dev(self, binary):
        binary = self._frames(binary, 1)
        out = self._dev((binary.shape[0], self.H, self.WW), self.torch.int32)
        _abi.check(self.lib, self.lib.ld_pack_mask(self._ctx_h, _ptr(binary), binary.shape[0], _ptr(out), self._stream()))
        return out

    def unpack_dev(self, mask):
        out = self._dev((mask.shape[0], self.H, self.W), self.torch.uint8)
        _abi.check(self.lib, self.lib.ld_unpack_mask(self._ctx_h, _ptr(mask), mask.shape[0], _ptr(out), self._stream()))
        return out

    def margin_search_dev(self, mask, fits):
        """helper.skip_sliding_window: fits = float64[n,2,3] (left, right)."""
        out = self._dev((mask.shape[0], 2, _abi.RECORD_DTYPE.itemsize), self.torch.uint8)
        _abi.check(self.lib, self.lib.ld_margin_search(self._ctx_h, _ptr(mask), _ptr(fits), mask.shape[0], _ptr(out),
                                                       self._stream()))
        return out

    def tracked_search_dev(self, mask, init_fits=None):
        """helper.skip_sliding_window as a tracked search over a sequence of packed masks (int32[n,H,W/32], frame order): frame k's
        fits set frame k+1's bands, on the device.  init_fits = float64[2,3] (CUDA) or None.  Returns (records uint8[n,2,280],
        fits float64[n,2,3], source int32[n]: 0 band search, 1 window-search fallback, 2 neither found 10 pixels per lane)."""
        n = mask.shape[0]
        rec = self._dev((n, 2, _abi.RECORD_DTYPE.itemsize), self.torch.uint8)
        fits = self._dev((n, 2, 3), self.torch.float64)
        source = self._dev((n,), self.torch.int32)
        _abi.check(self.lib, self.lib.ld_tracked_search(self._ctx_h, _ptr(mask), n, _ptr(init_fits) if init_fits is not None else None,
                                                        _ptr(rec), _ptr(fits), _ptr(source), self._stream()))
        return rec, fits, source

    def polyfit_points(self, y, x):
        """np.polyfit(y, x, 2) of host point arrays on the device (the stand-alone Line.get_fit): returns (float64[3], rank).
        rcond follows numpy: len(y) * eps of y's floating dtype (float64 for integer input)."""
        y, x = np.asarray(y), np.asarray(x)
        if y.ndim != 1:
            raise TypeError("expected 1D vector for x")              # numpy's message: polyfit's `x` is our y
        if x.ndim != 1 or x.shape[0] != y.shape[0]:
            raise TypeError("expected x and y to have same length")
        if y.size == 0:
            raise TypeError("expected non-empty vector for x")
        eps = float(np.finfo(y.dtype).eps) if np.issubdtype(y.dtype, np.floating) else float(np.finfo(np.float64).eps)
        dev = "cuda:%d" % self.device
        yd = self.torch.as_tensor(np.ascontiguousarray(y, np.float64)).to(dev)
        xd = self.torch.as_tensor(np.ascontiguousarray(x, np.float64)).to(dev)
        coef = self._dev((3,), self.torch.float64)
        rank = self._dev((1,), self.torch.int32)
        _abi.check(self.lib, self.lib.ld_polyfit_points(self._ctx_h, _ptr(xd), _ptr(yd), int(y.shape[0]), eps, _ptr(coef), _ptr(rank),
                                                        self._stream()))
        return coef.cpu().numpy(), int(rank.item())

    def polyfit_dev(self, records):
        n = records.numel() // _abi.RECORD_DTYPE.itemsize
        coef = self._dev((n, 3), self.torch.float64)
        rank = self._dev((n,), self.torch.int32)
        _abi.check(self.lib, self.lib.ld_polyfit(self._ctx_h, _ptr(records), n, _ptr(coef), _ptr(rank), self._stream()))
        return coef, rank

    def draw_dev(self, undist, verts, counts, thickness=30):
        """draw_area on already-undistorted frames with explicit integer polygon vertices
        (verts int32[n,max_verts,2], counts int32[n])."""
        undist = self._frames(undist)
        out = self.torch.empty_like(undist)
        _abi.check(self.lib, self.lib.ld_draw(self._ctx_h, _ptr(undist), _ptr(verts), _ptr(counts), verts.shape[1],
                                              thickness, undist.shape[0], _ptr(out), self._stream()))
        return out

    # ------------------------------------------------------------------ helper.py gradient / HLS / RGB thresholds
    def _s_table_dev(self):
        if getattr(self, "_s_table", None) is None:
            from .plan import hls_s_minmax_table
            self._s_table = self.torch.from_numpy(hls_s_minmax_table()).to("cuda:%d" % self.device)
        return self._s_table

    def gray_dev(self, img, warped=False):
        """cv2.cvtColor(img, RGB2GRAY) (helper.py:59); of warp(img) when warped (helper.py:109-111)."""
        img = self._frames(img)
        out = self._dev((img.shape[0], self.H, self.W), self.torch.uint8)
        _abi.check(self.lib, self.lib.ld_gray(self._ctx_h, _ptr(img), img.shape[0], int(bool(warped)), _ptr(out), self._stream()))
        return out

    def grad_binary_dev(self, img, gray, s_thresh=None, sx_thresh=(20, 100), mag_thresh=None):
        """img2binary's combined binary (s_thresh and mag_thresh given) or sobelx_filter's abs_bin (both None)."""
        img = self._frames(img)
        n = gray.shape[0]
        mx = self._dev((n, 2), self.torch.int32)
        _abi.check(self.lib, self.lib.ld_sobel_max(self._ctx_h, _ptr(gray), n, _ptr(mx), self._stream()))
        out = self._dev((n, self.H, self.W), self.torch.uint8)
        st = self._s_table_dev() if s_thresh is not None else None
        s0, s1 = (int(s_thresh[0]), int(s_thresh[1])) if s_thresh is not None else (0, 0)
        m0, m1 = (int(mag_thresh[0]), int(mag_thresh[1])) if mag_thresh is not None else (0, 0)
        _abi.check(self.lib, self.lib.ld_grad_binary(self._ctx_h, _ptr(img), _ptr(gray), _ptr(mx), n,
                                                     _ptr(st) if st is not None else None, s0, s1, int(sx_thresh[0]), int(sx_thresh[1]),
                                                     int(mag_thresh is not None), m0, m1, _ptr(out), self._stream()))
        return out

    def color_filter_dev(self, img, r_th, g_th, b_th, and_with=None):
        """color_filter (helper.py:493-500); AND-ed with `and_with` (combine_bin, helper.py:503-511) when given."""
        img = self._frames(img)
        out = self._dev((img.shape[0], self.H, self.W), self.torch.uint8)
        _abi.check(self.lib, self.lib.ld_color_filter(self._ctx_h, _ptr(img), _ptr(and_with) if and_with is not None else None,
                                                      img.shape[0], int(r_th), int(g_th), int(b_th), _ptr(out), self._stream()))
        return out

    # ------------------------------------------------------------------ host views
    @staticmethod
    def records_to_numpy(t):
        a = t.cpu().numpy()
        return a.view(_abi.RECORD_DTYPE).reshape(a.shape[:-1])

    @staticmethod
    def fits_to_numpy(t):
        a = t.cpu().numpy()
        return a.view(_abi.FIT_DTYPE).reshape(a.shape[:-1])
