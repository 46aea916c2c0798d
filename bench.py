#!/usr/bin/env python
"""bench.py -- 1280x720 frames/s of the lane-finding hot path (process_image) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): the 1260-frame project video.  The video file is absent from the
reference checkout, so the sequence is SYNTHETIC: the 8 project stills cycled, seeded integer x-shift
random walk + gain, Line tracking on, frame_num = 15.  One step = one pass of
process_image over the whole sequence (trackers reset at the start of the step).

  value  frames/s with the frames already resident in HBM (CUDA events, max over ranks)
  e2e    frames/s through ld_process_batch_host: pinned host frames in, pinned host frames out,
         H2D / kernels / D2H pipelined inside the timed call
  N > 1  weak scaling: one long video of N x 1260 frames, rank r owns frames [1260 r, 1260 (r+1));
         per-pixel stages run independently, the Line state blob is handed rank r-1 -> r over NCCL
         before the fit stage (advanced-lane-detection_b200/sharding.py)

--impl reference times the reference's CPU path (oracle/lane_port.py: the same cv2 / numpy calls in the
same order as Project.py) on the host cores, one process per core, on a bounded sample of the same frames.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "advanced-lane-detection_b200"
N_FRAMES = 1260
STILLS = ["straight_lines1", "straight_lines2", "test1", "test2", "test3", "test4", "test5", "test6"]
METRIC = "1280x720 frames/sec end-to-end (process_image, Line tracking on)"
UNIT = "frames/s"
WORKLOAD = "project_video 1260 frames 1280x720 pre-decoded (synthetic sequence from the 8 project stills, seed=rank), Line tracking, frame_num=15"


def load_stills():
    from PIL import Image
    out = []
    for n in STILLS:
        with Image.open(os.path.join(ROOT, "tests", "golden", "stills", n + ".jpg")) as im:
            out.append(np.asarray(im.convert("RGB")).copy())
    return out


def synth_frames(n, seed):
    """same generator as oracle.lane_port.synth_sequence (kept separate: the bench's GPU arm never imports oracle/)"""
    stills = load_stills()
    rng = np.random.default_rng(seed)
    out = np.empty((n, 720, 1280, 3), np.uint8)
    dx = 0
    lut_in = np.arange(256, dtype=np.float32)
    for i in range(n):
        base = stills[i % len(stills)]
        dx = int(np.clip(dx + rng.integers(-2, 3), -8, 8))
        gain = float(rng.uniform(0.9, 1.1))
        cols = np.clip(np.arange(1280) - dx, 0, 1279)
        lut = np.clip(np.rint(lut_in * np.float32(gain)), 0, 255).astype(np.uint8)
        out[i] = lut[base[:, cols]]
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.lines, self.proc = gpu, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1]))
                mx = max(mx, float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ CPU arm
_WORKER = {}


def _cpu_init(frames_per_proc):
    """per-process set-up, outside the timed region: decode the stills, synthesise this worker's frames, build the oracle"""
    import cv2
    cv2.setNumThreads(1)
    sys.path.insert(0, ROOT)
    from oracle import lane_port as O
    _WORKER["frames"] = synth_frames(frames_per_proc, os.getpid() % 9973)
    _WORKER["oracle"] = O.LaneOracle("project")


def _cpu_worker(_):
    import warnings
    orc = _WORKER["oracle"]
    orc.reset()
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for f in _WORKER["frames"]:
            orc.process(f)
    return time.perf_counter() - t0


def cpu_reference_run(steps, warmup, frames_per_proc=8):
    """whole-host aggregate of the reference CPU path: one single-threaded process per core, each on its own stream
    of pre-decoded frames held in RAM; a step = every process runs process_image over its frames once."""
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    procs = max(1, min(cores, 64))
    ctx = mp.get_context("spawn")
    times = []
    with ctx.Pool(procs, initializer=_cpu_init, initargs=(frames_per_proc,)) as pool:
        pool.map(_cpu_worker, range(procs))                      # touch every worker once (imports, first-call costs)
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, range(procs), chunksize=1)
            if s >= warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": procs * frames_per_proc * steps / total, "ms_per_step": 1e3 * total / steps, "cores": procs,
            "sample": "%d single-threaded processes x %d pre-decoded frames per step (oracle port: same cv2/numpy calls as Project.py)" % (procs, frames_per_proc),
            "frames_per_step": procs * frames_per_proc}


def cpu_img2binary_ms(n=4):
    """the reference's helper.img2binary (oracle port) on the host, ms per frame, one process"""
    from oracle import lane_port as O
    frames = synth_frames(n, 0)
    O.helper_img2binary(frames[0])
    t0 = time.perf_counter()
    for f in frames:
        O.helper_img2binary(f)
    return 1e3 * (time.perf_counter() - t0) / n


def cpu_baseline_inline(n=60):
    """single process, cv2 default threads, first n frames of the bench sequence"""
    import cv2
    import warnings
    from oracle import lane_port as O
    frames = synth_frames(n, 0)
    orc = O.LaneOracle("project")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        orc.process(frames[0])
        orc.reset()
        t0 = time.perf_counter()
        for f in frames:
            orc.process(f)
        dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": int(cv2.getNumThreads()), "kind": "port",
            "sample": "first %d frames of the workload, one process, cv2 %s threads=%d, numpy %s" % (n, cv2.__version__, cv2.getNumThreads(), np.__version__)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--frames", type=int, default=N_FRAMES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stages", action="store_true", help="also time each kernel stage (adds a `kernels` object)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if a.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(max(1, a.steps), max(0, a.warmup))
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": WORKLOAD, "frames_per_step": r["frames_per_step"]},
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"                   # NCCL prints its version banner on STDOUT; this script's stdout is one JSON line
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module(PKG)
    sharding = importlib.import_module(PKG + ".sharding")
    n = a.frames
    eng = pkg.LaneEngine("project", device=local, max_frames=n)
    frames_host = synth_frames(n, seed=rank)
    pin_in = torch.empty((n, 720, 1280, 3), dtype=torch.uint8, pin_memory=True)
    pin_in.numpy()[...] = frames_host
    del frames_host
    pin_out = torch.empty((n, 720, 1280, 3), dtype=torch.uint8, pin_memory=True)
    frames = pin_in.to("cuda", non_blocking=False)
    out = torch.empty_like(frames)
    fits = torch.empty((n, pkg._abi.FIT_DTYPE.itemsize), dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        eng.reset()
        if world == 1:
            eng.process_dev(frames, out=out, fits=fits)
        else:
            sharding.process_sharded(eng, frames, rank, world, out=out)

    # state reset; shared pass, k_mask_b, k_search; 5 fit kernels; k_raster; plain tiles, k_blend_px, text tiles
    launches_per_step = 1 + 3 + 5 + 1 + 3

    for _ in range(a.warmup):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        ev0.record()
        for _ in range(a.steps):
            step()
        ev1.record()
        barrier()
    eng.check()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * n * a.steps / (ms / 1e3)

    # ---- end to end through the C ABI host call (pinned host -> pinned host)
    eng.reset()
    eng.process_pinned(pin_in, pin_out)                      # warm-up
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(a.steps, 3))
    for _ in range(e2e_steps):
        eng.reset()
        eng.process_pinned(pin_in, pin_out)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = world * n * e2e_steps / float(t.item())
    same = bool((pin_out.to("cuda") == out).all().item()) if world == 1 else None

    # ---- per-kernel timings and rooflines, measured live (CUDA events on the launching stream, each kernel alone over
    #      the same n resident frames).  roofline = the kernel that dominates the step (the frame pass); mask_roofline =
    #      the fused mask kernel BASELINE.json's metric names.  Algorithmic bytes per frame: SURVEY.md 8(d) / DESIGN.md 4.
    peak, peak_src = peaks()
    alg = eng.plan.mask_algorithmic_bytes
    stage = {}

    def time_stage(name, fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        stage[name] = e0.elapsed_time(e1) / reps

    L, h, st = eng.lib, eng._ctx_h, eng._stream()
    mask_t = eng.mask_dev(frames)
    time_stage("k_mask", lambda: L.ld_mask(h, frames.data_ptr(), n, mask_t.data_ptr(), st))
    rec_t = eng.search_dev(mask_t)
    time_stage("k_search", lambda: L.ld_search(h, mask_t.data_ptr(), n, rec_t.data_ptr(), st))
    eng.reset()
    time_stage("k_fit*", lambda: L.ld_fit(h, rec_t.data_ptr(), n, fits.data_ptr(), st))
    time_stage("k_raster", lambda: L.ld_raster(h, fits.data_ptr(), n, st))
    # the launches the step is really made of (ld_stage): shared pass -> k_mask_b -> ... plain tiles, k_blend_px, text tiles
    SH, PL, BL, TX = "k_frame_pass_mf<3> (shared pass)", "k_frame_pass_mf<0> (plain tiles)", "k_blend_px", "k_frame_pass_mf<1> (text tiles)"
    time_stage(SH, lambda: L.ld_stage(h, 3, frames.data_ptr(), n, None, st))
    time_stage("k_mask_b", lambda: L.ld_stage(h, 7, None, n, None, st))
    time_stage(PL, lambda: L.ld_stage(h, 5, frames.data_ptr(), n, out.data_ptr(), st))
    time_stage(BL, lambda: L.ld_stage(h, 6, None, n, out.data_ptr(), st))
    time_stage(TX, lambda: L.ld_stage(h, 4, frames.data_ptr(), n, out.data_ptr(), st))
    n_share, n_plain, n_text = L.ld_stage_tiles(h, 3), L.ld_stage_tiles(h, 5), L.ld_stage_tiles(h, 4)
    time_stage("k_frame_pass_mf<1> (whole frame: ld_frame_pass / ld_overlay)", lambda: L.ld_frame_pass(h, frames.data_ptr(), n, out.data_ptr(), st))
    hist_t = eng.hist_dev(mask_t)
    time_stage("k_hist", lambda: L.ld_hist(h, mask_t.data_ptr(), n, hist_t.data_ptr(), st))
    und_t = torch.empty_like(frames)
    time_stage("k_frame_pass_mf<0> (cv2.undistort alone)", lambda: L.ld_undistort(h, frames.data_ptr(), n, und_t.data_ptr(), st))
    del und_t
    # helper.img2binary: gray -> per-image Sobel maxima -> thresholds, three kernels
    time_stage("img2binary (k_gray + k_sobel_max + k_grad_binary)",
               lambda: eng.grad_binary_dev(frames, eng.gray_dev(frames), s_thresh=(100, 255), sx_thresh=(20, 100), mag_thresh=(50, 255)), reps=3)
    eng.check()
    frame_bytes = 720 * 1280 * 3
    tile_px = 128 * 16
    bytes_per_frame = {"k_mask": alg, "k_search": 115200 + 560, "k_raster": 722 * 82 * 8,
                       "k_frame_pass_mf<1> (whole frame: ld_frame_pass / ld_overlay)": 2 * frame_bytes,
                       SH: n_share * tile_px * (3 + 4) + 43520,          # source pixels read, u32 pixels + decision bits written
                       "k_mask_b": 43520 + 115200, PL: n_plain * tile_px * 6, TX: n_text * tile_px * 6,
                       BL: n_share * tile_px * (4 + 3) + 722 * 82 * 8,   # u32 pixels + canvas read, RGB written
                       "k_hist": 115200 + 20480, "k_frame_pass_mf<0> (cv2.undistort alone)": 2 * frame_bytes,
                       # RGB read twice (gray, S channel), gray written once and read twice, binary written once
                       "img2binary (k_gray + k_sobel_max + k_grad_binary)": 2 * frame_bytes + 4 * 720 * 1280}
    kernels = {k: {"ms": v, "us_per_frame": 1e3 * v / n,
                   **({"gbs": bytes_per_frame[k] * n / (v * 1e-3) / 1e9, "frac": bytes_per_frame[k] * n / (v * 1e-3) / 1e9 / peak}
                      if k in bytes_per_frame else {})} for k, v in stage.items()}
    traffic = {}
    try:                                                   # DRAM bytes per frame from the committed `ncu --set full` captures
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f)
    except Exception:
        pass

    def roof(kernel, alg_bytes, traffic_key):
        ms_k = stage[kernel]
        gbs = alg_bytes * n / (ms_k * 1e-3) / 1e9
        tr = traffic.get(traffic_key)
        return {"kernel": kernel, "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                "traffic": tr * n if tr else None, "peak_source": peak_src, "algorithmic_bytes_per_frame": alg_bytes,
                "ms_per_launch": ms_k, "frames_per_launch": n, "share_of_step": ms_k / (ms / a.steps)}

    # the dominant launch of the step: the larger of the two big frame-pass launches
    dom = SH if stage[SH] >= stage[PL] else PL
    roofline = roof(dom, bytes_per_frame[dom], "k_frame_pass_shared_dram_bytes_per_frame" if dom == SH else "k_frame_pass_plain_dram_bytes_per_frame")
    mask_roofline = roof("k_mask", alg, "k_mask_dram_bytes_per_frame")
    mask_roofline["note"] = ("the fused mask kernel of ld_mask (north-star kernel 1), timed alone; inside ld_process_batch the mask comes "
                             "from the shared pass + k_mask_b")

    if rank == 0:
        cpu = None
        if not a.no_cpu_baseline and world == 1:
            cpu = cpu_baseline_inline()
            kernels["img2binary (k_gray + k_sobel_max + k_grad_binary)"]["cpu_port_ms_per_frame"] = cpu_img2binary_ms()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "frames_per_gpu": n, "l2": "inputs (3.5 GB/step) far larger than the 126 MB L2; no flush needed",
                           "parallelism": "frame-range sharding, Line state hand-off over NCCL" if world > 1 else "1 GPU"},
                "clocks": clocks.summary(),
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": n * frame_bytes, "d2h_bytes_per_step": n * frame_bytes,
                        "matches_device_path": same, "api": "ld_process_batch_host (pinned host in/out, 32-frame chunks, 3 streams)"},
                "gpu_launches": launches_per_step * a.steps,
                "roofline": roofline, "mask_roofline": mask_roofline, "kernels": kernels, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
