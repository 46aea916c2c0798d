#!/usr/bin/env python
"""bench.py -- 1280x720 frames/s of the lane-finding hot path (process_image) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): the 1260-frame project video.  The video file is absent from the
reference checkout, so the sequence is SYNTHETIC: the 8 project stills cycled, seeded integer x-shift
random walk + gain, Line tracking on, frame_num = 15.  One step = one pass of
process_image over the whole sequence (trackers reset at the start of the step).

  value  frames/s with the frames already resident in HBM (CUDA events, max over ranks)
  e2e    frames/s through the C ABI host call: pinned host frames in, pinned host frames out,
         H2D / kernels / D2H pipelined inside the timed call (ld_process_batch_host; ld_process_range_host at N > 1)
  N > 1  weak scaling: ONE video of N x 1260 frames, rank r owns frames [1260 r, 1260 (r+1)) plus the 30 frames before
         them as look-back; sharding.ShardedRun, default protocol "lookback" (zero communication, every rank rebuilds
         the Line trackers from its look-back); `sharding` reports the other two protocols (look-back records over
         NCCL/NVLink, serial tracker blob) and `shard_matches_serial` compares all of them byte for byte
  configs  the other BASELINE.json configs in the same line: c1 single-frame process_image latency through the drop-in,
         c3 harder pipeline (frame_num 5) sharded the same way, c4 3840x2160 mask + histogram, c5 independent streams

--impl reference times the reference's CPU path (oracle/lane_port.py: the same cv2 / numpy calls in the
same order as Project.py) on the host cores, one process per core, on the first frames of the same sequence.
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
HALO = 30
STILLS = ["straight_lines1", "straight_lines2", "test1", "test2", "test3", "test4", "test5", "test6"]
# the harder pipeline's stills in the order of tests/golden (two of them have an empty lane: the fallback path runs every cycle)
HARDER_STILLS = ["test_ch0", "test_ch1", "test_ch6", "test_ch2", "test_ch8", "test_ch8", "test_ch3", "test_ch4", "test_ch5", "test_ch7",
                 "test_ch9", "test_ch10", "test_ch11", "test_ch12", "test_ch13", "test_ch14"]
METRIC = "1280x720 frames/sec, process_image with Line tracking (value: frames resident in HBM; e2e: pinned host in -> pinned host out)"
UNIT = "frames/s"
WORKLOAD = ("project_video 1260 frames 1280x720 pre-decoded per GPU (synthetic sequence from the 8 project stills; frames of rank r = "
            "generator seed r), Line tracking, frame_num=15")
FRAME_BYTES = 720 * 1280 * 3


def load_stills(names=STILLS):
    from PIL import Image
    out = []
    for n in names:
        with Image.open(os.path.join(ROOT, "tests", "golden", "stills", n + ".jpg")) as im:
            out.append(np.asarray(im.convert("RGB")).copy())
    return out


def synth_params(n, seed):
    """per-frame (x shift, gain) of the SURVEY 8(d) generator -- the same random stream as oracle.lane_port.synth_sequence
    (kept separate: the bench's GPU arm never imports oracle/)"""
    rng = np.random.default_rng(seed)
    dx, dxs, gains = 0, np.empty(n, np.int64), np.empty(n, np.float64)
    for i in range(n):
        dx = int(np.clip(dx + rng.integers(-2, 3), -8, 8))
        dxs[i], gains[i] = dx, float(rng.uniform(0.9, 1.1))
    return dxs, gains


def synth_frames(n, seed, lo=0, hi=None, names=STILLS, out=None):
    """frames [lo, hi) of the n-frame sequence with this seed"""
    hi = n if hi is None else hi
    stills = load_stills(names)
    dxs, gains = synth_params(n, seed)
    if out is None:
        out = np.empty((hi - lo, 720, 1280, 3), np.uint8)
    lut_in = np.arange(256, dtype=np.float32)
    for i in range(lo, hi):
        base = stills[i % len(stills)]
        cols = np.clip(np.arange(1280) - int(dxs[i]), 0, 1279)
        lut = np.clip(np.rint(lut_in * np.float32(gains[i])), 0, 255).astype(np.uint8)
        out[i - lo] = lut[base[:, cols]]
    return out


class ClockSampler:
    """SM clock and throttle reasons of one GPU sampled DURING the timed region (B200_PROFILING.md's clocks line).  The region is
    a few tens of milliseconds, far below nvidia-smi's start-up time, so a thread polls NVML directly (about every millisecond);
    nvidia-smi -lms is the fallback when NVML cannot be loaded."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu, uuid=None):
        self.gpu, self.uuid = gpu, uuid
        self.sm, self.max_mhz, self.reasons, self.lines = [], None, set(), []
        self.handle = self.nv = self.proc = self.thread = None
        self.stop = threading.Event()
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.handle = nv.nvmlDeviceGetHandleByUUID(uuid) if uuid else nv.nvmlDeviceGetHandleByIndex(gpu)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
            self.nv = nv
        except Exception:
            self.handle = None

    def _poll(self):
        nv, h = self.nv, self.handle
        while not self.stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for bit, name in self.REASONS:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.001)

    def __enter__(self):
        if self.handle is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.thread.start()
            time.sleep(0.5)                                 # let nvidia-smi come up before the timed region starts
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        if self.thread:
            self.thread.join(timeout=2)

    def summary(self):
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                self.sm.append(float(p[1]))
                self.max_mhz = max(self.max_mhz or 0, float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": "nvml" if self.handle is not None else "nvidia-smi"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ CPU arm
_WORKER = {}


def _cpu_init(frames_per_proc, counter):
    """per-process set-up, outside the timed region: this worker's slice of the FIRST frames of the bench sequence (seed 0 =
    rank 0's video), the oracle"""
    import cv2
    cv2.setNumThreads(1)
    sys.path.insert(0, ROOT)
    from oracle import lane_port as O
    with counter.get_lock():
        idx = counter.value
        counter.value += 1
    lo = idx * frames_per_proc
    _WORKER["frames"] = synth_frames(N_FRAMES, 0, lo, lo + frames_per_proc)
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
    """whole-host aggregate of the reference CPU path: one single-threaded process per core, process p on frames
    [8p, 8p + 8) of the bench sequence held in RAM; a step = every process runs process_image over its frames once."""
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    procs = max(1, min(cores, 64, N_FRAMES // frames_per_proc))
    ctx = mp.get_context("spawn")
    counter = ctx.Value("i", 0)
    times = []
    with ctx.Pool(procs, initializer=_cpu_init, initargs=(frames_per_proc, counter)) as pool:
        pool.map(_cpu_worker, range(procs), chunksize=1)          # touch every worker once (imports, first-call costs)
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, range(procs), chunksize=1)
            if s >= warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": procs * frames_per_proc * steps / total, "ms_per_step": 1e3 * total / steps, "cores": procs,
            "sample": "the first %d frames of the workload (seed 0), %d single-threaded processes x %d consecutive frames each, fresh "
                      "trackers per process (oracle port: same cv2/numpy calls as Project.py)" % (procs * frames_per_proc, procs, frames_per_proc),
            "frames_per_step": procs * frames_per_proc}


def cpu_img2binary_ms(n=4):
    """the reference's helper.img2binary (oracle port) on the host, ms per frame, one process"""
    from oracle import lane_port as O
    frames = synth_frames(N_FRAMES, 0, 0, n)
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
    frames = synth_frames(N_FRAMES, 0, 0, n)
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


def cpu_single_frame_ms(frame, reps=5):
    """config 1 on the host: the oracle port's process_image on test1.jpg with fresh trackers, ms per call"""
    import warnings
    from oracle import lane_port as O
    orc = O.LaneOracle("project")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        orc.process(frame)
        t0 = time.perf_counter()
        for _ in range(reps):
            orc.reset()
            orc.process(frame)
    return 1e3 * (time.perf_counter() - t0) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--frames", type=int, default=N_FRAMES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs (c1, c3, c4)")
    ap.add_argument("--shard-mode", default=None, help="lookback | records | serial (default: LD_SHARD_MODE or lookback)")
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

    # This script's stdout is ONE JSON line, but libraries write there too (NCCL prints its version banner on stdout at the
    # VERSION and WARN debug levels): keep the real stdout aside and point file descriptor 1 at stderr until the line is printed.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
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
    dev = torch.device("cuda", local)
    try:
        gpu_uuid = "GPU-" + str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        gpu_uuid = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(flag):
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def device_rate(step_fn, finish_fn, steps, warm, frames_per_step):
        """frames/s over `steps` steps: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks"""
        for _ in range(warm):
            step_fn()
        finish_fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            step_fn()
        finish_fn()
        e1.record()
        barrier()
        ms = reduce_max(e0.elapsed_time(e1))
        return world * frames_per_step * steps / (ms / 1e3), ms / steps

    # ---- this rank's range of the video, and the look-back in front of it (the last HALO frames of the predecessor's range)
    def make_inputs(names, seed_base):
        pin = torch.empty((HALO + n, 720, 1280, 3), dtype=torch.uint8, pin_memory=True)
        h = HALO if rank > 0 else 0
        if h:
            synth_frames(n, seed_base + rank - 1, n - HALO, n, names, out=pin.numpy()[:HALO])
        synth_frames(n, seed_base + rank, 0, n, names, out=pin.numpy()[HALO:])
        return pin, h

    pin_all, h = make_inputs(STILLS, 0)
    pin_in, pin_look = pin_all[HALO:], (pin_all[HALO - h:HALO] if h else None)
    pin_out = torch.empty((n, 720, 1280, 3), dtype=torch.uint8, pin_memory=True)
    dev_all = pin_all[HALO - h:].to(dev)
    frames, lookback = dev_all[h:], (dev_all[:h] if h else None)
    out = torch.empty_like(frames)
    fits = torch.empty((n, pkg._abi.FIT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    eng = pkg.LaneEngine("project", device=local, max_frames=n + HALO)

    def fetch_from(pin_src_names, seed_base):
        def fetch(k):                                        # repair path of the look-back protocol: more frames of the predecessor
            k = min(k, n)
            return torch.from_numpy(synth_frames(n, seed_base + rank - 1, n - k, n, pin_src_names)).to(dev), False
        return fetch

    mode = a.shard_mode or os.environ.get("LD_SHARD_MODE", "lookback")
    run = sharding.ShardedRun(eng, rank, world, mode=mode, halo=HALO, fetch_lookback=fetch_from(STILLS, 0))

    # ---- the headline: device-resident frames/s
    for _ in range(a.warmup):
        run.step(frames, out=out, fits=fits, lookback=lookback)
    run.finish()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local, gpu_uuid) as clocks:
        barrier()
        ev0.record()
        for _ in range(a.steps):
            run.step(frames, out=out, fits=fits, lookback=lookback)
        run.finish()
        ev1.record()
        barrier()
    eng.check()
    ms = reduce_max(ev0.elapsed_time(ev1))
    value = world * n * a.steps / (ms / 1e3)
    repairs = run.repairs
    # kernels of one step (profiles/r02_launches.csv): state reset; plain tiles; shared pass (on these plans one launch: no share
    # tile is sampled by the mask without being touched by the inverse warp), k_mask_b, k_search; 5 fit kernels; k_raster_prep,
    # k_raster; k_blend_px, k_text_px; k_near_integer
    # (rank 0 prints: its range starts the video, so it has no look-back; the other ranks of a 'lookback' run add k_halo_start)
    launches_per_step = (1 + 1 + 3 + 5 + 2 + 2 + 1) if (world == 1 or mode != "records") else 16

    # ---- sharded results against the serial protocol and against each other, then the other protocols' rates
    shard = None
    if world > 1:
        out2, fits2 = torch.empty_like(out), torch.empty_like(fits)
        shard = {"default": mode, "protocols": {mode: {"frames_per_s": value, "ms_per_step": ms / a.steps, "repairs": repairs}}}
        same_all = True
        for m2 in sharding.MODES:
            if m2 == mode:
                continue
            r2 = sharding.ShardedRun(eng, rank, world, mode=m2, halo=HALO, fetch_lookback=fetch_from(STILLS, 0))
            step2 = lambda: r2.step(frames, out=out2, fits=fits2, lookback=lookback)
            fps2, ms2 = device_rate(step2, r2.finish, max(1, min(a.steps, 3)), 2, n)
            eng.check()
            same = all_true(bool((out2 == out).all().item()) and bool((fits2 == fits).all().item()))
            same_all = same_all and same
            shard["protocols"][m2] = {"frames_per_s": fps2, "ms_per_step": ms2, "repairs": r2.repairs, "identical_to_default": same}
        shard["shard_matches_serial"] = same_all
        # config 5 = the upper bound of any sharding protocol: independent streams, one per GPU, no look-back, no communication
        def indep():
            eng.reset()
            eng.process_dev(frames, out=out2, fits=fits2)
        fps5, ms5 = device_rate(indep, lambda: None, max(1, min(a.steps, 3)), 1, n)
        shard["independent_streams"] = {"frames_per_s": fps5, "ms_per_step": ms5,
                                        "note": "BASELINE config 5: one independent 1260-frame stream per GPU, zero communication"}
        del out2, fits2

    # ---- end to end through the C ABI host call (pinned host -> pinned host); at N > 1 the look-back travels with the range
    def e2e_call():
        if pin_look is None:                                 # the range starts the video: fresh trackers
            eng.reset()
            eng.process_pinned(pin_in, pin_out)
            return False
        return eng.process_range_pinned(pin_look, pin_in, pin_out)
    e2e_bad = e2e_call()                                     # warm-up
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(a.steps, 3))
    for _ in range(e2e_steps):
        e2e_bad = e2e_call() or e2e_bad
    torch.cuda.synchronize()
    dt = reduce_max(time.perf_counter() - t0)
    e2e = world * n * e2e_steps / dt
    same_e2e = all_true(bool((pin_out.to(dev) == out).all().item()) and not e2e_bad)
    # the platform's ceiling for this traffic pattern, same run: plain pinned copies of the same buffers both ways at once, no kernels
    scratch = torch.empty_like(out)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    pcie_fps = 0.0
    for _ in range(3):                                       # best of three rounds (the first one also warms the copy engines up)
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            scratch.copy_(pin_in, non_blocking=True)
        with torch.cuda.stream(s2):
            pin_out.copy_(frames, non_blocking=True)
        torch.cuda.synchronize()
        pcie_fps = max(pcie_fps, world * n / reduce_max(time.perf_counter() - t0))
    del scratch

    # ---- config 3: the harder pipeline (frame_num = 5), sharded the same way
    cfg = {}
    if not a.no_configs:
        pin3, h3 = make_inputs(HARDER_STILLS, 100)
        d3 = pin3[HALO - h3:].to(dev)
        del pin3
        eng3 = pkg.LaneEngine("harder", device=local, max_frames=n + HALO)
        run3 = sharding.ShardedRun(eng3, rank, world, mode=mode, halo=HALO, fetch_lookback=fetch_from(HARDER_STILLS, 100))
        fits3 = torch.empty_like(fits)
        step3 = lambda: run3.step(d3[h3:], out=out, fits=fits3, lookback=d3[:h3] if h3 else None)
        fps3, ms3 = device_rate(step3, run3.finish, max(1, min(a.steps, 3)), 2, n)
        eng3.check()
        cfg["c3_harder_pipeline"] = {"frames_per_s": fps3, "ms_per_step": ms3, "n_gpus": world, "sharding": mode if world > 1 else None,
                                     "repairs": run3.repairs,
                                     "workload": "Project_harder_challenge.py path, frame_num=5, 1260 synthetic frames per GPU from test_ch0..14 "
                                                 "(two stills with an empty lane: the previous-points fallback runs every 16 frames)"}
        eng3.close()
        del d3, fits3

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu": n, "l2": "inputs (3.5 GB/step) far larger than the 126 MB L2; no flush needed",
                       "parallelism": ("frame-range sharding, protocol '%s' (sharding.py)" % mode) if world > 1 else "1 GPU"},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": (n + h) * FRAME_BYTES, "d2h_bytes_per_step": n * FRAME_BYTES,
                    "matches_device_path": same_e2e, "pcie_ceiling_fps": pcie_fps,
                    "pcie_ceiling_note": "same run: plain pinned copies of the same buffers both ways at once, no kernels",
                    "api": "ld_process_batch_host" if world == 1 else "ld_process_range_host (range + 30-frame look-back per rank)"},
            "gpu_launches": launches_per_step * a.steps}
    if shard:
        line["sharding"] = shard
        line["shard_matches_serial"] = shard["shard_matches_serial"]

    if world == 1:
        line.update(single_gpu_extras(a, eng, pkg, torch, frames, out, fits, ms, n, cfg))
    line["configs"] = cfg
    if rank == 0:
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


def single_gpu_extras(a, eng, pkg, torch, frames, out, fits, ms, n, cfg):
    """per-kernel timings and rooflines (CUDA events on the launching stream, each kernel alone over the same n resident
    frames), the CPU baseline, and BASELINE configs 1 and 4.  roofline = the launch that dominates the step; mask_roofline =
    the fused mask kernel BASELINE.json's metric names.  Algorithmic bytes per frame: SURVEY.md 8(d) / DESIGN.md 4."""
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
    MK = "ld_mask (k_frame_pass_mf<4> mask-only pass + k_mask_b)"
    time_stage(MK, lambda: L.ld_mask(h, frames.data_ptr(), n, mask_t.data_ptr(), st))
    time_stage("k_frame_pass_mf<4> (mask-only pass)", lambda: L.ld_stage(h, 8, frames.data_ptr(), n, None, st))
    time_stage("k_mask (ld_mask_fused: the one-kernel route)", lambda: L.ld_mask_fused(h, frames.data_ptr(), n, mask_t.data_ptr(), st))
    rec_t = eng.search_dev(mask_t)
    time_stage("k_search", lambda: L.ld_search(h, mask_t.data_ptr(), n, rec_t.data_ptr(), st))
    eng.reset()
    time_stage("k_fit*", lambda: L.ld_fit(h, rec_t.data_ptr(), n, fits.data_ptr(), st))
    time_stage("k_raster", lambda: L.ld_raster(h, fits.data_ptr(), n, st))
    # the launches the step is really made of (ld_stage): shared pass -> k_mask_b -> ... plain tiles, k_blend_px, k_text_px
    SH, PL, BL, TX = "k_frame_pass_mf<3> + <5> (shared pass)", "k_frame_pass_mf<0> (plain tiles)", "k_blend_px", "k_text_px (putText pixels)"
    time_stage(SH, lambda: L.ld_stage(h, 3, frames.data_ptr(), n, out.data_ptr(), st))   # with the output frames, as the step runs it
    time_stage("k_mask_b", lambda: L.ld_stage(h, 7, None, n, None, st))
    time_stage(PL, lambda: L.ld_stage(h, 5, frames.data_ptr(), n, out.data_ptr(), st))
    time_stage(BL, lambda: L.ld_stage(h, 6, None, n, out.data_ptr(), st))
    time_stage(TX, lambda: L.ld_stage(h, 4, frames.data_ptr(), n, out.data_ptr(), st))
    n_share, n_plain, n_text = L.ld_stage_tiles(h, 3), L.ld_stage_tiles(h, 5), L.ld_stage_tiles(h, 4)
    blend_px = 128 * 16 * L.ld_stage_tiles(h, 9)           # pixels of the share tiles the inverse warp can touch (und_px -> k_blend_px)
    final_px = 128 * 16 * L.ld_stage_tiles(h, 12)          # share tiles it cannot touch but the mask samples: shared pass -> output frame
    # (the share tiles neither of them needs are plain tiles of this step: counted in n_plain)
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
    tile_px = 128 * 16
    bytes_per_frame = {MK: alg, "k_mask (ld_mask_fused: the one-kernel route)": alg, "k_frame_pass_mf<4> (mask-only pass)": alg - 115200 + 43520,
                       "k_search": 115200 + 560, "k_raster": 722 * 82 * 8,
                       "k_frame_pass_mf<1> (whole frame: ld_frame_pass / ld_overlay)": 2 * FRAME_BYTES,
                       SH: blend_px * (3 + 4) + final_px * (3 + 3) + 43520,   # source pixels read; u32 pixels / RGB + decision bits written
                       "k_mask_b": 43520 + 115200, PL: n_plain * tile_px * 6,
                       BL: blend_px * (4 + 3) + 722 * 82 * 8,            # u32 pixels + canvas read, RGB written
                       "k_hist": 115200 + 20480, "k_frame_pass_mf<0> (cv2.undistort alone)": 2 * FRAME_BYTES,
                       # RGB read twice (gray, S channel), gray written once and read twice, binary written once
                       "img2binary (k_gray + k_sobel_max + k_grad_binary)": 2 * FRAME_BYTES + 4 * 720 * 1280}
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
    mask_roofline = roof(MK, alg, "ld_mask_dram_bytes_per_frame")
    mask_roofline["note"] = ("the mask stage ld_mask (north-star kernel 1: undistort + warp + colour threshold -> packed mask), both of its "
                             "launches timed together; algorithmic bytes = the source rows the warp samples + the packed mask; inside "
                             "ld_process_batch the mask comes from the shared pass + k_mask_b")
    step_bytes = 6650880                                    # end-to-end device-resident algorithmic bytes per frame
    step_roofline = {"bound": "hbm", "achieved": step_bytes * n / (ms / a.steps * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "algorithmic_bytes_per_frame": step_bytes}
    step_roofline["frac"] = step_roofline["achieved"] / peak
    extras = {"roofline": roofline, "mask_roofline": mask_roofline, "step_roofline": step_roofline, "kernels": kernels}

    cpu = None
    if not a.no_cpu_baseline:
        cpu = cpu_baseline_inline()
        kernels["img2binary (k_gray + k_sobel_max + k_grad_binary)"]["cpu_port_ms_per_frame"] = cpu_img2binary_ms()
    extras["cpu_baseline"] = cpu
    if a.no_configs:
        return extras

    # ---- config 1: single-frame process_image latency through the drop-in module (what moviepy's fl_image calls)
    proj = importlib.import_module(PKG + ".Project")
    frame1 = load_stills(["test1"])[0]
    proj.reset()
    for _ in range(10):
        proj.process_image(frame1)
    lat = []
    for _ in range(200):
        t0 = time.perf_counter()
        proj.process_image(frame1)
        lat.append(1e3 * (time.perf_counter() - t0))
    lat = np.sort(np.array(lat))
    cfg["c1_single_frame"] = {"ms_p50": float(lat[len(lat) // 2]), "ms_p99": float(lat[int(len(lat) * 0.99) - 1]), "ms_min": float(lat[0]),
                              "calls": len(lat), "api": "Project.process_image(frame) -> frame (numpy in, numpy out, trackers carried)",
                              "cpu_port_ms": None if a.no_cpu_baseline else cpu_single_frame_ms(frame1)}

    # ---- config 4: 3840x2160 frames through the fused mask kernel + the band histogram kernel
    P = importlib.import_module(PKG + ".plan")
    k4, W4, H4, n4 = 3, 3840, 2160, 48
    plan4 = P.build_plan(P.scaled_variant(P.PROJECT, k4), P.scaled_calibration(P.CAL_MTX, k4), P.CAL_DIST, size=(W4, H4))
    eng4 = pkg.LaneEngine(plan=plan4, max_frames=n4)
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    f4 = torch.randint(0, 256, (n4, H4, W4, 3), dtype=torch.uint8, device="cuda", generator=g)
    up = torch.from_numpy(np.repeat(np.repeat(frame1, k4, axis=0), k4, axis=1)).cuda()
    f4[::2] = up                                            # half noise (worst case for the colour decision), half upscaled road scene
    m4 = eng4.mask_dev(f4)
    h4 = eng4.hist_dev(m4)
    L4, c4, st4 = eng4.lib, eng4._ctx_h, eng4._stream()

    def both():
        L4.ld_mask(c4, f4.data_ptr(), n4, m4.data_ptr(), st4)
        L4.ld_hist(c4, m4.data_ptr(), n4, h4.data_ptr(), st4)
    stage.clear()
    time_stage("c4", both, reps=10)
    alg4 = plan4.mask_algorithmic_bytes + W4 * 8 * 2        # + the eight band histograms (u16 per column)
    fps4 = n4 / (stage["c4"] * 1e-3)
    cfg["c4_4k_mask_hist"] = {"frames_per_s": fps4, "gbs": alg4 * fps4 / 1e9, "frac_of_peak": alg4 * fps4 / 1e9 / peak,
                              "algorithmic_bytes_per_frame": alg4, "frames": n4,
                              "workload": "3840x2160 RGB batch (half uniform noise, half x3 upscaled test1), k_mask + k_hist"}
    eng4.close()
    return extras


if __name__ == "__main__":
    main()
