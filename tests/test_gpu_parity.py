"""GPU parity: the CUDA path (through the C ABI) against the oracle and the golden vectors produced by the
unmodified reference.  Bars (BASELINE.json north_star): masks / histograms / moments bit-exact, fit
coefficients <= 1e-9 relative, curvature / offset <= 1e-6 relative, overlay within +-1 LSB."""
import hashlib
import warnings

import numpy as np
import pytest

from conftest import unhex, load_pkg
from oracle import lane_port as O

pytestmark = pytest.mark.gpu

PROJECT_STILLS = ["straight_lines1", "straight_lines2"] + ["test%d" % i for i in range(1, 7)]
ALL = PROJECT_STILLS + ["test_ch%d" % i for i in range(15)]
FIT_RTOL = 1e-9
POS_RTOL = 1e-6


@pytest.fixture(scope="module")
def engines():
    pkg = load_pkg()
    made = {}

    def get(variant):
        if variant not in made:
            made[variant] = pkg.LaneEngine(variant, max_frames=48)
        return made[variant]
    yield get
    for e in made.values():
        e.close()


def unpack(mask_t):
    a = mask_t.cpu().numpy().view(np.uint32)
    return np.unpackbits(a.view(np.uint8), axis=-1, bitorder="little").reshape(a.shape[0], a.shape[1], -1)


def assert_fit_close(got, want, what):
    got, want = np.asarray(got, float), np.asarray(want, float)
    err = np.abs(got - want) / np.abs(want)
    assert (err <= FIT_RTOL).all(), "%s: rel err %s (got %s want %s)" % (what, err, got, want)


@pytest.mark.parametrize("variant", ["project", "harder"])
def test_mask_hist_search_stills(variant, engines, stills, golden):
    eng = engines(variant)
    orc = O.LaneOracle(variant)
    frames = np.stack([stills(n) for n in ALL])
    dev = eng.to_device(frames)
    mask_t = eng.mask_dev(dev)
    masks = unpack(mask_t)
    hist = eng.hist_dev(mask_t).cpu().numpy().astype(np.int64)
    rec = eng.records_to_numpy(eng.search_dev(mask_t))
    for i, name in enumerate(ALL):
        g = golden["stills"][name][variant]
        want = O.binary_birdseye(O.undistort(frames[i], orc.mtx, orc.dist), orc.v)
        assert (masks[i] == want).all(), "%s mask differs in %d px" % (name, (masks[i] != want).sum())
        assert int(masks[i].sum()) == g["mask_popcount"]
        assert hashlib.sha1(np.packbits(masks[i].astype(bool), axis=1, bitorder="little").tobytes()).hexdigest() == g["mask_sha1"]
        assert (hist[i] == O.band_histograms(want)).all(), name
        for lane in (0, 1):
            x, y, wins = O.window_search(want, bool(lane), orc.v)
            r = rec[i, lane]
            assert [int(b) for b in r["win_base"][:8]] == [w[1] for w in wins], (name, lane)
            assert [int(b) for b in r["win_n"][:8]] == [w[2] for w in wins], (name, lane)
            assert [bool(b) for b in r["win_kept"][:8]] == [w[3] for w in wins], (name, lane)
            if g["lanes"][lane] is None:
                assert not r["ok"]
            else:
                assert r["ok"]
                assert [str(int(m)) for m in r["mom"]] == g["lanes"][lane]["moments"], (name, lane)
                rows = np.unique(y)
                got_rows = np.nonzero(np.unpackbits(r["rowmap"].view(np.uint8), bitorder="little"))[0]
                assert (rows == got_rows).all()


def test_mask_noise_and_extremes(engines):
    """uniform noise exercises all 1024 bilinear fractions and the whole colour table; flat frames the borders"""
    eng = engines("project")
    orc = O.LaneOracle("project")
    rng = np.random.default_rng(7)
    frames = [rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8) for _ in range(3)]
    frames += [np.zeros((720, 1280, 3), np.uint8), np.full((720, 1280, 3), 255, np.uint8)]
    yellow = np.zeros((720, 1280, 3), np.uint8)
    yellow[..., 0], yellow[..., 1] = 255, 220
    frames.append(yellow)
    frames = np.stack(frames)
    dev = eng.to_device(frames)
    masks = unpack(eng.mask_dev(dev))
    for i in range(len(frames)):
        want = O.binary_birdseye(O.undistort(frames[i], orc.mtx, orc.dist), orc.v)
        assert (masks[i] == want).all(), "frame %d differs in %d px" % (i, (masks[i] != want).sum())
    assert (unpack(eng.mask_dev(dev, fused=True)) == masks).all()         # the one-kernel route (k_mask)


def test_colour_decision_all_colours(engines):
    """all 2^24 colours through the stand-alone colour stage (helper.luv_lab_filter, no warp)"""
    import cv2
    eng = engines("project")
    r, g, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8),
                          indexing="ij")
    img = np.stack([r, g, b], -1).reshape(4096, 4096, 3)
    want = O.colour_mask(img, O.PROJECT.l_thresh, O.PROJECT.b_thresh).reshape(-1)
    flat = img.reshape(-1, 3)
    per = 720 * 1280
    n = (flat.shape[0] + per - 1) // per
    padded = np.zeros((n * per, 3), np.uint8)
    padded[:flat.shape[0]] = flat
    got = []
    for a in range(0, n, 6):
        chunk = padded[a * per:(a + 6) * per].reshape(-1, 720, 1280, 3)
        got.append(eng.colour_mask_dev(eng.to_device(chunk)).cpu().numpy().reshape(-1))
    got = np.concatenate(got)[:flat.shape[0]]
    assert (got == want).all(), "%d colours differ" % (got != want).sum()
    assert int(got.sum()) == 7267295


@pytest.mark.parametrize("variant", ["project", "harder"])
def test_stills_fit_position_overlay(variant, engines, stills, golden):
    eng = engines(variant)
    orc = O.LaneOracle(variant)
    for name in ALL:
        g = golden["stills"][name][variant]
        eng.reset()
        orc.reset()
        frame = stills(name)
        dev = eng.to_device(frame[None])
        if g["status"] != "ok":
            eng.process_dev(dev)
            with pytest.raises(TypeError, match="expected 1D vector for x"):
                eng.check()
            continue
        out_t, fit_t = eng.process_dev(dev)
        eng.check()
        fit = eng.fits_to_numpy(fit_t)[0]
        assert_fit_close(fit["fit"][0], [unhex(c) for c in g["left_fit"]], name + " left")
        assert_fit_close(fit["fit"][1], [unhex(c) for c in g["right_fit"]], name + " right")
        assert abs(fit["offset"] - unhex(g["offset"])) <= POS_RTOL * abs(unhex(g["offset"]))
        assert abs(fit["mean_curv"] - unhex(g["mean_curv"])) <= POS_RTOL * abs(unhex(g["mean_curv"]))
        assert [int(v) for v in fit["n_points"]] == g["n_fit_points"]
        assert int((np.asarray(fit["rank1"]) < 3).sum() + (np.asarray(fit["rank2"]) < 3).sum()) == g["rank_warnings"]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = orc.process(frame).out
        got = out_t[0].cpu().numpy()
        diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
        assert diff.max() <= 1, "%s: overlay differs by up to %d in %d px" % (name, diff.max(), (diff > 1).any(-1).sum())


@pytest.mark.parametrize("key,variant", [("project_seed0_40", "project"), ("harder_seed3_32", "harder"),
                                         ("project_on_ch_seed3_32", "project")])
def test_sequences_tracking(key, variant, engines, stills, golden):
    """Line tracking across frames (deque medians, empty-lane fallback) in ONE batched call, against the
    reference's frame-by-frame results; then the same sequence split into uneven chunks must give identical bits."""
    g = golden["sequences"][key]
    n = len(g["records"])
    frames = np.stack(list(O.synth_sequence([stills(s) for s in g["stills"]], n, seed=g["seed"])))
    eng = engines(variant)
    eng.reset()
    dev = eng.to_device(frames)
    out_t, fit_t = eng.process_dev(dev)
    eng.check()
    fits = eng.fits_to_numpy(fit_t)
    orc = O.LaneOracle(variant)
    out = out_t.cpu().numpy()
    for i, rec in enumerate(g["records"]):
        assert_fit_close(fits[i]["fit"][0], [unhex(c) for c in rec["left_fit"]], "%s[%d] left" % (key, i))
        assert_fit_close(fits[i]["fit"][1], [unhex(c) for c in rec["right_fit"]], "%s[%d] right" % (key, i))
        assert [bool(v) for v in fits[i]["detected"]] == rec["detected"], i
        assert [int(v) for v in fits[i]["n_points"]] == rec["n_fit_points"], i
        assert abs(fits[i]["offset"] - unhex(rec["offset"])) <= POS_RTOL * abs(unhex(rec["offset"]))
        assert abs(fits[i]["mean_curv"] - unhex(rec["mean_curv"])) <= POS_RTOL * abs(unhex(rec["mean_curv"]))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = orc.process(frames[i]).out
        diff = np.abs(out[i].astype(np.int16) - want.astype(np.int16))
        assert diff.max() <= 1, "%s[%d]: overlay differs by up to %d in %d px" % (key, i, diff.max(), (diff > 1).any(-1).sum())
    # chunked: 1 + 7 + rest, state carried in the context
    eng.reset()
    parts = [eng.process_dev(dev[a:b]) for a, b in ((0, 1), (1, 8), (8, n))]
    eng.check()
    fits2 = np.concatenate([eng.fits_to_numpy(p[1]) for p in parts])
    out2 = np.concatenate([p[0].cpu().numpy() for p in parts])
    assert fits2.tobytes() == fits.tobytes()
    assert (out2 == out).all()
    # host entry (pinned, pipelined copies) gives the same bytes as the device entry
    eng.reset()
    out3, fits3 = eng.process_host(frames)
    assert (out3 == out).all() and fits3.tobytes() == fits.tobytes()


def test_state_export_import_handoff(engines, stills, golden):
    """the tracker blob handed between GPUs at chunk boundaries reproduces the single-context run"""
    g = golden["sequences"]["project_seed0_40"]
    frames = np.stack(list(O.synth_sequence([stills(s) for s in g["stills"]], 40, seed=g["seed"])))
    eng = engines("project")
    dev = eng.to_device(frames)
    eng.reset()
    _, fit_all = eng.process_dev(dev)
    eng.check()
    want = eng.fits_to_numpy(fit_all)
    eng.reset()
    _, fa = eng.process_dev(dev[:17])
    blob = eng.export_state()
    eng.reset()
    eng.import_state(blob)
    _, fb = eng.process_dev(dev[17:])
    eng.check()
    got = np.concatenate([eng.fits_to_numpy(fa), eng.fits_to_numpy(fb)])
    assert got.tobytes() == want.tobytes()


def test_helper_surface_stages(engines, stills):
    import cv2
    eng = engines("project")
    orc = O.LaneOracle("project")
    frame = stills("test3")
    dev = eng.to_device(frame[None])
    und = O.undistort(frame, orc.mtx, orc.dist)
    assert (eng.undistort_dev(dev)[0].cpu().numpy() == und).all()
    assert (eng.warp_dev(eng.to_device(und[None]))[0].cpu().numpy() == O.birdseye(und, orc.v)).all()
    want = O.colour_mask(und, orc.v.l_thresh, orc.v.b_thresh)
    assert (eng.colour_mask_dev(eng.to_device(und[None]))[0].cpu().numpy() == want).all()
    m = O.binary_birdseye(und, orc.v)
    packed = eng.pack_dev(eng.to_device(m[None]))
    assert (eng.unpack_dev(packed)[0].cpu().numpy() == m).all()
    assert (packed.cpu().numpy().view(np.uint32)[0] == O.pack_mask(m)).all()


def test_prepass_postpass_equals_process(engines, stills, golden):
    """the split the multi-GPU path uses (mask+search | fit+overlay, state blob copied out after the fit stage)"""
    import torch
    g = golden["sequences"]["project_seed0_40"]
    frames = np.stack(list(O.synth_sequence([stills(s) for s in g["stills"]], 40, seed=g["seed"])))
    eng = engines("project")
    dev = eng.to_device(frames)
    eng.reset()
    out_a, fit_a = eng.process_dev(dev)
    eng.check()
    want_state = eng.export_state()
    eng.reset()
    blob = torch.zeros(eng.lib.ld_state_size(), dtype=torch.uint8, device=dev.device)
    eng.prepass_dev(dev)
    out_b, fit_b = eng.postpass_dev(dev, state_out=blob)
    eng.check()
    assert (out_a == out_b).all() and (fit_a == fit_b).all()
    assert bytes(blob.cpu().numpy().tobytes()) == want_state


def test_prepass_out_and_second_postpass_over_the_same_output(engines, stills, golden):
    """ld_prepass_out (the share tiles the inverse warp cannot touch are written by the prepass already) followed by the
    post-pass gives ld_process_batch's bytes; and a SECOND post-pass over the same output with other trackers -- what a sharded
    rank does when it repairs its boundary (skip_plain: the plain tiles are in place) -- leaves no pixel of the first one's text"""
    g = golden["sequences"]["project_seed0_40"]
    frames = np.stack(list(O.synth_sequence([stills(s) for s in g["stills"]], 40, seed=g["seed"])))
    eng = engines("project")
    dev = eng.to_device(frames)
    eng.reset()
    want, want_fit = eng.process_dev(dev[20:40])            # trackers start empty at frame 20: other medians, other text
    eng.check()
    eng.reset()
    full, _ = eng.process_dev(dev)
    eng.check()
    assert not (full[20:40] == want).all()                  # the two histories really print different numbers
    out = eng.new_like(dev[20:40])
    eng.reset()
    eng.process_dev(dev[:20])                               # trackers as after frame 19
    eng.prepass_dev(dev[20:40], out=out)
    out_b, fit_b = eng.postpass_dev(dev[20:40], out=out)
    eng.check()
    assert (out_b == full[20:40]).all()
    eng.reset()                                             # the "repair": empty trackers, same context records, same output buffer
    out[:, 74:139, 62:538] = 255                            # (whatever the first pass printed: the whole putText window white)
    out_c, fit_c = eng.postpass_dev(dev[20:40], out=out, skip_plain=True)
    eng.check()
    assert (fit_c == want_fit).all()
    assert (out_c == want).all()


def test_config4_4k_mask_and_histogram():
    """BASELINE.json configs[3]: 3840x2160 frames through the fused mask + histogram kernels (maps from the
    calibration scaled x3), against cv2 run directly at that size."""
    import cv2
    import importlib
    pkg = load_pkg()
    P = importlib.import_module("advanced-lane-detection_b200.plan")
    k, W, H = 3, 3840, 2160
    v = P.scaled_variant(P.PROJECT, k)
    mtx = P.scaled_calibration(P.CAL_MTX, k)
    plan = P.build_plan(v, mtx, P.CAL_DIST, size=(W, H))
    eng = pkg.LaneEngine(plan=plan, max_frames=2)
    rng = np.random.default_rng(1)
    noise = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    from conftest import read_still
    up = np.repeat(np.repeat(read_still("test1"), k, axis=0), k, axis=1)          # nearest-upscaled still
    frames = np.stack([noise, up])
    dev = eng.to_device(frames)
    mask_t = eng.mask_dev(dev)                                                     # tile route (3840 = 30 x 128)
    assert (eng.mask_dev(dev, fused=True) == mask_t).all()                         # the one-kernel route
    masks = unpack(mask_t)
    hist = eng.hist_dev(mask_t).cpu().numpy().astype(np.int64)
    M = cv2.getPerspectiveTransform(np.float32(v.src), np.float32(v.dst))
    for i in range(2):
        und = cv2.undistort(frames[i], mtx, P.CAL_DIST, None, mtx)
        warped = cv2.warpPerspective(und, M, (W, H), flags=cv2.INTER_NEAREST)
        want = O.colour_mask(warped, v.l_thresh, v.b_thresh)
        assert (masks[i] == want).all(), "4K frame %d differs in %d px" % (i, (masks[i] != want).sum())
        bh = H // 8
        for b in range(8):
            top = H - bh * (b + 1)
            assert (hist[i, b] == want[top:top + bh].sum(axis=0)).all()
    assert masks[1].sum() > 100000
    eng.close()


def test_config5_independent_streams(engines, stills, golden):
    """BASELINE.json configs[4]: independent video streams = independent contexts; interleaving their calls on one
    GPU must not mix tracker state (one context per stream is what maps one stream per GPU)."""
    pkg = load_pkg()
    g = golden["sequences"]["project_seed0_40"]
    a = np.stack(list(O.synth_sequence([stills(s) for s in g["stills"]], 20, seed=0)))
    b = np.stack(list(O.synth_sequence([stills(s) for s in g["stills"][::-1]], 20, seed=1)))
    e1, e2 = pkg.LaneEngine("project", max_frames=32), pkg.LaneEngine("project", max_frames=32)
    da, db = e1.to_device(a), e2.to_device(b)
    ref_a = e1.fits_to_numpy(e1.process_dev(da)[1]).copy()
    e1.reset()
    parts_a, parts_b = [], []
    for lo in range(0, 20, 5):
        parts_a.append(e1.fits_to_numpy(e1.process_dev(da[lo:lo + 5])[1]))
        parts_b.append(e2.fits_to_numpy(e2.process_dev(db[lo:lo + 5])[1]))
    e1.check(); e2.check()
    assert np.concatenate(parts_a).tobytes() == ref_a.tobytes()
    assert np.concatenate(parts_b).tobytes() != ref_a.tobytes()
    e1.close(); e2.close()


def test_shared_pass_equals_fused_kernels(engines, stills):
    """ld_process_batch takes the mask from the shared pass (k_frame_pass_mf<3> -> k_mask_b) and blends onto its pixels
    (k_blend_px); ld_mask (mask-only tile pass k_frame_pass_mf<4> -> k_mask_b), ld_mask_fused (the one-kernel k_mask) and
    ld_overlay (whole-frame k_frame_pass_mf<1>) must give the same bytes."""
    eng = engines("project")
    rng = np.random.default_rng(11)
    frames = np.stack([stills("test1"), stills("test5"), rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8),
                       np.full((720, 1280, 3), 255, np.uint8), stills("straight_lines2")])
    dev = eng.to_device(frames)
    eng.reset()
    out_t, fit_t = eng.process_dev(dev)
    try:
        eng.check()
    except TypeError:
        pass                                                    # the noise / white frames have no lanes; masks and pixels still compare
    ctx_masks = eng.context_masks(len(frames)).clone()
    assert (ctx_masks == eng.mask_dev(dev)).all()
    assert (ctx_masks == eng.mask_dev(dev, fused=True)).all()
    ok = eng.fits_to_numpy(fit_t)["status"] == 0
    assert ok.sum() >= 3
    assert (out_t[torch_idx(ok, out_t)] == eng.overlay_dev(dev, fit_t)[torch_idx(ok, out_t)]).all()


def torch_idx(mask, like):
    import torch
    return torch.as_tensor(np.nonzero(mask)[0], device=like.device)


def test_halo_rebuilds_tracker_state(engines, stills):
    """sharding without a hand-off: 30 look-back lane records in front of a range, fitted from empty trackers, give the
    same fits, frames and follow-up behaviour as processing the video from its start"""
    eng = engines("project")
    names = ["straight_lines1", "straight_lines2", "test1", "test2", "test3", "test4", "test5", "test6"]
    seq = np.stack(list(O.synth_sequence([stills(n) for n in names], 48, seed=4)))
    dev = eng.to_device(seq)
    eng.reset()
    out_full, fit_full = eng.process_dev(dev)
    eng.check()
    rec_full = eng.context_records(48).clone()
    fits_full = eng.fits_to_numpy(fit_full).copy()
    eng.reset()
    eng.process_dev(dev[40:48])                                 # leave unrelated trackers behind: the halo path must not use them
    eng.prepass_dev(dev[31:40])
    eng.halo_set(rec_full[1:31])
    out_a, fit_a = eng.postpass_dev(dev[31:40], use_halo=True)
    out_b, fit_b = eng.process_dev(dev[40:48])                  # continues from the rebuilt trackers
    eng.check()
    for got_t, lo, hi in ((fit_a, 31, 40), (fit_b, 40, 48)):
        got = eng.fits_to_numpy(got_t)
        for k in ("fit", "fit1", "fit2", "bottom", "top", "offset", "mean_curv"):
            assert (got[k] == fits_full[k][lo:hi]).all(), (k, lo)
    assert (out_a == out_full[31:40]).all() and (out_b == out_full[40:48]).all()
