"""Parity at the sizes bench.py times: the 1 M / 6 M, 1920x1080 frames of BASELINE.json's
configurations against the CPU oracle -- projected records / depth keys / visible set bit-exact, per-tile order, image
bars of DESIGN.md section 2 with the transmittance early-out on and off, and the 6 M fp16 target against the oracle's
round-after-every-blend half target (src/app.cpp:1010-1014; shader/splat_frag.glsl:18-42 for the fragment rule).

These are the scenes where the size-dependent machinery runs: multi-wave look-back, 14.5 M-instance tile passes, capacity
handling, two render targets with overlapped copies.  The oracle renders a 1 M frame in about a second on the GPU box.
"""
import os

import numpy as np
import pytest

import bench
from splatapult_b200 import api, scenes

pytestmark = pytest.mark.gpu

W, H = bench.WIDTH, bench.HEIGHT
VP, NF = [0, 0, W, H], [scenes.Z_NEAR, scenes.Z_FAR]
FIELDS = ("px", "py", "i00", "i01", "i10", "i11", "r", "g", "b", "alpha", "key")


@pytest.fixture(scope="module")
def scene_dir(tmp_path_factory):
    d = os.environ.get("SPLAT_SCENE_DIR") or str(tmp_path_factory.mktemp("baseline_scenes"))
    old = bench.SCENE_DIR
    bench.SCENE_DIR = d
    yield d
    bench.SCENE_DIR = old


def _oracle_frame(oracle, name, cam):
    aos, full = oracle.import_ply(bench.scene_path(name), True)
    proj = scenes.perspective(scenes.FOVY, W / H)
    want, rec, order, _ = oracle.render(aos, full, cam, proj, VP, NF, fp16=False, want_records=True)
    return want, rec, order


def _renderer(splat, name, fb_format, sort_mode=api.SORT_LSD):
    r = splat.SplatRenderer(fb_format=fb_format)
    r.InitFromPly(bench.scene_path(name), True, False)  # the bench's own load path (GPU PLY decode)
    r.SetOption(api.OPT_SORT_MODE, sort_mode)
    return r


def _frame(r, cam, early_out):
    proj = scenes.perspective(scenes.FOVY, W / H)
    r.SetOption(api.OPT_EARLY_OUT_T, early_out)
    r.Sort(cam, proj, VP, NF)
    r.Render(cam, proj, VP, NF)
    return r.ReadFramebuffer()


def _check_records(got, want):
    np.testing.assert_array_equal(got["visible"], want["visible"])
    np.testing.assert_array_equal(got["key"], want["key"])
    np.testing.assert_array_equal(got["alpha"], want["alpha"])
    vis = want["visible"] == 1
    for f in FIELDS:
        np.testing.assert_array_equal(got[f][vis].view(np.uint32), want[f][vis].view(np.uint32), err_msg=f)


def _check_tile_order(r, want_rec, want_order):
    """every tile's run is the oracle's back-to-front order restricted to that tile (ties included), ranges delimit the runs"""
    keys, vals, ranges = r.Tap(api.TAP_KEYS), r.Tap(api.TAP_VALS), r.Tap(api.TAP_TILE_RANGES)
    st = r.GetStats()
    assert not st.overflow and len(keys) == st.num_instances
    assert np.all(keys[1:] >= keys[:-1])
    np.testing.assert_array_equal((keys & np.uint64(0xFFFFFFFF)).astype(np.uint32), want_rec["key"][vals])
    rank = np.full(len(want_rec), -1, np.int64)
    rank[want_order] = np.arange(len(want_order))
    rv = rank[vals]
    assert np.all(rv >= 0)
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    same = tiles[1:] == tiles[:-1]
    assert np.all(rv[1:][same] > rv[:-1][same])
    b, e = ranges[:, 0].astype(np.int64), ranges[:, 1].astype(np.int64)
    nonempty = e > b
    assert int((e - b)[nonempty].sum()) == len(keys)
    assert np.all(tiles[b[nonempty]] == np.nonzero(nonempty)[0]) and np.all(tiles[e[nonempty] - 1] == np.nonzero(nonempty)[0])


def _image_stats(fb, want):
    d = np.abs(fb[..., :3] - want[..., :3])
    return float(d.max()), float(np.quantile(d.reshape(-1), 0.999)), float(d.mean())


@pytest.mark.parametrize("name", ["1m_deg0", "1m_deg3"])
def test_1m_frame_matches_the_oracle(splat, oracle, scene_dir, name):
    """BASELINE configs[1] / configs[2]: the frame bench.py times, bit-exact records and order, image bars on / off early-out"""
    cam = bench.camera_for_rank(name, 0, 1)
    want, wrec, worder = _oracle_frame(oracle, name, cam)
    r = _renderer(splat, name, api.FB_FP32)
    fb = _frame(r, cam, bench.EARLY_OUT_T)
    st = r.GetStats()
    assert st.num_gaussians == 1_000_000 and st.num_visible == int((wrec["visible"] == 1).sum()) and not st.overflow
    _check_records(r.Tap(api.TAP_RECORDS), wrec)
    _check_tile_order(r, wrec, worder)
    np.testing.assert_array_equal(fb[..., 3], 1.0)
    mx, p999, mean = _image_stats(fb, want)
    assert mx <= bench.BARS_FP32["max_abs"] and p999 <= bench.BARS_FP32["p999"] and mean <= bench.BARS_FP32["mean"], (mx, p999, mean)
    fb0 = _frame(r, cam, 0.0)  # the reference's full blend chain
    mx, p999, mean = _image_stats(fb0, want)
    assert mx <= bench.BARS_FP32_EXACT["max_abs"] and p999 <= bench.BARS_FP32_EXACT["p999"] and mean <= bench.BARS_FP32_EXACT["mean"], (mx, p999, mean)
    # the asynchronous path delivers the same pixels as the blocking read (two targets, copy stream)
    pin = [splat.host_alloc(W * H * 16), splat.host_alloc(W * H * 16)]
    proj = scenes.perspective(scenes.FOVY, W / H)
    r.SetOption(api.OPT_EARLY_OUT_T, bench.EARLY_OUT_T)
    for i in range(3):
        r.Sort(cam, proj, VP, NF)
        r.Render(cam, proj, VP, NF)
        r.ReadFramebufferAsync(pin[i & 1].ptr, W * H * 16)
    r.WaitCopies(0)
    for p in pin:
        np.testing.assert_array_equal(p.array(np.float32, (H, W, 4)), fb)


def test_1m_ring_view_matches_the_oracle(splat, oracle, scene_dir):
    """the view a rank > 0 renders at N > 1 (bench.camera_for_rank): nobody looked at those frames in round 1"""
    name = "1m_deg3"
    cam = bench.camera_for_rank(name, 3, 8)
    want, wrec, worder = _oracle_frame(oracle, name, cam)
    r = _renderer(splat, name, api.FB_FP32)
    fb = _frame(r, cam, bench.EARLY_OUT_T)
    _check_records(r.Tap(api.TAP_RECORDS), wrec)
    mx, p999, mean = _image_stats(fb, want)
    assert mx <= 5e-3 and p999 <= 1e-3 and mean <= 1e-4, (mx, p999, mean)


def test_6m_fp16_frame_matches_the_oracle(splat, oracle, scene_dir):
    """BASELINE configs[3]: 6 M gaussians, D = 14.5 M instances, half target.  The reference's RGBA16F target rounds after
    EVERY blend (src/app.cpp:1010-1014), the product once at the end: bar 2^-7 relative to max(1, |v|) against the oracle's
    round-every-blend image (about 100 layers per pixel here), and the fp32 bars' mean against the oracle's fp32 image."""
    name = "6m_deg3_fp16"
    cam = bench.camera_for_rank(name, 0, 1)
    want, wrec, worder = _oracle_frame(oracle, name, cam)
    want16 = oracle.composite(wrec, worder, W, H, fp16=True)
    r = _renderer(splat, name, api.FB_FP16)
    fb = _frame(r, cam, bench.EARLY_OUT_T)
    st = r.GetStats()
    assert st.num_gaussians == 6_000_000 and not st.overflow and st.num_instances > 10_000_000
    _check_records(r.Tap(api.TAP_RECORDS), wrec)
    _check_tile_order(r, wrec, worder)
    np.testing.assert_array_equal(fb[..., 3], 1.0)
    rel = np.abs(fb[..., :3] - want16[..., :3]) / np.maximum(1.0, np.abs(want16[..., :3]))
    assert rel.max() <= bench.BAR_FP16_REL, rel.max()
    d = np.abs(fb[..., :3] - want[..., :3])
    assert d.max() <= 2.0 ** -8 * max(1.0, float(np.abs(want[..., :3]).max())) + 5e-3 and d.mean() <= 1e-3, (d.max(), d.mean())


def test_config5_views_from_cameras_json(splat, oracle, scene_dir):
    """BASELINE configs[4] on one GPU: two of the eight cameras.json views of the 6 M cloud, read through CamerasConfig"""
    name = "6m_deg3_fp16"
    aos, full = oracle.import_ply(bench.scene_path(name), True)
    proj = scenes.perspective(scenes.FOVY, W / H)
    r = _renderer(splat, name, api.FB_FP16)
    for rank in (1, 6):
        cam, view, ncams = bench.config5_camera(name, rank, 8)
        assert ncams == 8 and view == rank
        want, wrec, worder, _ = oracle.render(aos, full, cam, proj, VP, NF, fp16=False, want_records=True)
        fb = _frame(r, cam, bench.EARLY_OUT_T)
        assert not r.GetStats().overflow
        _check_records(r.Tap(api.TAP_RECORDS), wrec)
        want16 = oracle.composite(wrec, worder, W, H, fp16=True)
        rel = np.abs(fb[..., :3] - want16[..., :3]) / np.maximum(1.0, np.abs(want16[..., :3]))
        assert rel.max() <= bench.BAR_FP16_REL, (rank, rel.max())
