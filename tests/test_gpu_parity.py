"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Bars (DESIGN.md "Parity"):
  * projected records, depth keys, visible set, sort order, tile lists: BIT-EXACT (integer/index work, and float
    work done with IEEE ops in the oracle's order).
  * framebuffer fp32: max |diff| <= 5e-3 (one 1/256 discard-threshold flip), 99.9 percentile <= 1e-3,
    mean <= 1e-4; with early-out disabled the same bars hold with mean <= 2e-5.
  * framebuffer fp16: max |diff| <= 2^-8 * max(1, |value|) against the oracle's fp32 image.
"""
import os

import numpy as np
import pytest

from splatapult_b200 import api, scenes

pytestmark = pytest.mark.gpu

NF = [scenes.Z_NEAR, scenes.Z_FAR]


SORT_MODE = api.SORT_BINNED


@pytest.fixture(params=[api.SORT_LSD, api.SORT_LSD_FLAT, api.SORT_BINNED], ids=["lsd", "lsd_flat", "binned"], autouse=True)
def sort_mode(request):
    """every test runs on both frame-sort paths: MSD-by-tile (default) and the device-wide LSD onesweep"""
    global SORT_MODE
    SORT_MODE = request.param
    return request.param


def _new(splat, **kw):
    r = splat.SplatRenderer(**kw)
    r.SetOption(api.OPT_SORT_MODE, SORT_MODE)
    return r


def _renderer(splat, aos, full, fb_format=0, early_out=None):
    gc = splat.GaussianCloud()
    gc.SetRawData(aos, full)
    r = _new(splat, fb_format=fb_format)
    assert r.Init(gc, False, False), getattr(r, "last_error", "")
    if early_out is not None:
        r.SetOption(api.OPT_EARLY_OUT_T, early_out)
    return r


def _frame(r, cam, proj, w, h):
    vp = [0, 0, w, h]
    r.Sort(cam, proj, vp, NF)
    r.Render(cam, proj, vp, NF)
    return r.ReadFramebuffer()


def _check_records(got, want):
    np.testing.assert_array_equal(got["visible"], want["visible"])
    vis = want["visible"] == 1
    for f in ("px", "py", "i00", "i01", "i10", "i11", "r", "g", "b", "alpha", "key"):
        np.testing.assert_array_equal(got[f][vis], want[f][vis], err_msg=f)
    # culled gaussians still carry their pre-sort key and alpha
    np.testing.assert_array_equal(got["key"], want["key"])
    np.testing.assert_array_equal(got["alpha"], want["alpha"])


def _check_tile_lists(r, want_rec, want_order, w, h):
    """sorted (tile|key) stream: every tile's run is the oracle's back-to-front order restricted to that tile."""
    keys, vals, ranges = r.Tap(api.TAP_KEYS), r.Tap(api.TAP_VALS), r.Tap(api.TAP_TILE_RANGES)
    st = r.GetStats()
    assert not st.overflow and len(keys) == st.num_instances
    assert np.all(keys[1:] >= keys[:-1])
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    np.testing.assert_array_equal((keys & np.uint64(0xFFFFFFFF)).astype(np.uint32), want_rec["key"][vals])
    rank = np.full(len(want_rec), -1, np.int64)
    rank[want_order] = np.arange(len(want_order))
    assert np.all(rank[vals] >= 0)  # only visible gaussians are instanced
    # within a tile, oracle rank strictly increases (stable ties included)
    same = tiles[1:] == tiles[:-1]
    assert np.all(rank[vals][1:][same] > rank[vals][:-1][same])
    # ranges delimit exactly the runs
    total = 0
    for t, (b, e) in enumerate(ranges):
        if e > b:
            assert np.all(tiles[b:e] == t)
            total += e - b
    assert total == len(keys)
    assert st.tiles_x == (w + st.tile_w - 1) // st.tile_w and st.tiles_y == (h + st.tile_h - 1) // st.tile_h
    return tiles, vals


def _coverage_is_conservative(want_rec, tiles, vals, w, h, tiles_x, tile_w=32, tile_h=16):
    """every pixel the oracle's fragment test accepts lies in a tile instanced for that gaussian."""
    inst = set(zip(vals.tolist(), tiles.tolist()))
    ys, xs = np.mgrid[0:h, 0:w]
    fx, fy = xs + 0.5, ys + 0.5
    for i in np.nonzero(want_rec["visible"])[0][:400]:
        s = want_rec[i]
        dx, dy = (fx - s["px"]).astype(np.float32), (fy - s["py"]).astype(np.float32)
        q = dx * (s["i00"] * dx + s["i10"] * dy) + dy * (s["i01"] * dx + s["i11"] * dy)
        hit = s["alpha"] * np.exp(-0.5 * q) > 1 / 256
        for t in np.unique((ys[hit] // tile_h) * tiles_x + xs[hit] // tile_w):
            assert (int(i), int(t)) in inst, (i, t)


def _check_image(fb, want, fp16=False, exact_mode=False):
    assert fb.shape == want.shape
    np.testing.assert_array_equal(fb[..., 3], 1.0)
    d = np.abs(fb[..., :3] - want[..., :3])
    if fp16:
        assert np.all(d <= 2.0 ** -8 * np.maximum(1.0, np.abs(want[..., :3])) + 5e-3 * (d > 2.0 ** -8)), d.max()
        assert d.mean() < 1e-3
        return
    assert d.max() <= 5e-3, d.max()
    assert np.quantile(d, 0.999) <= 1e-3
    assert d.mean() <= (2e-5 if exact_mode else 1e-4), d.mean()


def test_reference_fixture_config1(splat, oracle, golden_dir):
    """BASELINE config 1: data/test.ply, 640x480, the app's start-up camera."""
    g = np.load(os.path.join(golden_dir, "test_ply_640x480.npz"))
    gc = splat.GaussianCloud()
    assert gc.ImportPly(os.path.join(golden_dir, "reference_test.ply"))
    r = _new(splat)
    assert r.Init(gc, False, False)
    fb = _frame(r, g["camera"], g["proj"], 640, 480)
    _check_records(r.Tap(api.TAP_RECORDS), g["records"])
    tiles, vals = _check_tile_lists(r, g["records"], g["order"], 640, 480)
    _coverage_is_conservative(g["records"], tiles, vals, 640, 480, r.GetStats().tiles_x)
    _check_image(fb, g["fb"])
    st = r.GetStats()
    assert st.num_visible == 16 and st.sh_mode == 3 and st.kernel_launches >= 5
    # live oracle == committed golden
    fb_o = oracle.render(np.load(os.path.join(golden_dir, "test_ply_decode.npy")), True, g["camera"], g["proj"], [0, 0, 640, 480], NF)
    np.testing.assert_array_equal(fb_o, g["fb"])


def test_golden_synth_deg3(splat, golden_dir):
    g = np.load(os.path.join(golden_dir, "synth_2k_deg3_320x180.npz"))
    r = _renderer(splat, g["aos"], True)
    fb = _frame(r, g["camera"], g["proj"], 320, 180)
    _check_records(r.Tap(api.TAP_RECORDS), g["records"])
    tiles, vals = _check_tile_lists(r, g["records"], g["order"], 320, 180)
    _coverage_is_conservative(g["records"], tiles, vals, 320, 180, r.GetStats().tiles_x)
    _check_image(fb, g["fb"])


def test_debug_cloud(splat, oracle):
    """InitDebugCloud rendered from translate(0.5, 0.5, 3): stride 244 without full SH."""
    gc = splat.GaussianCloud()
    gc.InitDebugCloud()
    r = _new(splat)
    assert r.Init(gc, False, False)
    cam, proj = scenes.translate(0.5, 0.5, 3), scenes.perspective(scenes.FOVY, 640 / 480)
    fb = _frame(r, cam, proj, 640, 480)
    want, rec, order, _ = oracle.render(oracle.debug_cloud(), False, cam, proj, [0, 0, 640, 480], NF, want_records=True)
    _check_records(r.Tap(api.TAP_RECORDS), rec)
    _check_image(fb, want)
    assert r.GetStats().sh_mode == 0


@pytest.mark.parametrize("deg,n,w,h,cam_z", [(0, 30000, 640, 360, 10.0), (3, 30000, 640, 360, 10.0), (3, 8000, 333, 211, 6.0)])
def test_synthetic_scene(splat, oracle, scratch, deg, n, w, h, cam_z):
    """seeded clouds through the .ply load path; odd resolution exercises partial tiles."""
    path = str(scratch / f"scene_{deg}_{n}.ply")
    scenes.write_synthetic_ply(path, n, deg, seed=101 + deg)
    gc = splat.GaussianCloud()
    assert gc.ImportPly(path)
    r = _new(splat)
    assert r.Init(gc, False, False)
    cam, proj = scenes.translate(0, 0, cam_z), scenes.perspective(scenes.FOVY, w / h)
    fb = _frame(r, cam, proj, w, h)
    aos, full = oracle.import_ply(path, True)
    np.testing.assert_array_equal(gc.GetRawData(), aos)
    want, rec, order, _ = oracle.render(aos, full, cam, proj, [0, 0, w, h], NF, want_records=True)
    _check_records(r.Tap(api.TAP_RECORDS), rec)
    _check_tile_lists(r, rec, order, w, h)
    _check_image(fb, want)
    assert r.GetStats().num_visible == int(rec["visible"].sum())
    assert r.GetStats().sh_mode == (3 if deg == 3 else 0)
    # exact mode (no transmittance early-out): the full chain of the reference
    r.SetOption(api.OPT_EARLY_OUT_T, 0.0)
    _check_image(_frame(r, cam, proj, w, h), want, exact_mode=True)


def test_camera_inside_cloud_culls_and_huge_splats(splat, oracle, scratch):
    """camera in the middle of the cloud: near/guard culls bite, splats close to the eye cover hundreds of tiles."""
    a = scenes.synthetic_cloud_arrays(20000, 0, seed=33, half_extent=2.0)
    a["log_scale"][:200] += 2.5  # some very large gaussians
    path = scenes.write_ply_from_arrays(str(scratch / "inside.ply"), a)
    gc = splat.GaussianCloud()
    assert gc.ImportPly(path)
    r = _new(splat)
    assert r.Init(gc, False, False)
    cam = scenes.look_at_camera((0.3, 0.2, 0.1), (1.0, 0.4, -2.0))
    proj = scenes.perspective(scenes.FOVY, 512 / 384)
    fb = _frame(r, cam, proj, 512, 384)
    aos, full = oracle.import_ply(path, True)
    want, rec, order, _ = oracle.render(aos, full, cam, proj, [0, 0, 512, 384], NF, want_records=True)
    assert 0.05 < rec["visible"].mean() < 0.8
    _check_records(r.Tap(api.TAP_RECORDS), rec)
    _check_tile_lists(r, rec, order, 512, 384)
    _check_image(fb, want)


def test_capacity_overflow_regrows(splat, oracle, scratch):
    a = scenes.synthetic_cloud_arrays(3000, 0, seed=5, half_extent=1.0)
    a["log_scale"] += 2.0
    path = scenes.write_ply_from_arrays(str(scratch / "overflow.ply"), a)
    gc = splat.GaussianCloud()
    assert gc.ImportPly(path)
    r = _new(splat)
    r.SetOption(api.OPT_INSTANCES_PER_GAUSSIAN, 1.0)
    assert r.Init(gc, False, False)
    cam, proj = scenes.translate(0, 0, 4), scenes.perspective(scenes.FOVY, 640 / 480)
    r.Sort(cam, proj, [0, 0, 640, 480], NF)
    st = r.GetStats()
    assert st.overflow and st.num_instances > st.instance_capacity
    r.Render(cam, proj, [0, 0, 640, 480], NF)
    fb = r.ReadFramebuffer()  # grows the key buffers and redoes the frame
    st = r.GetStats()
    assert not st.overflow and st.instance_capacity >= st.num_instances
    aos, full = oracle.import_ply(path, True)
    _check_image(fb, oracle.render(aos, full, cam, proj, [0, 0, 640, 480], NF))


def test_fp16_target(splat, oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "synth_2k_deg3_320x180.npz"))
    r = _renderer(splat, g["aos"], True, fb_format=api.FB_FP16)
    fb = _frame(r, g["camera"], g["proj"], 320, 180)
    assert r.Framebuffer()[4] == api.FB_FP16 and r.Framebuffer()[3] == 320 * 8
    _check_image(fb, g["fb"], fp16=True)
    fb16 = oracle.composite(g["records"], g["order"], 320, 180, fp16=True)  # reference semantics: round every blend
    assert np.abs(fb[..., :3] - fb16[..., :3]).max() <= 2.0 ** -7


def test_srgb_mode(splat, oracle, golden_dir):
    """isFramebufferSRGBEnabled (VR path): radiance goes sRGB -> linear before blending (splat_vert.glsl:209-218)."""
    g = np.load(os.path.join(golden_dir, "synth_2k_deg3_320x180.npz"))
    gc = splat.GaussianCloud()
    gc.SetRawData(g["aos"], True)
    r = _new(splat)
    assert r.Init(gc, True, False)
    fb = _frame(r, g["camera"], g["proj"], 320, 180)
    want, rec, _, _ = oracle.render(g["aos"], True, g["camera"], g["proj"], [0, 0, 320, 180], NF, srgb=True, want_records=True)
    got = r.Tap(api.TAP_RECORDS)
    vis = rec["visible"] == 1
    for f in "rgb":  # powf differs by an ulp or two between libm and CUDA
        np.testing.assert_allclose(got[f][vis], rec[f][vis], rtol=1e-5, atol=1e-6, equal_nan=True)
    ok = np.isfinite(want).all(axis=-1)
    d = np.abs(fb - want)[ok]
    assert d.max() <= 5e-3


def test_idempotent_and_resolution_change(splat, golden_dir):
    g = np.load(os.path.join(golden_dir, "synth_2k_deg3_320x180.npz"))
    r = _renderer(splat, g["aos"], True)
    a = _frame(r, g["camera"], g["proj"], 320, 180)
    b = _frame(r, g["camera"], g["proj"], 320, 180)
    np.testing.assert_array_equal(a, b)  # deterministic: no atomics-order dependence anywhere
    c = _frame(r, g["camera"], scenes.perspective(scenes.FOVY, 400 / 300), 400, 300)
    assert c.shape == (300, 400, 4)
    d = _frame(r, g["camera"], g["proj"], 320, 180)
    np.testing.assert_array_equal(a, d)


def test_render_without_sort_and_changed_camera(splat, oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "synth_2k_deg3_320x180.npz"))
    r = _renderer(splat, g["aos"], True)
    vp = [0, 0, 320, 180]
    r.Render(g["camera"], g["proj"], vp, NF)  # Render alone: sorts implicitly
    _check_image(r.ReadFramebuffer(), g["fb"])
    cam2 = scenes.translate(0.5, -0.25, 9.0)
    r.Render(cam2, g["proj"], vp, NF)  # arguments differ from the last Sort: the image follows the arguments
    _check_image(r.ReadFramebuffer(), oracle.render(g["aos"], True, cam2, g["proj"], vp, NF))


def test_empty_and_all_culled(splat):
    r = _renderer(splat, np.zeros((0, 25), np.float32), False)
    fb = _frame(r, scenes.translate(0, 0, 3), scenes.perspective(scenes.FOVY, 1.0), 64, 64)
    np.testing.assert_array_equal(fb, np.broadcast_to(np.array([0, 0, 0, 1], np.float32), (64, 64, 4)))
    a = np.zeros((10, 25), np.float32)
    a[:, 2] = 50.0  # behind the camera
    a[:, 3] = 1.0
    a[:, 16] = a[:, 20] = a[:, 24] = 0.01
    r = _renderer(splat, a, False)
    fb = _frame(r, scenes.translate(0, 0, 3), scenes.perspective(scenes.FOVY, 1.0), 64, 64)
    assert fb[..., :3].max() == 0 and r.GetStats().num_visible == 0 and r.GetStats().num_instances == 0


def test_error_behaviour(splat, golden_dir):
    g = np.load(os.path.join(golden_dir, "synth_2k_deg3_320x180.npz"))
    r = _new(splat)
    with pytest.raises(splat.SplatError):  # Sort before Init
        r.Sort(g["camera"], g["proj"], [0, 0, 320, 180], NF)
    r = _renderer(splat, g["aos"], True)
    with pytest.raises(splat.SplatError):  # viewport origin must be 0
        r.Sort(g["camera"], g["proj"], [8, 0, 320, 180], NF)
    with pytest.raises(splat.SplatError):
        r.Sort(g["camera"], g["proj"], [0, 0, 0, 180], NF)
    with pytest.raises(splat.SplatError):  # nothing rendered yet
        r.ReadFramebuffer()
    gc = splat.GaussianCloud()
    with pytest.raises(splat.SplatError):
        gc.SetRawData(np.zeros((4, 30), np.float32), False)  # stride must be 100 or 244


@pytest.mark.parametrize("n,expect_path", [(1500, "small"), (6000, "large-smem"), (40000, "global")])
def test_dense_tile_segments(splat, oracle, scratch, n, expect_path):
    """thousands of gaussians inside ONE 16x16 tile: exercises every size class of the per-tile segment sort
    (<= 2048 and <= 12288 in shared memory, longer through global memory) and deep front-to-back lists."""
    a = scenes.synthetic_cloud_arrays(n, 0, seed=77, half_extent=0.02)
    a["log_scale"][:] = np.log(0.004)
    a["opacity"][:] = -3.0  # alpha ~ 0.047: thousands of layers stay above the transmittance early-out
    path = scenes.write_ply_from_arrays(str(scratch / f"dense_{n}.ply"), a)
    gc = splat.GaussianCloud()
    assert gc.ImportPly(path)
    r = _new(splat)
    r.SetOption(api.OPT_INSTANCES_PER_GAUSSIAN, 8.0)
    assert r.Init(gc, False, False)
    cam, proj = scenes.translate(0.0, 0.0, 6.0), scenes.perspective(scenes.FOVY, 1.0)
    fb = _frame(r, cam, proj, 64, 64)
    aos, full = oracle.import_ply(path, True)
    want, rec, order, _ = oracle.render(aos, full, cam, proj, [0, 0, 64, 64], NF, want_records=True)
    _check_records(r.Tap(api.TAP_RECORDS), rec)
    _check_tile_lists(r, rec, order, 64, 64)
    ranges = r.Tap(api.TAP_TILE_RANGES)
    longest = int((ranges[:, 1] - ranges[:, 0]).max())
    assert {"small": longest <= 2048, "large-smem": 2048 < longest <= 12288, "global": longest > 12288}[expect_path], longest
    _check_image(fb, want)
    r.SetOption(api.OPT_EARLY_OUT_T, 0.0)
    _check_image(_frame(r, cam, proj, 64, 64), want, exact_mode=True)


def test_sort_tile_variants_give_identical_lists(splat, golden_dir):
    """the onesweep sorter's two tile sizes (4096 / 2048 keys per CTA) are the same stable sort"""
    if SORT_MODE == api.SORT_BINNED:
        pytest.skip("LSD paths only")
    g = np.load(os.path.join(golden_dir, "synth_2k_deg3_320x180.npz"))
    out = []
    for tile in (4096, 2048):
        r = _renderer(splat, g["aos"], True)
        r.SetOption(api.OPT_SORT_TILE, tile)
        fb = _frame(r, g["camera"], g["proj"], 320, 180)
        _check_tile_lists(r, g["records"], g["order"], 320, 180)
        out.append((r.Tap(api.TAP_KEYS), r.Tap(api.TAP_VALS), fb))
    np.testing.assert_array_equal(out[0][0], out[1][0])
    np.testing.assert_array_equal(out[0][1], out[1][1])
    np.testing.assert_array_equal(out[0][2], out[1][2])
