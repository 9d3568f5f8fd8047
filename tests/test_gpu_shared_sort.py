"""the reference's VR call pattern -- Sort for eye 0, Render for both eyes (src/app.cpp:603-607).

Default (SPL_OPT_SHARED_SORT = 0): a Render whose arguments differ from the last Sort's sorts again, every image is that
of its own arguments (tests/test_gpu_parity.py::test_render_without_sort_and_changed_camera).
SPL_OPT_SHARED_SORT = 1: the reference's behaviour.  The second eye's draw runs over the list and in the order the Sort
left: list membership and depth key from the Sort's matrices (presort_compute.glsl:42-56), everything else from the
draw's (splat_vert.glsl / splat_geom.glsl).  Oracle: oracle.project_shared + sort + composite.  Same bars as the mono
path: records / keys / tile lists bit-exact, framebuffer within the DESIGN bars.
"""
import numpy as np
import pytest

from splatapult_b200 import api, scenes

from test_gpu_parity import _check_image, _check_records, _check_tile_lists

pytestmark = pytest.mark.gpu

NF = [scenes.Z_NEAR, scenes.Z_FAR]
W, H = 640, 360


def _eyes():
    """two eyes 6.4 cm apart with some toe-in, so that depths (and with them the exact orders) differ."""
    left = scenes.look_at_camera((-0.032, 0.10, 6.0), (0.25, 0.0, 0.0))
    right = scenes.look_at_camera((0.032, 0.10, 6.0), (-0.35, 0.1, 0.0))
    return left, right


def _cloud(splat, oracle, scratch, deg, n=20000, seed=77):
    path = str(scratch / f"shared_{deg}_{n}.ply")
    scenes.write_synthetic_ply(path, n, deg, seed=seed + deg)
    gc = splat.GaussianCloud()
    assert gc.ImportPly(path)
    aos, full = oracle.import_ply(path, True)
    return gc, aos, full


def _renderer(splat, gc, shared, sort_mode=api.SORT_LSD, fb_format=0):
    r = splat.SplatRenderer(fb_format=fb_format)
    r.SetOption(api.OPT_SORT_MODE, sort_mode)
    r.SetOption(api.OPT_SHARED_SORT, 1 if shared else 0)
    assert r.Init(gc, False, False), getattr(r, "last_error", "")
    return r


@pytest.mark.parametrize("deg", [0, 3])
def test_sorted_eye_is_unchanged_by_the_option(splat, oracle, scratch, deg):
    gc, aos, full = _cloud(splat, oracle, scratch, deg)
    left, _ = _eyes()
    proj, vp = scenes.perspective(scenes.FOVY, W / H), [0, 0, W, H]
    frames, recs, lists = [], [], []
    for shared in (False, True):
        r = _renderer(splat, gc, shared)
        r.Sort(left, proj, vp, NF)
        r.Render(left, proj, vp, NF)
        frames.append(r.ReadFramebuffer())
        recs.append(r.Tap(api.TAP_RECORDS))
        lists.append((r.Tap(api.TAP_KEYS), r.Tap(api.TAP_VALS), r.Tap(api.TAP_TILE_RANGES)))
    np.testing.assert_array_equal(frames[0], frames[1])
    assert recs[0].tobytes() == recs[1].tobytes()
    for a, b in zip(lists[0], lists[1]):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("deg", [0, 3])
def test_second_eye_in_the_first_eyes_order(splat, oracle, scratch, deg):
    gc, aos, full = _cloud(splat, oracle, scratch, deg)
    left, right = _eyes()
    proj, vp = scenes.perspective(scenes.FOVY, W / H), [0, 0, W, H]
    r = _renderer(splat, gc, True)
    r.SetOption(api.OPT_EARLY_OUT_T, 0.0)
    r.Sort(left, proj, vp, NF)
    r.Render(left, proj, vp, NF)
    fb_left = r.ReadFramebuffer()
    want_left, rec_left, order_left, _ = oracle.render(aos, full, left, proj, vp, NF, want_records=True)
    _check_image(fb_left, want_left, exact_mode=True)

    r.Render(right, proj, vp, NF)  # no Sort: the VR pattern
    fb_right = r.ReadFramebuffer()
    f_left, f_right = oracle.Frame(left, proj, vp, NF), oracle.Frame(right, proj, vp, NF)
    rec = oracle.project_shared(aos, full, f_left, f_right)
    order = oracle.sort(rec)
    _check_records(r.Tap(api.TAP_RECORDS), rec)
    _check_tile_lists(r, rec, order, W, H)
    _check_image(fb_right, oracle.composite(rec, order, W, H), exact_mode=True)
    # the second eye cost one projection, one expansion, the tile-digit passes and the composite: no depth passes
    st = r.GetStats()
    tiles = int(st.tiles_x * st.tiles_y)
    tile_passes = (max(tiles - 1, 1).bit_length() + 7) // 8
    assert st.kernel_launches == 2 + tile_passes + 1, st.kernel_launches
    # the test discriminates: the exact order for the right eye is a different one
    rec_exact = oracle.project(aos, full, f_right)
    order_exact = oracle.sort(rec_exact)
    assert len(order_exact) != len(order) or not np.array_equal(order_exact, order)

    # and back: Render for the sorted eye again re-projects to it (same list, same order as the Sort's own frame)
    r.Render(left, proj, vp, NF)
    np.testing.assert_array_equal(r.ReadFramebuffer(), fb_left)
    _check_records(r.Tap(api.TAP_RECORDS), rec_left)


def test_next_sort_starts_a_fresh_frame(splat, oracle, scratch):
    gc, aos, full = _cloud(splat, oracle, scratch, 0)
    left, right = _eyes()
    proj, vp = scenes.perspective(scenes.FOVY, W / H), [0, 0, W, H]
    r = _renderer(splat, gc, True)
    for _ in range(3):  # frames of a VR session
        r.Sort(left, proj, vp, NF)
        r.Render(left, proj, vp, NF)
        r.Render(right, proj, vp, NF)
    r.Sort(right, proj, vp, NF)
    r.Render(right, proj, vp, NF)
    _check_image(r.ReadFramebuffer(), oracle.render(aos, full, right, proj, vp, NF))
    rec = oracle.project(aos, full, oracle.Frame(right, proj, vp, NF))
    _check_tile_lists(r, rec, oracle.sort(rec), W, H)


def test_option_off_and_other_sort_paths_sort_again(splat, oracle, scratch):
    gc, aos, full = _cloud(splat, oracle, scratch, 0)
    left, right = _eyes()
    proj, vp = scenes.perspective(scenes.FOVY, W / H), [0, 0, W, H]
    want, rec, order, _ = oracle.render(aos, full, right, proj, vp, NF, want_records=True)
    for shared, mode in ((False, api.SORT_LSD), (True, api.SORT_LSD_FLAT), (True, api.SORT_BINNED)):
        r = _renderer(splat, gc, shared, sort_mode=mode)
        r.Sort(left, proj, vp, NF)
        r.Render(right, proj, vp, NF)
        _check_image(r.ReadFramebuffer(), want)
        _check_records(r.Tap(api.TAP_RECORDS), rec)
        _check_tile_lists(r, rec, order, W, H)


def test_other_viewport_sorts_again(splat, oracle, scratch):
    gc, aos, full = _cloud(splat, oracle, scratch, 0)
    left, right = _eyes()
    r = _renderer(splat, gc, True)
    r.Sort(left, scenes.perspective(scenes.FOVY, W / H), [0, 0, W, H], NF)
    proj2, vp2 = scenes.perspective(scenes.FOVY, 320 / 200), [0, 0, 320, 200]
    r.Render(right, proj2, vp2, NF)
    fb = r.ReadFramebuffer()
    assert fb.shape == (200, 320, 4)
    _check_image(fb, oracle.render(aos, full, right, proj2, vp2, NF))


def test_overflow_in_the_second_eye_is_redone(splat, oracle, scratch):
    a = scenes.synthetic_cloud_arrays(3000, 0, seed=5, half_extent=1.0)
    a["log_scale"] += 2.0
    path = scenes.write_ply_from_arrays(str(scratch / "shared_overflow.ply"), a)
    gc = splat.GaussianCloud()
    assert gc.ImportPly(path)
    aos, full = oracle.import_ply(path, True)
    r = splat.SplatRenderer()
    r.SetOption(api.OPT_SHARED_SORT, 1)
    r.SetOption(api.OPT_INSTANCES_PER_GAUSSIAN, 1.0)
    assert r.Init(gc, False, False)
    proj, vp = scenes.perspective(scenes.FOVY, 640 / 480), [0, 0, 640, 480]
    left, right = scenes.translate(-0.03, 0, 4), scenes.look_at_camera((0.03, 0.0, 4.0), (-0.2, 0.0, 0.0))
    r.Sort(left, proj, vp, NF)
    r.Render(left, proj, vp, NF)
    try:
        r.Render(right, proj, vp, NF)
    except splat.SplatError as e:  # the left eye's frame has already reported that it was incomplete (a notification:
        assert e.code == -4        # the Render itself was carried out)
    assert r.GetStats().overflow
    fb = r.ReadFramebuffer()  # grows, sorts for the left eye again, re-projects for the right one
    assert not r.GetStats().overflow
    rec = oracle.project_shared(aos, full, oracle.Frame(left, proj, vp, NF), oracle.Frame(right, proj, vp, NF))
    _check_image(fb, oracle.composite(rec, oracle.sort(rec), 640, 480))


def test_fp16_target_second_eye(splat, oracle, scratch):
    gc, aos, full = _cloud(splat, oracle, scratch, 3, n=8000)
    left, right = _eyes()
    proj, vp = scenes.perspective(scenes.FOVY, W / H), [0, 0, W, H]
    r = _renderer(splat, gc, True, fb_format=api.FB_FP16)
    r.Sort(left, proj, vp, NF)
    r.Render(left, proj, vp, NF)
    r.Render(right, proj, vp, NF)
    rec = oracle.project_shared(aos, full, oracle.Frame(left, proj, vp, NF), oracle.Frame(right, proj, vp, NF))
    _check_image(r.ReadFramebuffer().astype(np.float32), oracle.composite(rec, oracle.sort(rec), W, H), fp16=True)
