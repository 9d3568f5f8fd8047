"""CPU: the oracle's restatement of the reference's VR draw (Sort for eye 0, Render for both, src/app.cpp:603-607)."""
import numpy as np
import pytest

from splatapult_b200 import api, scenes

NF = [scenes.Z_NEAR, scenes.Z_FAR]
W, H = 320, 200


def _aos(oracle, scratch, deg=0, n=6000):
    path = str(scratch / f"oracle_shared_{deg}.ply")
    scenes.write_synthetic_ply(path, n, deg, seed=31 + deg)
    return oracle.import_ply(path, True)


def _frames(oracle):
    proj, vp = scenes.perspective(scenes.FOVY, W / H), [0, 0, W, H]
    left = scenes.look_at_camera((-0.032, 0.10, 6.0), (0.25, 0.0, 0.0))
    right = scenes.look_at_camera((0.032, 0.10, 6.0), (-0.35, 0.1, 0.0))
    return oracle.Frame(left, proj, vp, NF), oracle.Frame(right, proj, vp, NF)


@pytest.mark.parametrize("deg", [0, 3])
def test_same_frame_is_the_mono_path(oracle, scratch, deg):
    aos, full = _aos(oracle, scratch, deg)
    f0, _ = _frames(oracle)
    assert oracle.project(aos, full, f0).tobytes() == oracle.project_shared(aos, full, f0, f0).tobytes()


def test_membership_and_key_from_the_sort_geometry_from_the_draw(oracle, scratch):
    aos, full = _aos(oracle, scratch)
    f0, f1 = _frames(oracle)
    r0, r1, rs = oracle.project(aos, full, f0), oracle.project(aos, full, f1), oracle.project_shared(aos, full, f0, f1)
    # keys (set whenever the pre-sort cull passes) are the sorted eye's, for every gaussian
    np.testing.assert_array_equal(rs["key"], r0["key"])
    assert np.any(r0["key"] != r1["key"])
    vis = rs["visible"] == 1
    assert vis.sum() > 1000
    # whatever is drawn for both arguments has the draw eye's geometry and colour, bit for bit
    both = vis & (r1["visible"] == 1)
    assert both.sum() > 1000
    for f in ("px", "py", "i00", "i01", "i10", "i11", "r", "g", "b", "alpha"):
        np.testing.assert_array_equal(rs[f][both], r1[f][both], err_msg=f)
    # nothing is drawn that the Sort's cull rejected (key 0 and not visible there), even if the draw eye would see it
    rejected = (r0["key"] == 0) & (r0["visible"] == 0)
    assert not np.any(vis & rejected)


def test_order_is_the_sorted_eyes(oracle, scratch):
    aos, full = _aos(oracle, scratch)
    f0, f1 = _frames(oracle)
    r0, rs = oracle.project(aos, full, f0), oracle.project_shared(aos, full, f0, f1)
    o0, os_ = oracle.sort(r0), oracle.sort(rs)
    # the shared list is the sorted eye's list minus what the draw eye's geometry stage discards, in the same order ...
    keep = rs["visible"][o0] == 1
    extra = np.setdiff1d(os_, o0)  # ... plus gaussians only the sorted eye's geometry stage discarded, at their key's place
    np.testing.assert_array_equal(o0[keep], os_[~np.isin(os_, extra)])
    assert np.all(np.diff(rs["key"][os_].astype(np.int64)) >= 0)
    # and it is not the draw eye's own order
    r1 = oracle.project(aos, full, f1)
    o1 = oracle.sort(r1)
    assert len(o1) != len(os_) or not np.array_equal(o1, os_)
    # the two images differ only where the order matters
    a, b = oracle.composite(rs, os_, W, H), oracle.composite(r1, o1, W, H)
    assert 0 < np.abs(a - b).max() < 0.5 and np.abs(a - b).mean() < 1e-2


def test_option_is_accepted_without_a_gpu(splat):
    r = splat.SplatRenderer()
    r.SetOption(api.OPT_SHARED_SORT, 1)
    r.SetOption(api.OPT_SHARED_SORT, 0)
    with pytest.raises(splat.SplatError):
        r.SetOption(api.OPT_SHARED_SORT, 2)
