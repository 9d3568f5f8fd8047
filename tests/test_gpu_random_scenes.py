"""Randomised parity: 120 small scenes with extreme gaussians and cameras (the family tests/test_ref_pin.py runs through the
reference's own shader source and the oracle bit for bit) through the CUDA path -- records / keys / visible set bit-exact,
per-tile order, image bars of DESIGN.md section 2 -- on all three frame-sort modes, with the early-out on and off."""
import numpy as np
import pytest

from splatapult_b200 import api, scenes

pytestmark = pytest.mark.gpu
NF = [scenes.Z_NEAR, scenes.Z_FAR]
FIELDS = ("px", "py", "i00", "i01", "i10", "i11", "r", "g", "b", "alpha", "key")


def _scene(rng):
    n, deg = int(rng.integers(1, 400)), int(rng.choice([0, 3]))
    a = scenes.synthetic_cloud_arrays(n, deg, seed=int(rng.integers(0, 1 << 30)), half_extent=float(rng.choice([0.5, 2.0, 6.0])))
    k = int(rng.integers(0, 5))
    if k == 0:
        a["log_scale"][: n // 3] += rng.uniform(1.5, 5.0)  # splats covering much of the frame
    elif k == 1:
        a["log_scale"][: n // 3] -= rng.uniform(3.0, 9.0)  # sub-pixel splats (the +0.3 dilation keeps them alive)
    elif k == 2:
        a["log_scale"][:, 0] -= 8.0  # needles
    elif k == 3:
        a["opacity"][: n // 2] = rng.choice([-12.0, 12.0, 0.0, -5.54])  # alpha ~ 0, ~ 1, 0.5, just under 1/256
    w, h = [(64, 48), (97, 33), (33, 97), (160, 120), (640, 360)][int(rng.integers(0, 5))]
    ck = int(rng.integers(0, 5))
    if ck == 0:
        cam = scenes.translate(float(rng.normal()), float(rng.normal()), float(rng.uniform(0.2, 12)))
    elif ck == 1:
        cam = scenes.look_at_camera(tuple(rng.normal(size=3) * 3), (0, 0, 0))
    elif ck == 2:
        cam = scenes.look_at_camera(tuple(rng.normal(size=3) * 0.3), tuple(rng.normal(size=3)))  # inside the cloud
    elif ck == 3:
        cam = scenes.look_at_camera((0, 0, 40.0), (0, 0, 0))
    else:
        cam = scenes.look_at_camera(tuple(rng.normal(size=3) * 8), tuple(rng.normal(size=3) * 2))
    proj = scenes.perspective(float(rng.choice([scenes.FOVY, 0.3, 2.4])), w / h)
    return a, deg, w, h, cam, proj


@pytest.mark.parametrize("mode", [api.SORT_LSD, api.SORT_LSD_FLAT, api.SORT_BINNED], ids=["lsd", "lsd_flat", "binned"])
def test_random_extreme_scenes(splat, oracle, scratch, mode):
    rng = np.random.default_rng(424242)
    path = str(scratch / "random_gpu.ply")
    drawn = instances = 0
    for it in range(120):
        a, deg, w, h, cam, proj = _scene(rng)
        scenes.write_ply_from_arrays(path, a)
        aos, full = oracle.import_ply(path, True)
        want, wrec, worder, _ = oracle.render(aos, full, cam, proj, [0, 0, w, h], NF, want_records=True)
        gc = splat.GaussianCloud()
        assert gc.ImportPly(path) and gc.HasFullSH() == full
        np.testing.assert_array_equal(gc.GetRawData(), aos)
        r = splat.SplatRenderer()
        assert r.Init(gc, False, False), getattr(r, "last_error", "")
        r.SetOption(api.OPT_SORT_MODE, mode)
        for early_out in (None, 0.0):
            if early_out is not None:
                r.SetOption(api.OPT_EARLY_OUT_T, early_out)
            r.Sort(cam, proj, [0, 0, w, h], NF)
            r.Render(cam, proj, [0, 0, w, h], NF)
            fb = r.ReadFramebuffer()
            st = r.GetStats()
            assert not st.overflow, it
            got = r.Tap(api.TAP_RECORDS)
            np.testing.assert_array_equal(got["visible"], wrec["visible"], err_msg=f"scene {it}")
            np.testing.assert_array_equal(got["key"], wrec["key"], err_msg=f"scene {it}")
            vis = wrec["visible"] == 1
            for f in FIELDS:
                np.testing.assert_array_equal(got[f][vis].view(np.uint32), wrec[f][vis].view(np.uint32), err_msg=f"scene {it} {f}")
            # per-tile order = the oracle's back-to-front order restricted to the tile
            keys, vals = r.Tap(api.TAP_KEYS), r.Tap(api.TAP_VALS)
            assert len(keys) == st.num_instances and np.all(keys[1:] >= keys[:-1])
            rank = np.full(len(wrec), -1, np.int64)
            rank[worder] = np.arange(len(worder))
            rv = rank[vals]
            assert np.all(rv >= 0), it
            tiles = (keys >> np.uint64(32)).astype(np.int64)
            same = tiles[1:] == tiles[:-1]
            assert np.all(rv[1:][same] > rv[:-1][same]), it
            np.testing.assert_array_equal(fb[..., 3], 1.0)
            d = np.abs(fb[..., :3] - want[..., :3])
            assert d.max() <= 5e-3 and np.quantile(d, 0.999) <= 1e-3 and d.mean() <= (2e-5 if early_out == 0.0 else 1e-4), (it, d.max(), d.mean())
        drawn += len(worder)
        instances += st.num_instances
        r.close()
    assert drawn > 5000 and instances > drawn
