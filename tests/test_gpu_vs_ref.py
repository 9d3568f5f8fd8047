"""The CUDA path directly against the reference's own shader/host source (oracle/_ref, prebuilt; see
tests/test_ref_pin.py): records bit-exact, framebuffer within the stated tolerance."""
import os

import numpy as np
import pytest

from splatapult_b200 import api, scenes

pytestmark = pytest.mark.gpu
NF = [scenes.Z_NEAR, scenes.Z_FAR]


@pytest.fixture(scope="module")
def ref():
    from oracle import splat_ref

    if not splat_ref.available():
        pytest.skip("oracle/_ref is not built and /root/reference is absent")
    splat_ref.lib()
    return splat_ref


@pytest.mark.parametrize("sort_mode", [api.SORT_LSD, api.SORT_BINNED])
def test_config1_against_reference_source(splat, ref, golden_dir, sort_mode):
    """BASELINE config 1 (data/test.ply, 640x480, start-up camera): CUDA vs the reference's GLSL executed on the CPU."""
    ply = os.path.join(golden_dir, "reference_test.ply")
    aos, full = ref.import_ply(ply, True)  # the reference's ImportPly
    gc = splat.GaussianCloud()
    assert gc.ImportPly(ply)
    np.testing.assert_array_equal(gc.GetRawData(), aos)
    cam, proj, vp = scenes.test_ply_camera(), scenes.perspective(scenes.FOVY, 640 / 480), [0, 0, 640, 480]
    want_fb, want_rec, want_order = ref.render(aos, full, cam, proj, vp, NF)
    r = splat.SplatRenderer()
    r.SetOption(api.OPT_SORT_MODE, sort_mode)
    assert r.Init(gc, False, False)
    r.Sort(cam, proj, vp, NF)
    r.Render(cam, proj, vp, NF)
    fb = r.ReadFramebuffer()
    rec = r.Tap(api.TAP_RECORDS)
    for f in want_rec.dtype.names:
        np.testing.assert_array_equal(rec[f], want_rec[f], err_msg=f)
    d = np.abs(fb - want_fb)
    assert d.max() <= 5e-3 and d.mean() <= 1e-4


def test_synthetic_deg3_against_reference_source(splat, ref, scratch):
    p = str(scratch / "vsref.ply")
    scenes.write_synthetic_ply(p, 4000, 3, seed=44)
    aos, full = ref.import_ply(p, True)
    cam, proj, vp = scenes.translate(0.2, 0.1, 9.5), scenes.perspective(scenes.FOVY, 320 / 180), [0, 0, 320, 180]
    want_fb, want_rec, want_order = ref.render(aos, full, cam, proj, vp, NF)
    gc = splat.GaussianCloud()
    assert gc.ImportPly(p)
    r = splat.SplatRenderer()
    assert r.Init(gc, False, False)
    r.Sort(cam, proj, vp, NF)
    r.Render(cam, proj, vp, NF)
    fb, rec = r.ReadFramebuffer(), r.Tap(api.TAP_RECORDS)
    vis = want_rec["visible"] == 1
    np.testing.assert_array_equal(rec["visible"], want_rec["visible"])
    for f in ("px", "py", "i00", "i01", "i10", "i11", "r", "g", "b", "alpha", "key"):
        np.testing.assert_array_equal(rec[f][vis], want_rec[f][vis], err_msg=f)
    # the sorted tile lists follow the reference's draw order
    keys, vals = r.Tap(api.TAP_KEYS), r.Tap(api.TAP_VALS)
    rank = np.full(len(want_rec), -1, np.int64)
    rank[want_order] = np.arange(len(want_order))
    tiles = keys >> np.uint64(32)
    same = tiles[1:] == tiles[:-1]
    assert np.all(rank[vals] >= 0) and np.all(rank[vals][1:][same] > rank[vals][:-1][same])
    d = np.abs(fb - want_fb)
    assert d.max() <= 5e-3 and np.quantile(d, 0.999) <= 1e-3 and d.mean() <= 1e-4
