"""The compositor's per-entry sub-block mask (SPL_OPT_SUBBLOCK_CULL) only skips lanes whose weight the reference's
fragment shader discards anyway (w <= 1/256, shader/splat_frag.glsl:38): a culled lane adds tw = 0, which leaves the
accumulators bit-unchanged.  So the framebuffer must be bit-identical with culling off (0), footprint-box culling
(1) and exact-ellipse culling (2, default) -- any over-culling shows up as a differing pixel."""
import math
import os

import numpy as np
import pytest

from splatapult_b200 import api, scenes

pytestmark = pytest.mark.gpu
NF = [scenes.Z_NEAR, scenes.Z_FAR]


def _render_modes(splat, gc, cam, proj, w, h, early_out, fb_format=api.FB_FP32):
    out = []
    for mode in (0, 1, 2):
        r = splat.SplatRenderer(fb_format=fb_format)
        r.SetOption(api.OPT_SUBBLOCK_CULL, mode)
        r.SetOption(api.OPT_EARLY_OUT_T, early_out)
        assert r.Init(gc, False, False)
        r.Sort(cam, proj, [0, 0, w, h], NF)
        r.Render(cam, proj, [0, 0, w, h], NF)
        out.append(r.ReadFramebuffer().copy())
    return out


@pytest.mark.parametrize("early_out", [0.0, 1e-4])
def test_cull_modes_bit_identical_golden_scene(splat, golden_dir, early_out):
    g = np.load(os.path.join(golden_dir, "synth_2k_deg3_320x180.npz"))
    gc = splat.GaussianCloud()
    gc.SetRawData(g["aos"], True)
    none, box, exact = _render_modes(splat, gc, g["camera"], g["proj"], 320, 180, early_out)
    assert none[..., :3].max() > 0.1
    np.testing.assert_array_equal(box, none)
    np.testing.assert_array_equal(exact, none)


def test_cull_modes_bit_identical_reference_ply(splat, golden_dir):
    gc = splat.GaussianCloud()
    assert gc.ImportPly(os.path.join(golden_dir, "reference_test.ply"))
    proj = scenes.perspective(scenes.FOVY, 640 / 480)
    none, box, exact = _render_modes(splat, gc, scenes.test_ply_camera(), proj, 640, 480, 0.0)
    np.testing.assert_array_equal(box, none)
    np.testing.assert_array_equal(exact, none)


@pytest.mark.parametrize("fb_format", [api.FB_FP32, api.FB_FP16])
def test_cull_modes_bit_identical_needles_and_giants(splat, scratch, fb_format):
    """Stress for the ellipse test: needle-shaped splats (axis ratio up to 1e4: the conic's determinant cancels
    almost completely), screen-filling giants, near-transparent ones (r2 -> 0) and opaque ones, camera inside the
    cloud so that centres lie far outside the tiles they touch."""
    rng = np.random.default_rng(77)
    n = 6000
    a = scenes.synthetic_cloud_arrays(n, 0, seed=77, half_extent=3.0)
    ls = np.empty((n, 3), np.float32)
    ls[:, 0] = rng.uniform(math.log(0.3), math.log(6.0), n)        # long axis: up to metres
    ls[:, 1] = rng.uniform(math.log(0.0005), math.log(0.02), n)    # thin
    ls[:, 2] = rng.uniform(math.log(0.0005), math.log(0.5), n)
    third = n // 3
    ls[:third] = rng.uniform(math.log(0.002), math.log(0.08), (third, 3))  # ordinary blobs
    a["log_scale"] = ls
    a["opacity"] = rng.uniform(-5.6, 6.0, (n, 1)).astype(np.float32)  # alpha from ~1/270 (below the discard level) to ~1
    path = scenes.write_ply_from_arrays(os.path.join(str(scratch), "needles.ply"), a)
    gc = splat.GaussianCloud()
    assert gc.ImportPly(path)
    w, h = 500, 300  # not a multiple of the tile size
    proj = scenes.perspective(scenes.FOVY, w / h)
    for cam in (scenes.translate(0.0, 0.0, 7.0), scenes.translate(0.4, -0.2, 1.0)):
        none, box, exact = _render_modes(splat, gc, cam, proj, w, h, 0.0, fb_format)
        assert np.isfinite(none).all() and none[..., :3].max() > 0.05
        np.testing.assert_array_equal(box, none)
        np.testing.assert_array_equal(exact, none)


def test_cull_option_validation(splat):
    r = splat.SplatRenderer()
    with pytest.raises(api.SplatError):
        r.SetOption(api.OPT_SUBBLOCK_CULL, 3)
