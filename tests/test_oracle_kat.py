"""Known-answer tests that PIN THE ORACLE (CPU only).

The reference ships no tests or golden vectors for this path, so the pins are
(a) the reference's only data fixture, data/test.ply, cross-checked against the constants hard-coded in
GaussianCloud::InitDebugCloud (gaussiancloud.cpp:505-578), and (b) closed-form answers derived from the shader
source (SURVEY.md 8c "Golden vectors / KATs").
"""
import math
import os

import numpy as np
import pytest

from splatapult_b200 import scenes

SH_C0 = 0.28209479177387814


def test_decode_reference_fixture(oracle, golden_dir):
    aos, full = oracle.import_ply(os.path.join(golden_dir, "reference_test.ply"), True)
    assert full and aos.shape == (16, 61)
    dbg = oracle.debug_cloud()  # the reference's own generator of this cloud
    # positions, alpha and colours are what InitDebugCloud hard-codes (alpha = sigmoid(100) == 1.0f)
    np.testing.assert_allclose(aos[:, 0:3], dbg[:, 0:3], atol=1e-6)
    np.testing.assert_array_equal(aos[:, 3], np.ones(16, np.float32))
    for blk in (4, 8, 12):
        np.testing.assert_allclose(aos[:, blk], dbg[:, blk], rtol=1e-6)
        assert np.all(aos[:, blk + 1:blk + 4] == 0)
    # the file stores scale = ln(0.05) and identity rotation -> Sigma = 0.0025 I (InitDebugCloud in memory uses 0.005)
    cov = aos[:, 16:25].reshape(16, 3, 3)
    np.testing.assert_allclose(cov, np.broadcast_to(np.eye(3, dtype=np.float32) * 0.0025, cov.shape), atol=1e-9)
    # colours: 0.5 + C0 * (+-1/(2 C0)) = {1, 0}
    rgb = 0.5 + SH_C0 * aos[:, [4, 8, 12]]
    assert np.allclose(np.sort(np.unique(np.round(rgb, 5))), [0.0, 1.0])
    # full-SH slots are all zero in this file
    assert np.all(aos[:, 25:61] == 0)


def test_decode_matches_golden(oracle, golden_dir):
    aos, _ = oracle.import_ply(os.path.join(golden_dir, "reference_test.ply"), True)
    np.testing.assert_array_equal(aos, np.load(os.path.join(golden_dir, "test_ply_decode.npy")))


def test_decode_nosh_layout(oracle, golden_dir):
    """--nosh (importFullSH=false): base 100-byte layout, degree-1 slots zero (gaussiancloud.cpp:316-332)."""
    aos, full = oracle.import_ply(os.path.join(golden_dir, "reference_test.ply"), False)
    assert not full and aos.shape == (16, 25)
    assert np.all(aos[:, [5, 6, 7, 9, 10, 11, 13, 14, 15]] == 0)


def test_perspective_matches_glm_formula(oracle):
    p = oracle.perspective(scenes.FOVY, 640 / 480, 0.1, 1000.0).reshape(4, 4)  # [col,row]
    t = math.tan(scenes.FOVY / 2)
    assert p[0, 0] == pytest.approx(1 / (640 / 480 * t), rel=1e-6)
    assert p[1, 1] == pytest.approx(1 / t, rel=1e-6)
    assert p[2, 2] == pytest.approx(-(1000.1) / (999.9), rel=1e-6)
    assert p[2, 3] == -1.0
    assert p[3, 2] == pytest.approx(-2 * 1000 * 0.1 / 999.9, rel=1e-6)
    np.testing.assert_array_equal(p.reshape(-1), scenes.perspective(scenes.FOVY, 640 / 480, 0.1, 1000.0))


def _single(alpha=1.0, sigma2=0.0025, dc=(1.7724539, 1.7724539, 1.7724539), pos=(0, 0, 0)):
    a = np.zeros((1, 25), np.float32)
    a[0, 0:3] = pos
    a[0, 3] = alpha
    a[0, 4], a[0, 8], a[0, 12] = dc
    a[0, 16] = a[0, 20] = a[0, 24] = sigma2
    return a


def test_key_kat(oracle):
    """depth 3.0, far 1000 -> key 0xFF3B6459 (presort_compute.glsl:53, SURVEY 8c iii)."""
    fr = oracle.Frame(scenes.translate(0, 0, 3), scenes.perspective(scenes.FOVY, 640 / 480), [0, 0, 640, 480], [0.1, 1000.0])
    r = oracle.project(_single(), False, fr)
    assert r["visible"][0] == 1 and r["key"][0] == 0xFF3B6459


def test_on_axis_isotropic_gaussian(oracle):
    """sigma=0.05 at d=3, H=480: cov2D = (sigma*SY*H/(2d))^2 + 0.3 = 93.55; centre pixel = alpha*rgb (SURVEY 8c ii)."""
    proj = scenes.perspective(scenes.FOVY, 640 / 480)
    fr = oracle.Frame(scenes.translate(0, 0, 3), proj, [0, 0, 640, 480], [0.1, 1000.0])
    r = oracle.project(_single(), False, fr)
    sy = proj.reshape(4, 4)[1, 1]
    expect = (0.05 * sy * 480 / (2 * 3.0)) ** 2 + 0.3
    det = r["i00"][0] * r["i11"][0] - r["i01"][0] * r["i10"][0]
    assert r["i11"][0] / det == pytest.approx(expect, rel=1e-5)
    assert expect == pytest.approx(93.55, abs=0.01)
    assert (r["px"][0], r["py"][0]) == (320.0, 240.0)
    np.testing.assert_allclose([r["r"][0], r["g"][0], r["b"][0]], 1.0, atol=1e-6)
    fb = oracle.composite(r, oracle.sort(r), 640, 480)
    # the four pixels around the centre sit at d = (+-0.5, +-0.5): value = exp(-0.5 * 0.5 / 93.55)
    np.testing.assert_allclose(fb[239:241, 319:321, :3], math.exp(-0.25 / expect), rtol=1e-6)
    assert np.all(fb[..., 3] == 1.0)


def test_near_and_guard_culls(oracle):
    """K6: ndc.z < 0.25 <=> view depth < 0.26662 for n=0.1, f=1000 (splat_geom.glsl:48); K1: |ndc.xy| < 1.5, w > 0."""
    proj = scenes.perspective(scenes.FOVY, 640 / 480)
    vis = []
    for d in (0.2660, 0.2672, 5.0, 999.0, 1000.5, -1.0):
        fr = oracle.Frame(scenes.translate(0, 0, d), proj, [0, 0, 640, 480], [0.1, 1000.0])
        vis.append(int(oracle.project(_single(), False, fr)["visible"][0]))
    assert vis == [0, 1, 1, 1, 0, 0]
    fr = oracle.Frame(scenes.translate(0, 0, 3), proj, [0, 0, 640, 480], [0.1, 1000.0])
    half_w = 3 * math.tan(scenes.FOVY / 2) * 640 / 480  # ndc.x == 1 here
    for x, expect in ((1.49 * half_w, 1), (1.51 * half_w, 0), (-1.51 * half_w, 0)):
        assert int(oracle.project(_single(pos=(x, 0, 0)), False, fr)["visible"][0]) == expect


def test_sh_basis_matches_inria_constants(oracle):
    """degree-3 basis signs/constants (splat_vert.glsl:65-105) == INRIA SH_C0..C3 for a unit direction."""
    rng = np.random.default_rng(3)
    proj = scenes.perspective(scenes.FOVY, 640 / 480)
    C1 = 0.4886025119029199
    C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
    C3 = [-0.5900435899266435, 2.8906114426405543, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
          1.445305721320277, -0.5900435899266435]
    for _ in range(8):
        a = np.zeros((1, 61), np.float32)
        a[0, 0:3] = rng.uniform(-0.5, 0.5, 3)
        a[0, 3] = 1.0
        a[0, 16] = a[0, 20] = a[0, 24] = 0.0025
        sh = rng.normal(0, 0.3, (3, 16)).astype(np.float32)
        blocks = [[4, 25, 29, 33], [8, 37, 41, 45], [12, 49, 53, 57]]
        for c in range(3):
            for k in range(16):
                a[0, blocks[c][k >> 2] + (k & 3)] = sh[c, k]
        cam = scenes.translate(0, 0, 3)
        fr = oracle.Frame(cam, proj, [0, 0, 640, 480], [0.1, 1000.0])
        r = oracle.project(a, True, fr)
        d = a[0, 0:3].astype(np.float64) - np.array([0, 0, 3.0])
        x, y, z = d / np.linalg.norm(d)
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        for c, name in enumerate("rgb"):
            s = sh[c].astype(np.float64)
            res = SH_C0 * s[0] - C1 * y * s[1] + C1 * z * s[2] - C1 * x * s[3]
            res += C2[0] * xy * s[4] + C2[1] * yz * s[5] + C2[2] * (2 * zz - xx - yy) * s[6] + C2[3] * xz * s[7] + C2[4] * (xx - yy) * s[8]
            res += (C3[0] * y * (3 * xx - yy) * s[9] + C3[1] * xy * z * s[10] + C3[2] * y * (4 * zz - xx - yy) * s[11] +
                    C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * s[12] + C3[4] * x * (4 * zz - xx - yy) * s[13] +
                    C3[5] * z * (xx - yy) * s[14] + C3[6] * x * (xx - 3 * yy) * s[15])
            assert r[name][0] == pytest.approx(res + 0.5, abs=2e-6)


def test_two_overlapping_splats_closed_form_over(oracle):
    """'over' of two splats at the same pixel: C = w_near*c_near + (1-w_near)*w_far*c_far (SURVEY 8c vi)."""
    a = np.concatenate([_single(alpha=0.6, dc=(1.7724539, -1.7724539, -1.7724539), pos=(0, 0, 0)),      # red, far
                        _single(alpha=0.5, dc=(-1.7724539, 1.7724539, -1.7724539), pos=(0, 0, 1.0))])  # green, near
    proj = scenes.perspective(scenes.FOVY, 640 / 480)
    fr = oracle.Frame(scenes.translate(0, 0, 3), proj, [0, 0, 640, 480], [0.1, 1000.0])
    rec = oracle.project(a, False, fr)
    order = oracle.sort(rec)
    assert list(order) == [0, 1]  # back to front: far (depth 3) first, then near (depth 2)
    fb = oracle.composite(rec, order, 640, 480)

    def w(i):
        r = rec[i]
        dx, dy = 320.5 - r["px"], 240.5 - r["py"]
        q = r["i00"] * dx * dx + (r["i01"] + r["i10"]) * dx * dy + r["i11"] * dy * dy
        return float(r["alpha"]) * math.exp(-0.5 * q)

    w_far, w_near = w(0), w(1)
    np.testing.assert_allclose(fb[240, 320, :3], [(1 - w_near) * w_far, w_near, 0.0], atol=2e-6)


def test_discard_threshold(oracle):
    """a splat with alpha <= 1/256 never contributes; alpha slightly above contributes only near its centre."""
    proj = scenes.perspective(scenes.FOVY, 640 / 480)
    fr = oracle.Frame(scenes.translate(0, 0, 3), proj, [0, 0, 640, 480], [0.1, 1000.0])
    for alpha, lit in ((1 / 256, False), (1 / 256 + 1e-4, True)):
        rec = oracle.project(_single(alpha=alpha), False, fr)
        fb = oracle.composite(rec, oracle.sort(rec), 640, 480)
        assert bool(fb[..., :3].max() > 0) == lit


def test_binned_equals_bruteforce(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "synth_2k_deg3_320x180.npz"))
    fb_b = oracle.composite(g["records"], g["order"], 320, 180)
    fb_f = oracle.composite(g["records"], g["order"], 320, 180, bruteforce=True)
    np.testing.assert_array_equal(fb_b, fb_f)
    np.testing.assert_array_equal(fb_b, g["fb"])


def test_render_matches_golden(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "test_ply_640x480.npz"))
    aos = np.load(os.path.join(golden_dir, "test_ply_decode.npy"))
    fb, rec, order, _ = oracle.render(aos, True, g["camera"], g["proj"], [0, 0, 640, 480], [0.1, 1000.0], want_records=True)
    np.testing.assert_array_equal(order, g["order"])
    np.testing.assert_array_equal(rec, g["records"])
    np.testing.assert_array_equal(fb, g["fb"])
    # all 16 debug splats are in view, the image is not empty and alpha stays 1
    assert len(order) == 16 and fb[..., :3].max() > 0.9 and np.all(np.abs(fb[..., 3] - 1) < 1e-6)


def test_fp16_target_rounds_each_blend(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "synth_2k_deg3_320x180.npz"))
    fb32 = g["fb"]
    fb16 = oracle.composite(g["records"], g["order"], 320, 180, fp16=True)
    assert np.array_equal(fb16, fb16.astype(np.float16).astype(np.float32))  # every value is a binary16
    assert 0 < np.abs(fb16 - fb32).max() < 2 ** -7


def test_sort_oracle_is_stable(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "sort_keys_45bit.npz"))
    ko, vo = oracle.sort_pairs_u64(g["keys"], g["vals"], 0, 45)
    idx = np.argsort(g["keys"], kind="stable")
    np.testing.assert_array_equal(ko, g["keys"][idx])
    np.testing.assert_array_equal(vo, g["vals"][idx])
    np.testing.assert_array_equal(ko, g["keys_sorted"])
