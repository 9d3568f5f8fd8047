"""Framebuffer egress: float target -> 8-bit PPM, top row first, clamped like the RGBA8 window target."""
import os

import numpy as np
import pytest


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_write_ppm(splat, scratch, golden_dir, dtype):
    fb = np.load(os.path.join(golden_dir, "test_ply_640x480.npz"))["fb"].astype(dtype)
    fb[0, 0, :3] = [2.0, -1.0, 0.5]  # out-of-range values clamp (float targets are unclamped)
    path = str(scratch / f"out_{np.dtype(dtype).name}.ppm")
    splat.write_ppm(path, fb)
    raw = open(path, "rb").read()
    header = b"P6\n640 480\n255\n"
    assert raw.startswith(header) and len(raw) == len(header) + 640 * 480 * 3
    img = np.frombuffer(raw[len(header):], np.uint8).reshape(480, 640, 3)
    want = np.floor(np.clip(fb[::-1, :, :3].astype(np.float32), 0, 1) * 255.0 + 0.5).astype(np.uint8)
    np.testing.assert_array_equal(img, want)
    np.testing.assert_array_equal(img[-1, 0], [255, 0, 128])  # framebuffer row 0 is the LAST image row
    assert img.max() == 255 and (img.sum(axis=2) > 0).mean() > 0.1  # the debug splats are in the picture


def test_write_ppm_rejects_bad_input(splat, scratch):
    with pytest.raises(ValueError):
        splat.write_ppm(str(scratch / "x.ppm"), np.zeros((4, 4, 3), np.float32))
    with pytest.raises(splat.SplatError):
        splat.write_ppm(str(scratch / "no_such_dir" / "x.ppm"), np.zeros((4, 4, 4), np.float32))
