"""Benchmark / test scenes: the seeded synthetic clouds and cameras of SURVEY.md section 8(d).

Clouds are written as INRIA-layout binary PLY files so that the SAME .ply load path (GaussianCloud.ImportPly)
feeds the renderer, exactly as a splatapult user would.  Cameras are glm-layout column-major float32[16]
camera-to-world matrices, the type SplatRenderer::Sort/Render take (src/splatrenderer.h:32-37).
"""
from __future__ import annotations

import json
import math
import os

import numpy as np

# constants of the reference app (src/app.cpp:73-75)
Z_NEAR = 0.1
Z_FAR = 1000.0
FOVY = math.radians(45.0)

SEED_1M = 20240916
SEED_6M = 20240917


def perspective(fovy: float, aspect: float, z_near: float = Z_NEAR, z_far: float = Z_FAR) -> np.ndarray:
    """glm::perspective (RH, clip z in [-1,1]) as called at src/app.cpp:1042, float32 arithmetic."""
    f = np.float32
    t = f(math.tan(f(fovy) / f(2.0)))
    m = np.zeros((4, 4), np.float32)  # m[col, row]
    m[0, 0] = f(1.0) / (f(aspect) * t)
    m[1, 1] = f(1.0) / t
    m[2, 2] = -(f(z_far) + f(z_near)) / (f(z_far) - f(z_near))
    m[2, 3] = f(-1.0)
    m[3, 2] = -(f(2.0) * f(z_far) * f(z_near)) / (f(z_far) - f(z_near))
    return m.reshape(-1)


def translate(x: float, y: float, z: float) -> np.ndarray:
    m = np.eye(4, dtype=np.float32)
    m[3, :3] = (x, y, z)  # column 3
    return m.reshape(-1)


def look_at_camera(eye, target, up=(0.0, 1.0, 0.0)) -> np.ndarray:
    """camera-to-world matrix of a GL camera (looks down -Z, Y up) at `eye` looking at `target`."""
    eye, target, up = (np.asarray(v, np.float64) for v in (eye, target, up))
    fwd = target - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    upv = np.cross(right, fwd)
    m = np.eye(4, dtype=np.float32)
    m[0, :3] = right
    m[1, :3] = upv
    m[2, :3] = -fwd
    m[3, :3] = eye
    return m.reshape(-1)


def test_ply_camera() -> np.ndarray:
    """Camera the desktop app starts with for data/test.ply: data/test_vr.json floorMat lifted 1.5 along its
    Y axis (src/app.cpp:486-498); values from SURVEY.md 8(d) config 1."""
    return np.array([-0.705639, -1.64858e-05, 0.708571, 0.0,
                     -0.312374, 0.897588, -0.311061, 0.0,
                     -0.636, -0.440836, -0.633378, 0.0,
                     -1.041896, -0.442548, -0.978728, 1.0], np.float32)


test_ply_camera.__test__ = False  # not a pytest test


def synthetic_cloud_arrays(n: int, sh_degree: int, seed: int = SEED_1M, half_extent: float = 4.0) -> dict:
    """SURVEY 8(d) generator.  Draw order (all stored float32): pos ~ U(-h,h)^3, log_scale ~ U(ln .005, ln .05)^3,
    rot ~ N(0,1)^4 normalised (w,x,y,z), opacity logit ~ U(-2,4), f_dc ~ U(-1.5,1.5)^3, deg 3: f_rest ~ N(0,.05)^45."""
    rng = np.random.default_rng(seed)
    pos = rng.uniform(-half_extent, half_extent, (n, 3)).astype(np.float32)
    log_scale = rng.uniform(math.log(0.005), math.log(0.05), (n, 3)).astype(np.float32)
    rot = rng.standard_normal((n, 4))
    rot = (rot / np.linalg.norm(rot, axis=1, keepdims=True)).astype(np.float32)
    opacity = rng.uniform(-2.0, 4.0, (n, 1)).astype(np.float32)
    f_dc = rng.uniform(-1.5, 1.5, (n, 3)).astype(np.float32)
    out = {"pos": pos, "log_scale": log_scale, "rot": rot, "opacity": opacity, "f_dc": f_dc}
    if sh_degree >= 3:
        out["f_rest"] = (rng.standard_normal((n, 45)) * 0.05).astype(np.float32)
    return out


def ply_property_names(sh_degree: int) -> list[str]:
    names = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"]
    if sh_degree >= 3:
        names += [f"f_rest_{i}" for i in range(45)]
    names += ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    return names


def write_ply_from_arrays(path: str, a: dict) -> str:
    n = len(a["pos"])
    full = "f_rest" in a
    cols = [a["pos"], np.zeros((n, 3), np.float32), a["f_dc"]]
    if full:
        cols.append(a["f_rest"])
    cols += [a["opacity"], a["log_scale"], a["rot"]]
    body = np.ascontiguousarray(np.concatenate(cols, axis=1), dtype="<f4")
    names = ply_property_names(3 if full else 0)
    assert body.shape[1] == len(names)
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % n
    header += "".join(f"property float {p}\n" for p in names) + "end_header\n"
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(header.encode("ascii"))
        body.tofile(f)
    os.replace(tmp, path)
    return path


def write_synthetic_ply(path: str, n: int, sh_degree: int, seed: int = SEED_1M, half_extent: float = 4.0) -> str:
    """INRIA-layout PLY: 62 float props for degree 3, 17 (no f_rest) for degree 0 so HasFullSH()==false."""
    return write_ply_from_arrays(path, synthetic_cloud_arrays(n, sh_degree, seed, half_extent))


def ring_cameras(count: int = 8, radius: float = 18.0, height: float = 3.0) -> list[np.ndarray]:
    """config 5: cameras on a ring looking at the origin."""
    cams = []
    for k in range(count):
        ang = 2.0 * math.pi * k / count
        cams.append(look_at_camera((radius * math.sin(ang), height, radius * math.cos(ang)), (0.0, 0.0, 0.0)))
    return cams


def write_cameras_json(path: str, cams: list[np.ndarray], width: int = 1920, height: int = 1080, fx: float = 1303.7) -> str:
    """INRIA cameras.json (schema parsed at src/camerasconfig.cpp:31-56): rotation = camera-to-world with
    y down / z forward, i.e. the GL matrix with its Y and Z columns negated (:50-53)."""
    out = []
    for i, c in enumerate(cams):
        m = np.asarray(c, np.float64).reshape(4, 4)  # m[col, row]
        r = np.stack([m[0, :3], -m[1, :3], -m[2, :3]], axis=1)  # rows x cols
        out.append({"id": i, "img_name": f"synthetic_{i:03d}", "width": width, "height": height,
                    "position": [float(v) for v in m[3, :3]], "rotation": [[float(v) for v in row] for row in r],
                    "fx": fx, "fy": fx})
    with open(path, "w") as f:
        json.dump(out, f)
    return path


SPECIAL_FLOATS = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 1e-38, 3.4e38, -3.4e38, 88.7, -88.7, 89.0, -104.0, 16.6, -16.7, 1e-20, 0.5], np.float32)


def poison_arrays(a, rng) -> None:
    """parity-test helper: 15 % of every property replaced by special values (zeros, infinities, NaN, denormals, the edges of expf), 5 % by values
    of random magnitude 1e-30 .. 1e30"""
    for arr in a.values():
        m = rng.random(arr.shape) < 0.15
        arr[m] = rng.choice(SPECIAL_FLOATS, size=int(m.sum()))
        m2 = rng.random(arr.shape) < 0.05
        arr[m2] = (rng.standard_normal(int(m2.sum())) * 10.0 ** rng.uniform(-30, 30, int(m2.sum()))).astype(np.float32)
