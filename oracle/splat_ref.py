"""ctypes front-end of oracle/_ref/libsplat_ref.so: the reference's OWN host C++ and GLSL shader source, compiled
through the shims in oracle/ref_shim and driven by oracle/ref_harness.cpp.  TEST INFRASTRUCTURE ONLY -- used to pin
the oracle's restatement (tests/test_ref_pin.py); slow, single-threaded, small scenes."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build_ref
from .splat_oracle import RECORD_DTYPE

_lib = None


def available() -> bool:
    return build_ref.available() or os.path.exists(build_ref.LIB)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = build_ref.build()
        if path is None or not os.path.exists(path):
            raise FileNotFoundError("oracle/_ref/libsplat_ref.so is not built and /root/reference is absent")
        L = C.CDLL(path)
        L.ref_import_ply.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.c_void_p, C.c_size_t]
        L.ref_import_ply.restype = C.c_int
        L.ref_init_debug_cloud.argtypes = [C.c_void_p, C.c_size_t]
        L.ref_init_debug_cloud.restype = C.c_int
        L.ref_prune_splats.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_float), C.c_uint32, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t]
        L.ref_prune_splats.restype = C.c_int
        L.ref_export_ply.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p]
        L.ref_export_ply.restype = C.c_int
        L.ref_cameras_import.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.ref_cameras_import.restype = C.c_int
        fp = C.POINTER(C.c_float)
        L.ref_render.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, fp, fp, fp, fp, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_render.restype = C.c_size_t
        _lib = L
    return _lib


def import_ply(path: str, import_full_sh: bool = True):
    """GaussianCloud::ImportPly, executed by the reference's own gaussiancloud.cpp / ply.cpp."""
    n, stride, full = C.c_size_t(), C.c_size_t(), C.c_int()
    rc = lib().ref_import_ply(path.encode(), int(import_full_sh), C.byref(n), C.byref(stride), C.byref(full), None, 0)
    if rc:
        raise IOError(f"reference ImportPly returned false for {path}")
    aos = np.zeros((n.value, stride.value // 4), np.float32)
    rc = lib().ref_import_ply(path.encode(), int(import_full_sh), C.byref(n), C.byref(stride), C.byref(full), aos.ctypes.data, aos.size)
    assert rc == 0
    return aos, bool(full.value)


def prune_splats(path: str, origin, keep: int, import_full_sh: bool = True) -> np.ndarray:
    """GaussianCloud::PruneSplats of the reference on an imported file: the kept records in the reference's order"""
    o = (C.c_float * 3)(*[float(x) for x in origin])
    n, stride = C.c_size_t(), C.c_size_t()
    rc = lib().ref_prune_splats(path.encode(), int(import_full_sh), o, int(keep), C.byref(n), C.byref(stride), None, 0)
    if rc:
        raise IOError(f"reference ImportPly returned false for {path}")
    aos = np.zeros((n.value, stride.value // 4), np.float32)
    rc = lib().ref_prune_splats(path.encode(), int(import_full_sh), o, int(keep), C.byref(n), C.byref(stride), aos.ctypes.data, aos.size)
    assert rc == 0
    return aos


def export_ply(in_path: str, out_path: str, import_full_sh: bool = True, export_full_sh: bool = False) -> None:
    """GaussianCloud::ExportPly of the reference (its own Ply::Dump writes the file)"""
    rc = lib().ref_export_ply(in_path.encode(), int(import_full_sh), int(export_full_sh), out_path.encode())
    if rc:
        raise IOError(f"reference ExportPly failed ({rc}) for {in_path} -> {out_path}")


def json_is_real() -> bool:
    """True when the reference's camerasconfig.cpp was compiled against the real nlohmann/json (found in the image)"""
    L = lib()
    L.ref_json_is_real.restype = C.c_int
    return bool(L.ref_json_is_real())


ATTRIB_NAMES = ("posWithAlpha", "r_sh0", "r_sh1", "r_sh2", "r_sh3", "g_sh0", "g_sh1", "g_sh2", "g_sh3", "b_sh0", "b_sh1", "b_sh2", "b_sh3",
                "cov3_col0", "cov3_col1", "cov3_col2")


def attrib_table(path: str, import_full_sh: bool = True):
    """{name: (type, size, offset)} of the reference's InitAttribs after importing `path`, and its stride"""
    L = lib()
    L.ref_attrib_table.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
    L.ref_attrib_table.restype = C.c_int
    out = np.zeros(48, np.uint64)
    stride = L.ref_attrib_table(path.encode(), int(import_full_sh), out.ctypes.data)
    if stride < 0:
        raise IOError(f"reference ImportPly returned false for {path}")
    return {n: tuple(int(v) for v in out[i * 3:i * 3 + 3]) for i, n in enumerate(ATTRIB_NAMES)}, stride


def cameras_import_state(path: str):
    """(ImportJson's return value, cameras held afterwards) -- the reference keeps the entries read before a bad one"""
    L = lib()
    L.ref_cameras_import_state.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
    L.ref_cameras_import_state.restype = C.c_int
    ok = C.c_int()
    n = L.ref_cameras_import_state(path.encode(), C.byref(ok))
    return bool(ok.value), n


def cameras_import(path: str):
    """CamerasConfig::ImportJson + EstimateFloorPlane of the reference: None when ImportJson returns false, else
    (mats [n,16] column-major, fovs [n,2], floor normal [3], floor position [3])"""
    n = lib().ref_cameras_import(path.encode(), None, None, 0, None)
    if n < 0:
        return None
    mats, fovs, floor = np.zeros((max(n, 1), 16), np.float32), np.zeros((max(n, 1), 2), np.float32), np.zeros(6, np.float32)
    n2 = lib().ref_cameras_import(path.encode(), mats.ctypes.data, fovs.ctypes.data, max(n, 1), floor.ctypes.data)
    assert n2 == n
    return mats[:n], fovs[:n], floor[:3].copy(), floor[3:].copy()


def debug_cloud() -> np.ndarray:
    aos = np.zeros((16, 61), np.float32)
    n = lib().ref_init_debug_cloud(aos.ctypes.data, aos.size)
    assert n == 16
    return aos


def render(aos: np.ndarray, full_sh: bool, camera_mat, proj_mat, viewport, near_far, srgb: bool = False):
    """one frame through the reference's shader source: returns (fb [H,W,4] row 0 = bottom, records, draw order)"""
    aos = np.ascontiguousarray(aos, np.float32)
    n, nf = aos.shape
    f = lambda a, k: np.ascontiguousarray(np.asarray(a, np.float32).reshape(-1)[:k])
    cam, proj, vp, nfar = f(camera_mat, 16), f(proj_mat, 16), f(viewport, 4), f(near_far, 2)
    w, h = int(vp[2]), int(vp[3])
    fb = np.zeros((h, w, 4), np.float32)
    rec = np.zeros(max(n, 1), RECORD_DTYPE)
    order = np.zeros(max(n, 1), np.uint32)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    v = lib().ref_render(aos.ctypes.data, n, nf * 4, int(full_sh), int(srgb), p(cam), p(proj), p(vp), p(nfar), fb.ctypes.data, rec.ctypes.data, order.ctypes.data)
    return fb, rec[:n], order[:v].copy()
