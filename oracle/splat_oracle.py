"""ctypes front-end of the CPU oracle (oracle/splat_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module; nothing under splatapult_b200/ does.  Every function maps 1:1 onto a C entry point whose header
cites the reference file:line it restates.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

RECORD_DTYPE = np.dtype(
    [
        ("px", "<f4"), ("py", "<f4"),
        ("i00", "<f4"), ("i01", "<f4"), ("i10", "<f4"), ("i11", "<f4"),
        ("r", "<f4"), ("g", "<f4"), ("b", "<f4"),
        ("alpha", "<f4"),
        ("key", "<u4"), ("visible", "<u4"),
    ]
)


def build(force: bool = False) -> str:
    """Compile liboracle.so with oracle/Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "splat_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        fp = C.POINTER(C.c_float)
        L.orc_perspective.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, fp]
        L.orc_import_ply.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                     C.POINTER(C.c_int), C.c_void_p, C.c_size_t]
        L.orc_import_ply.restype = C.c_int
        L.orc_debug_cloud.argtypes = [C.c_void_p]
        L.orc_frame_constants.argtypes = [fp, fp, fp, fp, C.c_void_p]
        L.orc_project.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_project_shared.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_sort.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_sort.restype = C.c_size_t
        for f in (L.orc_composite, L.orc_composite_bruteforce):
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_render.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, fp, fp, fp, fp, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.orc_render.restype = C.c_size_t
        L.orc_sort_pairs_u64.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_record_size.restype = C.c_size_t
        L.orc_frame_size.restype = C.c_size_t
        assert L.orc_record_size() == RECORD_DTYPE.itemsize
        _lib = L
    return _lib


def _f(a) -> "C.POINTER(C.c_float)":
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f32(a, n) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))
    assert a.size == n, (a.size, n)
    return a


def perspective(fovy: float, aspect: float, z_near: float, z_far: float) -> np.ndarray:
    """glm::perspective as the reference calls it (app.cpp:1042); column-major float[16]."""
    out = np.zeros(16, np.float32)
    lib().orc_perspective(fovy, aspect, z_near, z_far, _f(out))
    return out


def import_ply(path: str, import_full_sh: bool = True):
    """GaussianCloud::ImportPly -> (aos float32 [n, 25|61], has_full_sh)."""
    n, stride, full = C.c_size_t(), C.c_size_t(), C.c_int()
    rc = lib().orc_import_ply(path.encode(), int(import_full_sh), C.byref(n), C.byref(stride), C.byref(full), None, 0)
    if rc:
        raise IOError(f"oracle ply parse failed rc={rc}: {path}")
    nf = stride.value // 4
    aos = np.zeros((n.value, nf), np.float32)
    rc = lib().orc_import_ply(path.encode(), int(import_full_sh), C.byref(n), C.byref(stride), C.byref(full),
                              aos.ctypes.data, aos.size)
    if rc:
        raise IOError(f"oracle ply decode failed rc={rc}: {path}")
    return aos, bool(full.value)


def debug_cloud() -> np.ndarray:
    """GaussianCloud::InitDebugCloud -> aos float32 [16, 61] (stride 244, hasFullSH == false)."""
    aos = np.zeros((16, 61), np.float32)
    lib().orc_debug_cloud(aos.ctypes.data)
    return aos


class Frame:
    """per-frame constants: view=inverse(cameraMat), mvp=proj*view, eye, W,H,near,far."""

    def __init__(self, camera_mat, proj_mat, viewport, near_far):
        self.camera = _f32(camera_mat, 16)
        self.proj = _f32(proj_mat, 16)
        self.viewport = _f32(viewport, 4)
        self.near_far = _f32(near_far, 2)
        self.buf = np.zeros(lib().orc_frame_size() // 4, np.float32)
        lib().orc_frame_constants(_f(self.camera), _f(self.proj), _f(self.viewport), _f(self.near_far), self.buf.ctypes.data)

    @property
    def view(self):
        return self.buf[0:16]

    @property
    def mvp(self):
        return self.buf[16:32]

    @property
    def eye(self):
        return self.buf[48:51]

    @property
    def width(self):
        return int(self.viewport[2])

    @property
    def height(self):
        return int(self.viewport[3])


def project(aos: np.ndarray, full_sh: bool, frame: Frame, srgb: bool = False) -> np.ndarray:
    aos = np.ascontiguousarray(aos, np.float32)
    n, nf = aos.shape
    rec = np.zeros(n, RECORD_DTYPE)
    lib().orc_project(aos.ctypes.data, n, nf * 4, int(full_sh), int(srgb), frame.buf.ctypes.data, rec.ctypes.data)
    return rec


def project_shared(aos: np.ndarray, full_sh: bool, sort_frame: Frame, draw_frame: Frame, srgb: bool = False) -> np.ndarray:
    """records of a draw that reuses another view's Sort (src/app.cpp:603-607): cull + key from sort_frame, the rest from draw_frame."""
    aos = np.ascontiguousarray(aos, np.float32)
    n, nf = aos.shape
    rec = np.zeros(n, RECORD_DTYPE)
    lib().orc_project_shared(aos.ctypes.data, n, nf * 4, int(full_sh), int(srgb), sort_frame.buf.ctypes.data,
                             draw_frame.buf.ctypes.data, rec.ctypes.data)
    return rec


def sort(rec: np.ndarray) -> np.ndarray:
    """back-to-front order (ascending key, ties ascending index) of the visible gaussians."""
    order = np.zeros(max(len(rec), 1), np.uint32)
    v = lib().orc_sort(rec.ctypes.data, len(rec), order.ctypes.data)
    return order[:v].copy()


def composite(rec: np.ndarray, order: np.ndarray, width: int, height: int, fp16: bool = False,
              bruteforce: bool = False) -> np.ndarray:
    fb = np.zeros((height, width, 4), np.float32)
    order = np.ascontiguousarray(order, np.uint32)
    fn = lib().orc_composite_bruteforce if bruteforce else lib().orc_composite
    fn(rec.ctypes.data, order.ctypes.data, len(order), width, height, int(fp16), fb.ctypes.data)
    return fb


def render(aos: np.ndarray, full_sh: bool, camera_mat, proj_mat, viewport, near_far, fp16: bool = False,
           srgb: bool = False, want_records: bool = False):
    """SplatRenderer::Sort + Render into a cleared float target. Returns fb[H,W,4] (row 0 = bottom),
    and optionally (records, order, stage_ms)."""
    aos = np.ascontiguousarray(aos, np.float32)
    n, nf = aos.shape
    cam, proj, vp, nfar = _f32(camera_mat, 16), _f32(proj_mat, 16), _f32(viewport, 4), _f32(near_far, 2)
    w, h = int(vp[2]), int(vp[3])
    fb = np.zeros((h, w, 4), np.float32)
    rec = np.zeros(max(n, 1), RECORD_DTYPE)
    order = np.zeros(max(n, 1), np.uint32)
    ms = (C.c_double * 3)()
    v = lib().orc_render(aos.ctypes.data, n, nf * 4, int(full_sh), int(srgb), _f(cam), _f(proj), _f(vp), _f(nfar),
                         int(fp16), fb.ctypes.data, rec.ctypes.data, order.ctypes.data, ms)
    if want_records:
        return fb, rec[:n], order[:v].copy(), [ms[0], ms[1], ms[2]]
    return fb


def sort_pairs_u64(keys: np.ndarray, vals: np.ndarray, begin_bit: int = 0, end_bit: int = 64):
    keys = np.ascontiguousarray(keys, np.uint64)
    vals = np.ascontiguousarray(vals, np.uint32)
    ko, vo = np.empty_like(keys), np.empty_like(vals)
    lib().orc_sort_pairs_u64(keys.ctypes.data, vals.ctypes.data, len(keys), begin_bit, end_bit, ko.ctypes.data, vo.ctypes.data)
    return ko, vo


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(n)
