"""ctypes binding of include/splat_b200.h, shaped like the reference's C++ classes."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsplat_b200.so")

FB_FP32 = 0
FB_FP16 = 1

OPT_EARLY_OUT_T = 0
OPT_PROFILE = 1
OPT_INSTANCES_PER_GAUSSIAN = 2
OPT_SORT_MODE = 3
OPT_SORT_TILE = 4
OPT_SUBBLOCK_CULL = 5
OPT_SHARED_SORT = 6
OPT_PDL = 7
OPT_GRAPH = 8
SORT_BINNED = 0
SORT_LSD = 1
SORT_LSD_FLAT = 2

TAP_RECORDS = 0
TAP_KEYS = 1
TAP_VALS = 2
TAP_TILE_RANGES = 3

ATTR_NAMES = ["posWithAlpha", "r_sh0", "r_sh1", "r_sh2", "r_sh3", "g_sh0", "g_sh1", "g_sh2", "g_sh3",
              "b_sh0", "b_sh1", "b_sh2", "b_sh3", "cov3_col0", "cov3_col1", "cov3_col2"]

# bit-compatible with oracle.splat_oracle.RECORD_DTYPE (declared independently: the product never imports the oracle)
RECORD_DTYPE = np.dtype([("px", "<f4"), ("py", "<f4"), ("i00", "<f4"), ("i01", "<f4"), ("i10", "<f4"), ("i11", "<f4"),
                         ("r", "<f4"), ("g", "<f4"), ("b", "<f4"), ("alpha", "<f4"), ("key", "<u4"), ("visible", "<u4")])


class SplatError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [
        ("num_gaussians", C.c_uint64), ("num_visible", C.c_uint64), ("num_instances", C.c_uint64),
        ("instance_capacity", C.c_uint64), ("overflow", C.c_uint32),
        ("tiles_x", C.c_uint32), ("tiles_y", C.c_uint32), ("tile_w", C.c_uint32), ("tile_h", C.c_uint32),
        ("sort_passes", C.c_uint32), ("sort_bits", C.c_uint32), ("sh_mode", C.c_uint32), ("sort_mode", C.c_uint32),
        ("kernel_launches", C.c_uint32), ("stage_ms", C.c_float * 8),
    ]

    def as_dict(self) -> dict:
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "stage_ms"}
        d["stage_ms"] = [float(v) for v in self.stage_ms]
        return d


# every symbol include/splat_b200.h declares: (name, restype, argtypes)
_FP = C.POINTER(C.c_float)
_SIGNATURES = [
    ("spl_last_error", C.c_char_p, []),
    ("spl_version", C.c_int, []),
    ("spl_device_count", C.c_int, []),
    ("spl_cloud_create", C.c_void_p, [C.c_int, C.c_int]),
    ("spl_cloud_destroy", None, [C.c_void_p]),
    ("spl_cloud_import_ply", C.c_int, [C.c_void_p, C.c_char_p]),
    ("spl_cloud_import_ply_gpu", C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    ("spl_renderer_init_ply", C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int]),
    ("spl_cloud_export_ply", C.c_int, [C.c_void_p, C.c_char_p]),
    ("spl_cloud_init_debug_cloud", C.c_int, [C.c_void_p]),
    ("spl_cloud_prune_splats", C.c_int, [C.c_void_p, _FP, C.c_uint32]),
    ("spl_cloud_set_raw", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int]),
    ("spl_cloud_num_gaussians", C.c_size_t, [C.c_void_p]),
    ("spl_cloud_stride", C.c_size_t, [C.c_void_p]),
    ("spl_cloud_total_size", C.c_size_t, [C.c_void_p]),
    ("spl_cloud_raw_data", C.c_void_p, [C.c_void_p]),
    ("spl_cloud_has_full_sh", C.c_int, [C.c_void_p]),
    ("spl_cloud_attrib", C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    ("spl_renderer_create", C.c_void_p, []),
    ("spl_renderer_destroy", None, [C.c_void_p]),
    ("spl_renderer_configure", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    ("spl_renderer_init", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    ("spl_renderer_init_device_aos", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int]),
    ("spl_renderer_sort", C.c_int, [C.c_void_p, _FP, _FP, _FP, _FP]),
    ("spl_renderer_render", C.c_int, [C.c_void_p, _FP, _FP, _FP, _FP]),
    ("spl_renderer_framebuffer", C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                           C.POINTER(C.c_size_t), C.POINTER(C.c_int)]),
    ("spl_renderer_read_framebuffer_async", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    ("spl_renderer_read_framebuffer", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    ("spl_renderer_read_rgba8_async", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    ("spl_renderer_read_rgba8", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    ("spl_renderer_fence_copies", C.c_int, [C.c_void_p]),
    ("spl_renderer_wait_copies", C.c_int, [C.c_void_p, C.c_int]),
    ("spl_renderer_synchronize", C.c_int, [C.c_void_p]),
    ("spl_renderer_stats", C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    ("spl_renderer_reserve", C.c_int, [C.c_void_p, C.c_uint64]),
    ("spl_renderer_set_option", C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    ("spl_renderer_tap", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    ("spl_cameras_create", C.c_void_p, []),
    ("spl_cameras_destroy", None, [C.c_void_p]),
    ("spl_cameras_import_json", C.c_int, [C.c_void_p, C.c_char_p]),
    ("spl_cameras_count", C.c_size_t, [C.c_void_p]),
    ("spl_cameras_get", C.c_int, [C.c_void_p, C.c_size_t, _FP, _FP]),
    ("spl_cameras_estimate_floor_plane", C.c_int, [C.c_void_p, _FP, _FP]),
    ("spl_sorter_create", C.c_void_p, [C.c_int, C.c_uint64]),
    ("spl_sorter_destroy", None, [C.c_void_p]),
    ("spl_sorter_sort", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p]),
    ("spl_sort_pairs_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    ("spl_write_ppm", C.c_int, [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    ("spl_host_alloc", C.c_void_p, [C.c_size_t]),
    ("spl_host_alloc_near", C.c_void_p, [C.c_size_t, C.c_int]),
    ("spl_host_free", None, [C.c_void_p]),
    ("spl_device_numa_node", C.c_int, [C.c_int]),
    ("spl_bind_thread_near", C.c_int, [C.c_int]),
    ("spl_comm_get_unique_id", C.c_int, [C.c_char_p]),
    ("spl_comm_create", C.c_void_p, [C.c_char_p, C.c_int, C.c_int, C.c_int]),
    ("spl_comm_destroy", None, [C.c_void_p]),
    ("spl_comm_nccl_version", C.c_int, []),
    ("spl_renderer_init_broadcast", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, _FP]),
]
COMM_ID_BYTES = 128
EXPORTED_SYMBOLS = [s[0] for s in _SIGNATURES]

_lib = None


def library_path() -> str:
    return _LIB_PATH


def lib() -> C.CDLL:
    """Load libsplat_b200.so.  Fails loudly if it has not been built: there is no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError(f"{_LIB_PATH} is missing: run `python -m splatapult_b200.build` (needs nvcc). "
                              "splatapult_b200 has no CPU or pure-Python fallback.")
        L = C.CDLL(_LIB_PATH)
        for name, res, args in _SIGNATURES:
            fn = getattr(L, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise SplatError(rc, (lib().spl_last_error() or b"").decode(errors="replace"))


def device_count() -> int:
    return lib().spl_device_count()


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(_FP)


def _f32(a, n: int) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))
    if a.size != n:
        raise ValueError(f"expected {n} floats, got {a.size}")
    return a


class GaussianCloud:
    """Mirror of the reference's GaussianCloud (src/gaussiancloud.h:17-91)."""

    def __init__(self, import_full_sh: bool = True, export_full_sh: bool = True):
        self._h = lib().spl_cloud_create(int(import_full_sh), int(export_full_sh))
        if not self._h:
            raise MemoryError("spl_cloud_create failed")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.spl_cloud_destroy(h)

    # --- reference methods
    def ImportPly(self, ply_filename: str) -> bool:
        return lib().spl_cloud_import_ply(self._h, os.fsencode(ply_filename)) == 0

    def ImportPlyGpu(self, ply_filename: str, device: int = 0) -> bool:
        """ImportPly with the per-vertex decode run as a CUDA kernel (same records bit for bit); False without a GPU."""
        return lib().spl_cloud_import_ply_gpu(self._h, os.fsencode(ply_filename), int(device)) == 0

    def ExportPly(self, ply_filename: str) -> bool:
        return lib().spl_cloud_export_ply(self._h, os.fsencode(ply_filename)) == 0

    def InitDebugCloud(self) -> None:
        _check(lib().spl_cloud_init_debug_cloud(self._h))

    def PruneSplats(self, origin, num_gaussians: int) -> None:
        o = _f32(origin, 3)
        _check(lib().spl_cloud_prune_splats(self._h, _fptr(o), int(num_gaussians)))

    def GetNumGaussians(self) -> int:
        return lib().spl_cloud_num_gaussians(self._h)

    def GetStride(self) -> int:
        return lib().spl_cloud_stride(self._h)

    def GetTotalSize(self) -> int:
        return lib().spl_cloud_total_size(self._h)

    def HasFullSH(self) -> bool:
        return bool(lib().spl_cloud_has_full_sh(self._h))

    def GetRawData(self) -> np.ndarray:
        """zero-copy float32 view [N, stride/4] of GetRawDataPtr()."""
        n, stride = self.GetNumGaussians(), self.GetStride()
        if n == 0:
            return np.zeros((0, max(stride // 4, 1)), np.float32)
        ptr = lib().spl_cloud_raw_data(self._h)
        buf = (C.c_float * (n * stride // 4)).from_address(ptr)
        arr = np.frombuffer(buf, dtype=np.float32).reshape(n, stride // 4)
        arr.flags.writeable = True
        self._keep = buf
        return arr

    def GetAttrib(self, name: str):
        """(type, size, offset) of one of the 16 BinaryAttribute accessors, e.g. 'r_sh1'."""
        t, s, o = C.c_int(), C.c_size_t(), C.c_size_t()
        _check(lib().spl_cloud_attrib(self._h, ATTR_NAMES.index(name), C.byref(t), C.byref(s), C.byref(o)))
        return t.value, s.value, o.value

    def LastError(self) -> str:
        return (lib().spl_last_error() or b"").decode(errors="replace")

    # --- extension: adopt an AoS built elsewhere
    def SetRawData(self, aos: np.ndarray, has_full_sh: bool) -> None:
        aos = np.ascontiguousarray(aos, np.float32)
        n, nf = aos.shape
        _check(lib().spl_cloud_set_raw(self._h, aos.ctypes.data, n, nf * 4, int(has_full_sh)))


class SplatRenderer:
    """Mirror of the reference's SplatRenderer (src/splatrenderer.h:23-67): Init, then Sort + Render per frame."""

    def __init__(self, device: int = 0, fb_format: int = FB_FP32, cuda_stream: int | None = None):
        self._h = lib().spl_renderer_create()
        if not self._h:
            raise MemoryError("spl_renderer_create failed")
        self.fb_format = fb_format
        self.numBlocksPerWorkgroup = 1024  # source compatibility, unused (src/splatrenderer.h:39)
        _check(lib().spl_renderer_configure(self._h, device, fb_format, C.c_void_p(cuda_stream or 0)))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.spl_renderer_destroy(h)

    def close(self) -> None:
        self.__del__()

    # --- reference methods
    def Init(self, gaussian_cloud: GaussianCloud, is_framebuffer_srgb_enabled: bool = False, use_rgc_sort_override: bool = False) -> bool:
        rc = lib().spl_renderer_init(self._h, gaussian_cloud._h, int(is_framebuffer_srgb_enabled), int(use_rgc_sort_override))
        if rc:
            self.last_error = (lib().spl_last_error() or b"").decode(errors="replace")
        return rc == 0

    def Sort(self, camera_mat, proj_mat, viewport, near_far) -> None:
        c, p, v, n = _f32(camera_mat, 16), _f32(proj_mat, 16), _f32(viewport, 4), _f32(near_far, 2)
        _check(lib().spl_renderer_sort(self._h, _fptr(c), _fptr(p), _fptr(v), _fptr(n)))

    def Render(self, camera_mat, proj_mat, viewport, near_far) -> None:
        c, p, v, n = _f32(camera_mat, 16), _f32(proj_mat, 16), _f32(viewport, 4), _f32(near_far, 2)
        _check(lib().spl_renderer_render(self._h, _fptr(c), _fptr(p), _fptr(v), _fptr(n)))

    # --- explicit replacements for implicit GL state
    def InitFromDevice(self, device_ptr: int, num_gaussians: int, stride: int, has_full_sh: bool, srgb: bool = False) -> None:
        _check(lib().spl_renderer_init_device_aos(self._h, C.c_void_p(device_ptr), num_gaussians, stride, int(has_full_sh), int(srgb)))

    def InitBroadcast(self, comm: "SplatComm", cloud_on_root: "GaussianCloud | None", root: int = 0, srgb: bool = False) -> float:
        """collective: every rank calls it; `root` passes its imported cloud, the others None.  One ncclBroadcast of the
        AoS inside the library (no torch tensor on the path).  Returns the device time of that broadcast in ms."""
        ms = C.c_float(0.0)
        _check(lib().spl_renderer_init_broadcast(self._h, comm._h, cloud_on_root._h if cloud_on_root is not None else None, int(root), int(srgb),
                                                 C.byref(ms)))
        return float(ms.value)

    def InitFromPly(self, ply_filename: str, import_full_sh: bool = True, srgb: bool = False) -> None:
        """ImportPly + Init without the host decode: the PLY body is decoded on the device and packed there."""
        _check(lib().spl_renderer_init_ply(self._h, os.fsencode(ply_filename), int(import_full_sh), int(srgb)))

    def Framebuffer(self):
        """(device_ptr, width, height, row_pitch_bytes, fb_format) of the last Render."""
        p, w, h, pitch, fmt = C.c_void_p(), C.c_int(), C.c_int(), C.c_size_t(), C.c_int()
        _check(lib().spl_renderer_framebuffer(self._h, C.byref(p), C.byref(w), C.byref(h), C.byref(pitch), C.byref(fmt)))
        return p.value, w.value, h.value, pitch.value, fmt.value

    def ReadFramebuffer(self) -> np.ndarray:
        """float32 [H, W, 4] (fp16 targets are widened); row 0 = bottom row, like glReadPixels."""
        _, w, h, _, fmt = self.Framebuffer()
        if fmt == FB_FP16:
            raw = np.empty((h, w, 4), np.float16)
        else:
            raw = np.empty((h, w, 4), np.float32)
        _check(lib().spl_renderer_read_framebuffer(self._h, raw.ctypes.data, raw.nbytes))
        return raw.astype(np.float32) if fmt == FB_FP16 else raw

    def ReadFramebufferAsync(self, host_ptr: int, nbytes: int) -> None:
        _check(lib().spl_renderer_read_framebuffer_async(self._h, C.c_void_p(host_ptr), nbytes))

    def ReadPresented8(self) -> np.ndarray:
        """uint8 [H, W, 4]: the presented frame (the reference's RenderDesktop blit into an 8-bit window); row 0 = bottom."""
        _, w, h, _, _ = self.Framebuffer()
        raw = np.empty((h, w, 4), np.uint8)
        _check(lib().spl_renderer_read_rgba8(self._h, raw.ctypes.data, raw.nbytes))
        return raw

    def ReadPresented8Async(self, host_ptr: int, nbytes: int) -> None:
        _check(lib().spl_renderer_read_rgba8_async(self._h, C.c_void_p(host_ptr), nbytes))

    def FenceCopies(self) -> None:
        """device-side join: the render stream waits for every framebuffer copy still in flight."""
        _check(lib().spl_renderer_fence_copies(self._h))

    def WaitCopies(self, keep_in_flight: int = 0) -> None:
        """host-side wait until at most keep_in_flight (0 or 1) asynchronous framebuffer copies are pending."""
        _check(lib().spl_renderer_wait_copies(self._h, int(keep_in_flight)))

    def Synchronize(self) -> None:
        _check(lib().spl_renderer_synchronize(self._h))

    def GetStats(self) -> Stats:
        s = Stats()
        _check(lib().spl_renderer_stats(self._h, C.byref(s)))
        return s

    def Reserve(self, instance_capacity: int) -> None:
        _check(lib().spl_renderer_reserve(self._h, int(instance_capacity)))

    def SetOption(self, option: int, value: float) -> None:
        _check(lib().spl_renderer_set_option(self._h, option, float(value)))

    def Tap(self, what: int) -> np.ndarray:
        st = self.GetStats()
        if what == TAP_RECORDS:
            out = np.zeros(st.num_gaussians, RECORD_DTYPE)
        elif what == TAP_KEYS:
            out = np.zeros(min(st.num_instances, st.instance_capacity), np.uint64)
        elif what == TAP_VALS:
            out = np.zeros(min(st.num_instances, st.instance_capacity), np.uint32)
        elif what == TAP_TILE_RANGES:
            out = np.zeros((st.tiles_y * st.tiles_x, 2), np.uint32)
        else:
            raise ValueError(what)
        if out.nbytes:
            _check(lib().spl_renderer_tap(self._h, what, out.ctypes.data, out.nbytes))
        return out


_nccl_preloaded = False


def _preload_nccl() -> None:
    """libsplat_b200.so binds NCCL at run time by SONAME (libnccl.so.2).  A Python process may import torch AFTER that, and
    torch's libtorch_cuda.so needs the libnccl.so.2 it was built with (the pip package nvidia-nccl, same SONAME as a system
    copy): if that package exists, load it first, so that one copy serves both.  A C++ host simply gets the system's."""
    global _nccl_preloaded
    if _nccl_preloaded:
        return
    _nccl_preloaded = True
    try:
        import importlib.util

        spec = importlib.util.find_spec("nvidia")
        for base in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(base, "nccl", "lib", "libnccl.so.2")
            if os.path.exists(cand):
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                return
    except Exception:  # noqa: BLE001 -- best effort: the library's own dlopen still finds a system copy
        pass


class SplatComm:
    """NCCL communicator of the multi-GPU start-up (include/splatapult_b200/splatcomm.h): one rank per GPU."""

    def __init__(self, unique_id: bytes, rank: int, nranks: int, device: int):
        if len(unique_id) != COMM_ID_BYTES:
            raise ValueError("unique id must be 128 bytes")
        _preload_nccl()
        self._h = lib().spl_comm_create(unique_id, int(rank), int(nranks), int(device))
        if not self._h:
            raise SplatError(-3, (lib().spl_last_error() or b"").decode(errors="replace"))
        self.rank, self.nranks, self.device = rank, nranks, device

    @staticmethod
    def GetUniqueId() -> bytes:
        _preload_nccl()
        buf = C.create_string_buffer(COMM_ID_BYTES)
        _check(lib().spl_comm_get_unique_id(buf))
        return buf.raw

    @staticmethod
    def NcclVersion() -> int:
        _preload_nccl()
        return lib().spl_comm_nccl_version()

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.spl_comm_destroy(h)

    def close(self) -> None:
        self.__del__()


class CamerasConfig:
    """Mirror of the reference's CamerasConfig (src/camerasconfig.h:12-32): INRIA cameras.json -> cameraMat list."""

    def __init__(self):
        self._h = lib().spl_cameras_create()
        if not self._h:
            raise MemoryError("spl_cameras_create failed")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.spl_cameras_destroy(h)

    def ImportJson(self, json_filename: str) -> bool:
        return lib().spl_cameras_import_json(self._h, os.fsencode(json_filename)) == 0

    def GetNumCameras(self) -> int:
        return lib().spl_cameras_count(self._h)

    def GetCamera(self, index: int):
        """(mat float32[16] column-major camera-to-world, fov float32[2]) of GetCameraVec()[index]."""
        m, f = np.zeros(16, np.float32), np.zeros(2, np.float32)
        _check(lib().spl_cameras_get(self._h, index, _fptr(m), _fptr(f)))
        return m, f

    def EstimateFloorPlane(self):
        n, p = np.zeros(3, np.float32), np.zeros(3, np.float32)
        _check(lib().spl_cameras_estimate_floor_plane(self._h, _fptr(n), _fptr(p)))
        return n, p

    def LastError(self) -> str:
        return (lib().spl_last_error() or b"").decode(errors="replace")


def sort_pairs_host(keys: np.ndarray, vals: np.ndarray, begin_bit: int = 0, end_bit: int = 64):
    """stable ascending radix sort of (u64, u32) pairs on the GPU (host arrays in/out)."""
    keys = np.ascontiguousarray(keys, np.uint64)
    vals = np.ascontiguousarray(vals, np.uint32)
    ko, vo = np.empty_like(keys), np.empty_like(vals)
    _check(lib().spl_sort_pairs_host(keys.ctypes.data, vals.ctypes.data, len(keys), begin_bit, end_bit, ko.ctypes.data, vo.ctypes.data))
    return ko, vo


def write_ppm(path: str, fb: np.ndarray) -> None:
    """8-bit binary PPM of a framebuffer [H, W, 4] float32 / float16 (row 0 = bottom); top row first, clamped to [0,1]."""
    fb = np.ascontiguousarray(fb)
    if fb.ndim != 3 or fb.shape[2] != 4 or fb.dtype not in (np.float32, np.float16):
        raise ValueError("expected [H, W, 4] float32 or float16")
    _check(lib().spl_write_ppm(os.fsencode(path), fb.ctypes.data, fb.shape[1], fb.shape[0], FB_FP16 if fb.dtype == np.float16 else FB_FP32))


class _Pinned:
    def __init__(self, nbytes: int, near_device: int | None = None):
        self.ptr = lib().spl_host_alloc(nbytes) if near_device is None else lib().spl_host_alloc_near(nbytes, int(near_device))
        if not self.ptr:
            raise MemoryError("spl_host_alloc failed")
        self.nbytes = nbytes

    def array(self, dtype, shape) -> np.ndarray:
        buf = (C.c_uint8 * self.nbytes).from_address(self.ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def __del__(self):
        p, self.ptr = getattr(self, "ptr", None), None
        if p and _lib is not None:
            _lib.spl_host_free(C.c_void_p(p))


def host_alloc(nbytes: int, near_device: int | None = None) -> _Pinned:
    """pinned host memory for the end-to-end path; near_device = d puts its pages on GPU d's NUMA node."""
    return _Pinned(nbytes, near_device)


def bind_thread_near(device: int) -> int:
    """pin the calling thread to the CPUs of GPU `device`'s NUMA node; returns the node or -1 (nothing changed)."""
    return lib().spl_bind_thread_near(int(device))


def device_numa_node(device: int) -> int:
    return lib().spl_device_numa_node(int(device))
