"""splatapult_b200 -- B200-native drop-in for splatapult's splat hot path.

Python here is plumbing only: a ctypes binding of the C ABI (include/splat_b200.h) with classes named after the
reference's C++ classes (GaussianCloud: src/gaussiancloud.h:17-91, SplatRenderer: src/splatrenderer.h:23-67) so
the parity tests read like calls into the reference.  All compute is in libsplat_b200.so (CUDA, sm_100a); there
is no CPU fallback -- without the library or without a GPU the renderer raises.
"""
from .api import (  # noqa: F401
    CamerasConfig,
    FB_FP16,
    FB_FP32,
    GaussianCloud,
    SplatComm,
    SplatError,
    SplatRenderer,
    Stats,
    bind_thread_near,
    device_count,
    device_numa_node,
    host_alloc,
    lib,
    library_path,
    sort_pairs_host,
    write_ppm,
)
from .scenes import (  # noqa: F401
    FOVY,
    Z_FAR,
    Z_NEAR,
    look_at_camera,
    perspective,
    ring_cameras,
    synthetic_cloud_arrays,
    test_ply_camera,
    translate,
    write_cameras_json,
    write_synthetic_ply,
)

__all__ = [n for n in dir() if not n.startswith("_")]
