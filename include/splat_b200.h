/*
 * splat_b200.h -- C ABI of libsplat_b200.so: the drop-in boundary for splatapult's splat hot path.
 *
 * The reference (hyperlogic/splatapult) has no plugin/FFI layer: the boundary is two C++ classes,
 *   GaussianCloud  src/gaussiancloud.h:17-91
 *   SplatRenderer  src/splatrenderer.h:23-67
 * called from App::Init (src/app.cpp:549-559) and App::Render (src/app.cpp:1067-1068).  Each entry point
 * below names the reference member it stands in for.  The C++ classes with the reference's own method
 * names live in include/splatapult_b200/{gaussiancloud,splatrenderer}.h and are thin owners of these
 * handles' implementation; INTEGRATION.md shows the patch a splatapult maintainer would apply.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types.
 *   - every int-returning call: 0 = ok, negative = error (spl_last_error() holds the message, per thread).
 *   - matrices are glm-layout column-major float[16]; cameraMat is camera-to-world (the renderer inverts it,
 *     src/splatrenderer.cpp:161,327); viewport = (x, y, width, height) with x = y = 0;
 *     nearFar must equal the near/far baked into projMat.
 *   - sort/render are asynchronous on the renderer's CUDA stream; there is no host sync on the frame path
 *     (the reference stalls on a counter read-back, src/splatrenderer.cpp:195-204).
 *   - the framebuffer is GL-oriented: row 0 is the BOTTOM row, pixel centres at (i+0.5, j+0.5), RGBA,
 *     alpha == 1 (clear colour (0,0,0,1) + "over" blending, src/app.cpp:153-160).
 *   - there is NO CPU fallback: without a CUDA device every renderer call fails with SPL_ERR_CUDA.
 */
#ifndef SPLAT_B200_H
#define SPLAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SPL_API __declspec(dllexport)
#else
#define SPL_API __attribute__((visibility("default")))
#endif

typedef struct spl_cloud spl_cloud;       /* GaussianCloud  */
typedef struct spl_renderer spl_renderer; /* SplatRenderer  */
typedef struct spl_sorter spl_sorter;     /* stand-alone (tile|depth) u64 radix sorter */

enum {
    SPL_OK = 0,
    SPL_ERR_ARG = -1,      /* bad argument / state */
    SPL_ERR_IO = -2,       /* file open / parse failure (GaussianCloud::ImportPly returns false) */
    SPL_ERR_CUDA = -3,     /* CUDA runtime error or no device */
    SPL_ERR_OVERFLOW = -4, /* tile-instance capacity exceeded (see spl_renderer_stats) */
    SPL_ERR_ALLOC = -5
};

enum { SPL_FB_FP32 = 0 /* GL_RGBA32F, --fp32 */, SPL_FB_FP16 = 1 /* GL_RGBA16F, --fp16 */ }; /* src/app.cpp:1000-1035 */

/* attribute ids for spl_cloud_attrib: the 16 BinaryAttribute accessors, src/gaussiancloud.h:42-57 */
enum {
    SPL_ATTR_POS_WITH_ALPHA = 0,
    SPL_ATTR_R_SH0, SPL_ATTR_R_SH1, SPL_ATTR_R_SH2, SPL_ATTR_R_SH3,
    SPL_ATTR_G_SH0, SPL_ATTR_G_SH1, SPL_ATTR_G_SH2, SPL_ATTR_G_SH3,
    SPL_ATTR_B_SH0, SPL_ATTR_B_SH1, SPL_ATTR_B_SH2, SPL_ATTR_B_SH3,
    SPL_ATTR_COV3_COL0, SPL_ATTR_COV3_COL1, SPL_ATTR_COV3_COL2,
    SPL_ATTR_COUNT
};

SPL_API const char* spl_last_error(void);
SPL_API int spl_version(void);
SPL_API int spl_device_count(void); /* 0 when no CUDA device is usable */

/* ---------------------------------------------------------------- GaussianCloud (host, src/gaussiancloud.h) */
SPL_API spl_cloud* spl_cloud_create(int import_full_sh, int export_full_sh);           /* GaussianCloud(Options) :26 */
SPL_API void spl_cloud_destroy(spl_cloud*);
SPL_API int spl_cloud_import_ply(spl_cloud*, const char* ply_filename);                /* ImportPly :28 */
/* ImportPly with the per-vertex decode (sigmoid / exp / quaternion -> Sigma / SH repack, gaussiancloud.cpp:254-361)
 * run as a CUDA kernel on `device`: same records bit for bit; SPL_ERR_CUDA without a device (no CPU fallback) */
SPL_API int spl_cloud_import_ply_gpu(spl_cloud*, const char* ply_filename, int device);
SPL_API int spl_cloud_export_ply(const spl_cloud*, const char* ply_filename);          /* ExportPly :29 */
SPL_API int spl_cloud_init_debug_cloud(spl_cloud*);                                    /* InitDebugCloud :31 */
SPL_API int spl_cloud_prune_splats(spl_cloud*, const float origin[3], uint32_t count); /* PruneSplats :34 */
/* adopt a copy of a caller-built AoS in the reference layout (stride 100 or 244 bytes) */
SPL_API int spl_cloud_set_raw(spl_cloud*, const void* aos, size_t num_gaussians, size_t stride, int has_full_sh);
SPL_API size_t spl_cloud_num_gaussians(const spl_cloud*);                              /* GetNumGaussians :36 */
SPL_API size_t spl_cloud_stride(const spl_cloud*);                                     /* GetStride :37 */
SPL_API size_t spl_cloud_total_size(const spl_cloud*);                                 /* GetTotalSize :38 */
SPL_API void* spl_cloud_raw_data(spl_cloud*);                                          /* GetRawDataPtr :39 */
SPL_API int spl_cloud_has_full_sh(const spl_cloud*);                                   /* HasFullSH :62 */
/* BinaryAttribute {type,size,offset}; type 0 = Unknown, 7 = Float (src/core/binaryattribute.h:15-27) */
SPL_API int spl_cloud_attrib(const spl_cloud*, int attr_id, int* type, size_t* size, size_t* offset);

/* ---------------------------------------------------------------- SplatRenderer (src/splatrenderer.h) */
SPL_API spl_renderer* spl_renderer_create(void);                                       /* SplatRenderer() :26 */
SPL_API void spl_renderer_destroy(spl_renderer*);
/* optional, before init: CUDA device ordinal, framebuffer format, and an external cudaStream_t to run on
 * (NULL = the renderer creates its own non-blocking stream). */
SPL_API int spl_renderer_configure(spl_renderer*, int device, int fb_format, void* cuda_stream);
/* Init :29-30.  Copies the cloud to the device and packs it; keeps no reference to host memory. */
SPL_API int spl_renderer_init(spl_renderer*, const spl_cloud*, int is_framebuffer_srgb_enabled, int use_rgc_sort_override);
/* same, from an AoS that is already in device memory (e.g. the destination of an NCCL broadcast) */
SPL_API int spl_renderer_init_device_aos(spl_renderer*, const void* device_aos, size_t num_gaussians, size_t stride,
                                         int has_full_sh, int is_framebuffer_srgb_enabled);
/* GaussianCloud::ImportPly + Init in one step, without the host decode and without a host copy of the cloud: the PLY
 * header is parsed on the host, the raw body is decoded by a CUDA kernel into device memory and packed there */
SPL_API int spl_renderer_init_ply(spl_renderer*, const char* ply_filename, int import_full_sh, int is_framebuffer_srgb_enabled);
/* Sort :32-33: cull + project + SH + (tile|depth) key emission + device-wide radix sort + tile ranges */
SPL_API int spl_renderer_sort(spl_renderer*, const float camera_mat[16], const float proj_mat[16],
                              const float viewport[4], const float near_far[2]);
/* Render :36-37: clears the target to (0,0,0,1) and composites the sorted lists front to back */
SPL_API int spl_renderer_render(spl_renderer*, const float camera_mat[16], const float proj_mat[16],
                                const float viewport[4], const float near_far[2]);
/* the render target of the last Render (device pointer; valid until the next Render / resize / destroy) */
SPL_API int spl_renderer_framebuffer(spl_renderer*, void** device_ptr, int* width, int* height,
                                     size_t* row_pitch_bytes, int* fb_format);
/* device->host copy of the target: enqueued on the renderer's copy stream after the frame (host should be
 * pinned); two targets alternate, so the next Sort/Render overlaps the copy ... */
SPL_API int spl_renderer_read_framebuffer_async(spl_renderer*, void* host, size_t bytes);
/* ... or copy and wait; re-runs the frame once with a larger capacity if it overflowed */
SPL_API int spl_renderer_read_framebuffer(spl_renderer*, void* host, size_t bytes);
/* the PRESENTED frame: 8-bit RGBA, 4 bytes per pixel, row 0 = bottom -- the reference's RenderDesktop blit of the float
 * target into the window (src/app.cpp:1076-1082, shader/desktop_frag.glsl:32-37): clamp to [0,1], x255, round.
 * Converted on the device; same streaming rules as the float reads above. */
SPL_API int spl_renderer_read_rgba8_async(spl_renderer*, void* host, size_t bytes);
SPL_API int spl_renderer_read_rgba8(spl_renderer*, void* host, size_t bytes);
/* device-side join: the render stream waits for every framebuffer copy still in flight (no host block) */
SPL_API int spl_renderer_fence_copies(spl_renderer*);
/* host-side wait: returns when at most keep_in_flight (0 or 1) async framebuffer copies are still pending; with 1,
 * the previous frame's pixels are in host memory while the current frame's copy is still running */
SPL_API int spl_renderer_wait_copies(spl_renderer*, int keep_in_flight);
/* host waits for the render stream and the copy stream */
SPL_API int spl_renderer_synchronize(spl_renderer*);

typedef struct spl_stats {
    uint64_t num_gaussians;    /* N */
    uint64_t num_visible;      /* V: gaussians that pass K1+K6 (the reference's sortCount after the geometry cull) */
    uint64_t num_instances;    /* D: (gaussian,tile) keys emitted this frame */
    uint64_t instance_capacity;
    uint32_t overflow;         /* 1 if D exceeded capacity (frame incomplete; call spl_renderer_reserve) */
    uint32_t tiles_x, tiles_y, tile_w, tile_h; /* composite tile grid; a tile is tile_w x tile_h pixels (32 x 16) */
    uint32_t sort_passes, sort_bits;
    uint32_t sh_mode;          /* 0 = DC only, 1 = degree 1 (base layout), 3 = degree 3 (FULL_SH) */
    uint32_t sort_mode;        /* SPL_SORT_* the last Sort used */
    uint32_t kernel_launches;  /* kernels launched by the last Sort+Render */
    float stage_ms[8];         /* profiling on: [0] project  [1] LSD: depth-digit passes on the gaussians + emission (flat: 0) |
                                  binned: tile scan + scatter  [2] LSD: tile-digit passes (flat: all passes) | binned:
                                  segment sort  [3] tile ranges (0 when the last sort pass writes them)  [4] composite
                                  [7] whole frame */
} spl_stats;
SPL_API int spl_renderer_stats(spl_renderer*, spl_stats* out); /* synchronizes the stream */
SPL_API int spl_renderer_reserve(spl_renderer*, uint64_t instance_capacity);

enum {
    SPL_OPT_EARLY_OUT_T = 0,  /* stop a pixel when its transmittance drops below this (default 1e-4; 0 = exact) */
    SPL_OPT_PROFILE = 1,      /* 1: bracket stages with CUDA events (stage_ms) */
    SPL_OPT_INSTANCES_PER_GAUSSIAN = 2, /* initial capacity factor, before init (default 16) */
    SPL_OPT_SORT_TILE = 4,    /* keys per onesweep tile: 0 (default) = automatic (2048 for a sort that would not fill one wave of
                                 CTAs with 4096-key tiles, else 4096), or 2048 / 4096 for every sort of the frame */
    SPL_OPT_SORT_MODE = 3,    /* SPL_SORT_LSD (default), SPL_SORT_LSD_FLAT or SPL_SORT_BINNED; all give bit-identical tile lists */
    SPL_OPT_SUBBLOCK_CULL = 5, /* compositor's per-entry 8x4 sub-block mask: 0 none, 1 footprint box, 2 exact ellipse (default);
                                 culling only skips lanes whose weight is discarded anyway: images are bit-identical */
    SPL_OPT_GRAPH = 8,        /* 1 (default): the device work of a Sort (control-block memset + projection + all sort passes) is
                                 captured and replayed as ONE CUDA graph; the executable graph is updated in place every frame
                                 (the camera changes the kernel arguments).  0: individual launches (also used automatically
                                 while the caller is capturing the stream itself).  Results are identical. */
    SPL_OPT_PDL = 7,          /* which kernels of a frame are launched as programmatic dependents of the kernel before them (a
                                 kernel's prologue overlaps the predecessor's tail; griddepcontrol.wait guards every dependent
                                 read): bit mask, 1 = depth passes, 2 = emission, 4 = tile passes, 8 = compositor.  Default 5
                                 (measured best); 0 = plain stream order.  Results are identical. */
    SPL_OPT_SHARED_SORT = 6   /* what a Render does whose arguments differ from the last Sort's (the VR pattern: Sort for
                                 eye 0, Render for both eyes, src/app.cpp:603-607).  0 (default): sort again for the new
                                 view -- every image is exactly ordered.  1: the reference's behaviour -- keep the list and
                                 the depth order of the last Sort and only re-project (default sort path, same viewport);
                                 the four depth passes of the second eye are saved */
};
enum {
    SPL_SORT_BINNED = 0, /* MSD: counting sort on the tile id, then per-tile shared-memory sort of (depthKey, id) */
    SPL_SORT_LSD = 1,    /* device-wide onesweep LSD radix sort of the packed (tile << 32 | depthKey) u64 keys; the four
                            depth-digit passes run on the N gaussians before tile expansion, the tile-digit passes on the D instances */
    SPL_SORT_LSD_FLAT = 2 /* same sort with every digit pass on the D instances (the textbook arrangement; slower) */
};
SPL_API int spl_renderer_set_option(spl_renderer*, int option, double value);

/* parity-test taps: copy an intermediate buffer of the last frame to the host (synchronizes) */
enum {
    SPL_TAP_RECORDS = 0,    /* N x 48 B projected records {px,py,i00,i01,i10,i11,r,g,b,alpha,key,visible} */
    SPL_TAP_KEYS = 1,       /* D x u64 sorted (tile<<32 | key) */
    SPL_TAP_VALS = 2,       /* D x u32 sorted gaussian ids */
    SPL_TAP_TILE_RANGES = 3 /* tiles x {u32 begin, u32 end} */
};
SPL_API int spl_renderer_tap(spl_renderer*, int what, void* host, size_t bytes);

/* ---------------------------------------------------------------- multi-GPU start-up
 * One renderer per GPU (a process or a thread per GPU); views shard one per GPU; the cloud is replicated ONCE by an
 * ncclBroadcast over NVLink / NVSwitch -- the multi-GPU form of App::Init's load + SplatRenderer::Init
 * (src/app.cpp:535-559).  No call on the frame path is collective.  NCCL is loaded at run time (libnccl.so.2). */
typedef struct spl_comm spl_comm;
#define SPL_COMM_ID_BYTES 128 /* sizeof(ncclUniqueId) */
SPL_API int spl_comm_get_unique_id(unsigned char id[SPL_COMM_ID_BYTES]); /* on one rank; ship the bytes to the others */
SPL_API spl_comm* spl_comm_create(const unsigned char id[SPL_COMM_ID_BYTES], int rank, int nranks, int device); /* collective */
SPL_API void spl_comm_destroy(spl_comm*);
SPL_API int spl_comm_nccl_version(void); /* ncclGetVersion() of the loaded copy; 0: NCCL could not be loaded */
/* collective: every rank calls it, `root` passes its imported cloud, the others NULL.  = spl_renderer_init on every rank
 * from one host copy of the cloud.  The renderer must have been configured for the communicator's device.
 * broadcast_ms (optional): device time of the ncclBroadcast of the AoS. */
SPL_API int spl_renderer_init_broadcast(spl_renderer*, spl_comm*, const spl_cloud* cloud_on_root, int root,
                                        int is_framebuffer_srgb_enabled, float* broadcast_ms);

/* ---------------------------------------------------------------- CamerasConfig (src/camerasconfig.h:12-32) */
typedef struct spl_cameras spl_cameras;
SPL_API spl_cameras* spl_cameras_create(void);
SPL_API void spl_cameras_destroy(spl_cameras*);
SPL_API int spl_cameras_import_json(spl_cameras*, const char* json_filename);           /* ImportJson :19 */
SPL_API size_t spl_cameras_count(const spl_cameras*);                                   /* GetNumCameras :22 */
/* Camera{mat, fov} of GetCameraVec()[index]: camera-to-world column-major float[16] (what Sort/Render take), fov[2] */
SPL_API int spl_cameras_get(const spl_cameras*, size_t index, float camera_mat[16], float fov[2]);
SPL_API int spl_cameras_estimate_floor_plane(const spl_cameras*, float normal_out[3], float pos_out[3]); /* :24 */

/* ---------------------------------------------------------------- stand-alone u64|u32 radix sort */
SPL_API spl_sorter* spl_sorter_create(int device, uint64_t capacity);
SPL_API void spl_sorter_destroy(spl_sorter*);
/* stable ascending LSD sort of n (u64 key, u32 val) pairs on bits [begin_bit, end_bit); device pointers;
 * result lands in (keys_out, vals_out) which must not alias the inputs (inputs are clobbered).
 * All four pointers must be 16-byte aligned (SPL_ERR_ARG otherwise) and every buffer must extend at least 16 bytes past
 * its n-th element: tiles are staged by the TMA engine (cp.async.bulk), which moves whole 16-byte units, so the last tile
 * reads up to 12 bytes past the last value / 8 past the last key (never used). */
SPL_API int spl_sorter_sort(spl_sorter*, void* d_keys_in, void* d_vals_in, void* d_keys_out, void* d_vals_out,
                            uint64_t n, int begin_bit, int end_bit, void* cuda_stream);
/* convenience for tests: host arrays in, host arrays out */
SPL_API int spl_sort_pairs_host(const uint64_t* keys, const uint32_t* vals, uint64_t n, int begin_bit, int end_bit,
                                uint64_t* keys_out, uint32_t* vals_out);

/* egress: write a host copy of the float target as an 8-bit binary PPM ("P6"), top row first, values
 * clamped to [0,1] like the reference's default RGBA8 window target (the 1:1 blit of src/app.cpp:1076-1082,
 * shader/desktop_frag.glsl:32-37).  fb_format tells whether `host_fb` holds float4 or half4 pixels (row 0 = bottom). */
SPL_API int spl_write_ppm(const char* path, const void* host_fb, int width, int height, int fb_format);

/* pinned host memory for the end-to-end path */
SPL_API void* spl_host_alloc(size_t bytes);
/* same, with its pages on the NUMA node the GPU `device` hangs off (2-socket hosts: a framebuffer destination on the
 * other socket makes every D->H copy cross the socket interconnect); plain pinned memory when the node is unknown */
SPL_API void* spl_host_alloc_near(size_t bytes, int device);
SPL_API void spl_host_free(void*); /* frees blocks of either allocator */
SPL_API int spl_device_numa_node(int device); /* NUMA node of the GPU, -1 if the platform reports none */
/* move the calling thread to the CPUs of the GPU's NUMA node (sched_setaffinity); returns the node, or -1 if nothing
 * was changed.  Call it first thing in a rank's thread / process, before allocating host buffers. */
SPL_API int spl_bind_thread_near(int device);

#ifdef __cplusplus
}
#endif
#endif /* SPLAT_B200_H */
