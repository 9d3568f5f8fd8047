// capi.cu -- extern "C" surface of libsplat_b200.so (include/splat_b200.h): plain handles over the C++ classes.
#include <cuda_fp16.h>

#include <cstdio>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "kernels.h"
#include "multigpu.h"
#include "splat_b200.h"
#include "splatapult_b200/gaussiancloud.h"
#include "splatapult_b200/splatrenderer.h"

struct spl_cloud
{
    std::shared_ptr<GaussianCloud> gc;
};
struct spl_renderer
{
    SplatRenderer r;
};

namespace
{
thread_local std::string g_lastError;
int set_error(int code, const std::string& msg)
{
    g_lastError = msg;
    return code;
}
// no C++ exception crosses the C ABI: allocation failures and anything else a host-side call throws become status codes
template <typename F>
int guarded(F&& f) noexcept
{
    try
    {
        return f();
    }
    catch (const std::bad_alloc&)
    {
        return set_error(SPL_ERR_ALLOC, "out of host memory");
    }
    catch (const std::exception& e)
    {
        return set_error(SPL_ERR_ARG, std::string("C++ exception: ") + e.what());
    }
    catch (...)
    {
        return set_error(SPL_ERR_ARG, "unknown C++ exception");
    }
}
int from_renderer(spl_renderer* h, bool okFlag)
{
    if (okFlag) return SPL_OK;
    g_lastError = h->r.LastError();
    return h->r.LastStatus() ? h->r.LastStatus() : SPL_ERR_ARG;
}
}  // namespace

extern "C" {

const char* spl_last_error(void) { return g_lastError.c_str(); }
int spl_version(void) { return 100; }
int spl_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ------------------------------------------------------------------ cloud
spl_cloud* spl_cloud_create(int import_full_sh, int export_full_sh)
{
    spl_cloud* c = new (std::nothrow) spl_cloud;
    if (!c) return nullptr;
    GaussianCloud::Options opt{import_full_sh != 0, export_full_sh != 0};
    c->gc = std::make_shared<GaussianCloud>(opt);
    return c;
}
void spl_cloud_destroy(spl_cloud* c) { delete c; }

int spl_cloud_import_ply(spl_cloud* c, const char* path)
{
    return guarded([&]() -> int {
    if (!c || !path) return set_error(SPL_ERR_ARG, "null argument");
    if (!c->gc->ImportPly(path)) return set_error(SPL_ERR_IO, c->gc->GetError());
    return SPL_OK;
    });
}
int spl_cloud_import_ply_gpu(spl_cloud* c, const char* path, int device)
{
    return guarded([&]() -> int {
    if (!c || !path) return set_error(SPL_ERR_ARG, "null argument");
    if (spl_device_count() < 1) return set_error(SPL_ERR_CUDA, "no CUDA device: spl_cloud_import_ply_gpu has no CPU fallback (use spl_cloud_import_ply)");
    if (!c->gc->ImportPlyGpu(path, device)) return set_error(SPL_ERR_IO, c->gc->GetError());
    return SPL_OK;
    });
}
int spl_cloud_export_ply(const spl_cloud* c, const char* path)
{
    return guarded([&]() -> int {
    if (!c || !path) return set_error(SPL_ERR_ARG, "null argument");
    if (!c->gc->ExportPly(path)) return set_error(SPL_ERR_IO, c->gc->GetError());
    return SPL_OK;
    });
}
int spl_cloud_init_debug_cloud(spl_cloud* c)
{
    return guarded([&]() -> int {
    if (!c) return set_error(SPL_ERR_ARG, "null argument");
    c->gc->InitDebugCloud();
    return SPL_OK;
    });
}
int spl_cloud_prune_splats(spl_cloud* c, const float origin[3], uint32_t count)
{
    return guarded([&]() -> int {
    if (!c || !origin) return set_error(SPL_ERR_ARG, "null argument");
    c->gc->PruneSplats(splat::vec3{origin[0], origin[1], origin[2]}, count);
    return SPL_OK;
    });
}
int spl_cloud_set_raw(spl_cloud* c, const void* aos, size_t n, size_t stride, int has_full_sh)
{
    return guarded([&]() -> int {
    if (!c || (!aos && n)) return set_error(SPL_ERR_ARG, "null argument");
    if (!c->gc->SetRawData(aos, n, stride, has_full_sh != 0)) return set_error(SPL_ERR_ARG, c->gc->GetError());
    return SPL_OK;
    });
}
size_t spl_cloud_num_gaussians(const spl_cloud* c) { return c ? c->gc->GetNumGaussians() : 0; }
size_t spl_cloud_stride(const spl_cloud* c) { return c ? c->gc->GetStride() : 0; }
size_t spl_cloud_total_size(const spl_cloud* c) { return c ? c->gc->GetTotalSize() : 0; }
void* spl_cloud_raw_data(spl_cloud* c) { return c ? c->gc->GetRawDataPtr() : nullptr; }
int spl_cloud_has_full_sh(const spl_cloud* c) { return c && c->gc->HasFullSH(); }
int spl_cloud_attrib(const spl_cloud* c, int id, int* type, size_t* size, size_t* offset)
{
    if (!c || id < 0 || id >= SPL_ATTR_COUNT) return set_error(SPL_ERR_ARG, "bad attribute id");
    const GaussianCloud& g = *c->gc;
    const BinaryAttribute* table[SPL_ATTR_COUNT] = {
        &g.GetPosWithAlphaAttrib(),
        &g.GetR_SH0Attrib(), &g.GetR_SH1Attrib(), &g.GetR_SH2Attrib(), &g.GetR_SH3Attrib(),
        &g.GetG_SH0Attrib(), &g.GetG_SH1Attrib(), &g.GetG_SH2Attrib(), &g.GetG_SH3Attrib(),
        &g.GetB_SH0Attrib(), &g.GetB_SH1Attrib(), &g.GetB_SH2Attrib(), &g.GetB_SH3Attrib(),
        &g.GetCov3_Col0Attrib(), &g.GetCov3_Col1Attrib(), &g.GetCov3_Col2Attrib()};
    const BinaryAttribute& a = *table[id];
    if (type) *type = (int)a.type;
    if (size) *size = a.size;
    if (offset) *offset = a.offset;
    return SPL_OK;
}

// ------------------------------------------------------------------ renderer
spl_renderer* spl_renderer_create(void) { return new (std::nothrow) spl_renderer; }
void spl_renderer_destroy(spl_renderer* h) { delete h; }

int spl_renderer_configure(spl_renderer* h, int device, int fb_format, void* cuda_stream)
{
    if (!h || (fb_format != SPL_FB_FP32 && fb_format != SPL_FB_FP16)) return set_error(SPL_ERR_ARG, "bad argument");
    return from_renderer(h, h->r.Configure(device, (SplatRenderer::FrameBuffer)fb_format, cuda_stream));
}
int spl_renderer_init(spl_renderer* h, const spl_cloud* c, int srgb, int use_rgc)
{
    return guarded([&]() -> int {
    if (!h || !c) return set_error(SPL_ERR_ARG, "null argument");
    return from_renderer(h, h->r.Init(c->gc, srgb != 0, use_rgc != 0));
    });
}
int spl_renderer_init_device_aos(spl_renderer* h, const void* device_aos, size_t n, size_t stride, int has_full_sh, int srgb)
{
    return guarded([&]() -> int {
    if (!h || (!device_aos && n)) return set_error(SPL_ERR_ARG, "null argument");
    return from_renderer(h, h->r.InitFromDevice(device_aos, n, stride, has_full_sh != 0, srgb != 0));
    });
}
int spl_renderer_init_ply(spl_renderer* h, const char* path, int import_full_sh, int srgb)
{
    return guarded([&]() -> int {
    if (!h || !path) return set_error(SPL_ERR_ARG, "null argument");
    return from_renderer(h, h->r.InitFromPly(path, import_full_sh != 0, srgb != 0));
    });
}
int spl_renderer_sort(spl_renderer* h, const float cam[16], const float proj[16], const float vp[4], const float nf[2])
{
    return guarded([&]() -> int {
    if (!h || !cam || !proj || !vp || !nf) return set_error(SPL_ERR_ARG, "null argument");
    h->r.Sort(splat::mat4::from(cam), splat::mat4::from(proj), splat::vec4{vp[0], vp[1], vp[2], vp[3]}, splat::vec2{nf[0], nf[1]});
    return from_renderer(h, h->r.LastStatus() == 0);
    });
}
int spl_renderer_render(spl_renderer* h, const float cam[16], const float proj[16], const float vp[4], const float nf[2])
{
    return guarded([&]() -> int {
    if (!h || !cam || !proj || !vp || !nf) return set_error(SPL_ERR_ARG, "null argument");
    h->r.Render(splat::mat4::from(cam), splat::mat4::from(proj), splat::vec4{vp[0], vp[1], vp[2], vp[3]}, splat::vec2{nf[0], nf[1]});
    return from_renderer(h, h->r.LastStatus() == 0);
    });
}
int spl_renderer_framebuffer(spl_renderer* h, void** device_ptr, int* w, int* hgt, size_t* pitch, int* fmt)
{
    if (!h) return set_error(SPL_ERR_ARG, "null argument");
    SplatRenderer::FrameBuffer f;
    if (!h->r.GetFramebuffer(device_ptr, w, hgt, pitch, &f)) return set_error(SPL_ERR_ARG, "no frame rendered yet");
    if (fmt) *fmt = (int)f;
    return SPL_OK;
}
int spl_renderer_read_framebuffer_async(spl_renderer* h, void* host, size_t bytes)
{
    if (!h || !host) return set_error(SPL_ERR_ARG, "null argument");
    return from_renderer(h, h->r.ReadFramebufferAsync(host, bytes));
}
int spl_renderer_read_framebuffer(spl_renderer* h, void* host, size_t bytes)
{
    return guarded([&]() -> int {
    if (!h || !host) return set_error(SPL_ERR_ARG, "null argument");
    return from_renderer(h, h->r.ReadFramebuffer(host, bytes));
    });
}
int spl_renderer_read_rgba8_async(spl_renderer* h, void* host, size_t bytes)
{
    if (!h || !host) return set_error(SPL_ERR_ARG, "null argument");
    return from_renderer(h, h->r.ReadPresented8Async(host, bytes));
}
int spl_renderer_read_rgba8(spl_renderer* h, void* host, size_t bytes)
{
    return guarded([&]() -> int {
    if (!h || !host) return set_error(SPL_ERR_ARG, "null argument");
    return from_renderer(h, h->r.ReadPresented8(host, bytes));
    });
}
int spl_renderer_fence_copies(spl_renderer* h)
{
    if (!h) return set_error(SPL_ERR_ARG, "null argument");
    return from_renderer(h, h->r.FenceCopies());
}
int spl_renderer_wait_copies(spl_renderer* h, int keep_in_flight)
{
    if (!h) return set_error(SPL_ERR_ARG, "null argument");
    return from_renderer(h, h->r.WaitCopies(keep_in_flight));
}
int spl_renderer_synchronize(spl_renderer* h)
{
    if (!h) return set_error(SPL_ERR_ARG, "null argument");
    return from_renderer(h, h->r.Synchronize());
}
int spl_renderer_stats(spl_renderer* h, spl_stats* out)
{
    if (!h || !out) return set_error(SPL_ERR_ARG, "null argument");
    return from_renderer(h, h->r.GetStats(out));
}
int spl_renderer_reserve(spl_renderer* h, uint64_t cap)
{
    return guarded([&]() -> int {
    if (!h) return set_error(SPL_ERR_ARG, "null argument");
    return from_renderer(h, h->r.Reserve(cap));
    });
}
int spl_renderer_set_option(spl_renderer* h, int option, double value)
{
    if (!h) return set_error(SPL_ERR_ARG, "null argument");
    return from_renderer(h, h->r.SetOption(option, value));
}
int spl_renderer_tap(spl_renderer* h, int what, void* host, size_t bytes)
{
    return guarded([&]() -> int {
    if (!h || !host) return set_error(SPL_ERR_ARG, "null argument");
    return from_renderer(h, h->r.Tap(what, host, bytes));
    });
}

}  // extern "C"

// ------------------------------------------------------------------ cameras.json
#include "splatapult_b200/camerasconfig.h"
struct spl_cameras
{
    CamerasConfig cfg;
};
extern "C" {
spl_cameras* spl_cameras_create(void) { return new (std::nothrow) spl_cameras; }
void spl_cameras_destroy(spl_cameras* c) { delete c; }
int spl_cameras_import_json(spl_cameras* c, const char* path)
{
    return guarded([&]() -> int {
    if (!c || !path) return set_error(SPL_ERR_ARG, "null argument");
    if (!c->cfg.ImportJson(path)) return set_error(SPL_ERR_IO, c->cfg.GetError());
    return SPL_OK;
    });
}
size_t spl_cameras_count(const spl_cameras* c) { return c ? c->cfg.GetNumCameras() : 0; }
int spl_cameras_get(const spl_cameras* c, size_t index, float camera_mat[16], float fov[2])
{
    if (!c || index >= c->cfg.GetNumCameras()) return set_error(SPL_ERR_ARG, "camera index out of range");
    const Camera& cam = c->cfg.GetCameraVec()[index];
    if (camera_mat) std::memcpy(camera_mat, cam.mat.m, sizeof cam.mat.m);
    if (fov) { fov[0] = cam.fov.x; fov[1] = cam.fov.y; }
    return SPL_OK;
}
int spl_cameras_estimate_floor_plane(const spl_cameras* c, float normal_out[3], float pos_out[3])
{
    if (!c || !normal_out || !pos_out) return set_error(SPL_ERR_ARG, "null argument");
    splat::vec3 n, p;
    c->cfg.EstimateFloorPlane(n, p);
    normal_out[0] = n.x; normal_out[1] = n.y; normal_out[2] = n.z;
    pos_out[0] = p.x; pos_out[1] = p.y; pos_out[2] = p.z;
    return SPL_OK;
}
}  // extern "C"

// ------------------------------------------------------------------ stand-alone sorter
struct spl_sorter
{
    int device = 0;
    int numSMs = 148;
    uint64_t capacity = 0;
    void* control = nullptr;  // n | tickets | histograms
    unsigned long long* status = nullptr;
    unsigned long long* groups = nullptr;  // group aggregates of the look-back, zeroed per sort
    size_t groupWordsPerPass = 0;
    uint32_t generation = 0;
    int keysPerThread = 16;
};

namespace
{
constexpr size_t kSorterControlBytes = 256 + splat::kMaxSortPasses * splat::kRadixSize * sizeof(uint32_t);
#define SORTER_TRY(expr)                                                                                   \
    do                                                                                                     \
    {                                                                                                      \
        cudaError_t e__ = (expr);                                                                          \
        if (e__ != cudaSuccess) return set_error(SPL_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)
}  // namespace

extern "C" {

spl_sorter* spl_sorter_create(int device, uint64_t capacity)
{
    if (cudaSetDevice(device) != cudaSuccess) { set_error(SPL_ERR_CUDA, "cudaSetDevice failed (no CPU fallback exists)"); cudaGetLastError(); return nullptr; }
    spl_sorter* s = new (std::nothrow) spl_sorter;
    if (!s) return nullptr;
    s->device = device;
    s->capacity = capacity < (uint64_t)splat::kSortTileKeys ? splat::kSortTileKeys : capacity;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) s->numSMs = prop.multiProcessorCount;
    s->groupWordsPerPass = splat::sort_group_words(s->capacity, s->keysPerThread);
    if (cudaMalloc(&s->control, kSorterControlBytes) != cudaSuccess ||
        cudaMalloc(&s->status, splat::sort_status_bytes(s->capacity, splat::kMaxSortPasses)) != cudaSuccess ||
        cudaMalloc(&s->groups, splat::kMaxSortPasses * s->groupWordsPerPass * sizeof(unsigned long long)) != cudaSuccess)
    {
        set_error(SPL_ERR_ALLOC, "cudaMalloc failed");
        cudaFree(s->control);
        cudaFree(s->status);
        cudaFree(s->groups);
        delete s;
        return nullptr;
    }
    cudaMemset(s->status, 0, splat::sort_status_bytes(s->capacity, splat::kMaxSortPasses));  // look-back words start invalid
    return s;
}
void spl_sorter_destroy(spl_sorter* s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    cudaFree(s->control);
    cudaFree(s->status);
    cudaFree(s->groups);
    delete s;
}

int spl_sorter_sort(spl_sorter* s, void* d_keys_in, void* d_vals_in, void* d_keys_out, void* d_vals_out, uint64_t n, int begin_bit, int end_bit,
                    void* cuda_stream)
{
    if (!s || !d_keys_in || !d_vals_in || !d_keys_out || !d_vals_out) return set_error(SPL_ERR_ARG, "null argument");
    if (n > s->capacity || n > splat::kMaxInstances) return set_error(SPL_ERR_ARG, "n exceeds sorter capacity");
    if (begin_bit < 0 || end_bit > 64 || end_bit <= begin_bit) return set_error(SPL_ERR_ARG, "bad bit range");
    if ((reinterpret_cast<uintptr_t>(d_keys_in) | reinterpret_cast<uintptr_t>(d_vals_in) | reinterpret_cast<uintptr_t>(d_keys_out) |
         reinterpret_cast<uintptr_t>(d_vals_out)) & 15u)
        return set_error(SPL_ERR_ARG, "sort buffers must be 16-byte aligned (tiles are staged by cp.async.bulk)");
    SORTER_TRY(cudaSetDevice(s->device));
    cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
    const int passes = (end_bit - begin_bit + splat::kRadixBits - 1) / splat::kRadixBits;
    SORTER_TRY(cudaMemsetAsync(s->control, 0, kSorterControlBytes, stream));
    const uint32_t n32 = (uint32_t)n;
    SORTER_TRY(cudaMemcpyAsync(s->control, &n32, sizeof n32, cudaMemcpyHostToDevice, stream));
    uint32_t* ctl = static_cast<uint32_t*>(s->control);
    if (((++s->generation) & splat::kSortGenerationMask) == 0)
    {
        SORTER_TRY(cudaMemsetAsync(s->status, 0, splat::sort_status_bytes(s->capacity, splat::kMaxSortPasses), stream));
        ++s->generation;
    }
    splat::SortWorkspace ws{ctl + 64, s->status, ctl + 8, s->capacity, s->generation, s->keysPerThread, 0};
    SORTER_TRY(cudaMemsetAsync(s->groups, 0, (size_t)passes * s->groupWordsPerPass * sizeof(unsigned long long), stream));
    ws.groupAgg = s->groups;
    ws.groupWordsPerPass = s->groupWordsPerPass;
    auto* kin = static_cast<unsigned long long*>(d_keys_in);
    auto* kout = static_cast<unsigned long long*>(d_keys_out);
    auto* vin = static_cast<uint32_t*>(d_vals_in);
    auto* vout = static_cast<uint32_t*>(d_vals_out);
    SORTER_TRY(splat::launch_radix_sort(kin, vin, kout, vout, ctl, s->capacity, begin_bit, end_bit, ws, s->numSMs, stream, nullptr));
    if ((passes & 1) == 0)
    {
        // an even number of passes ends in the input buffers
        SORTER_TRY(cudaMemcpyAsync(kout, kin, n * 8, cudaMemcpyDeviceToDevice, stream));
        SORTER_TRY(cudaMemcpyAsync(vout, vin, n * 4, cudaMemcpyDeviceToDevice, stream));
    }
    return SPL_OK;
}

int spl_sort_pairs_host(const uint64_t* keys, const uint32_t* vals, uint64_t n, int begin_bit, int end_bit, uint64_t* keys_out, uint32_t* vals_out)
{
    return guarded([&]() -> int {
    if ((!keys || !vals || !keys_out || !vals_out) && n) return set_error(SPL_ERR_ARG, "null argument");
    if (n == 0) return SPL_OK;
    spl_sorter* s = spl_sorter_create(0, n);
    if (!s) return SPL_ERR_CUDA;
    const uint64_t padded = n + splat::kSortTileKeys;
    void *dk0 = nullptr, *dk1 = nullptr, *dv0 = nullptr, *dv1 = nullptr;
    int rc = SPL_OK;
    auto run = [&]() -> int {
        SORTER_TRY(cudaMalloc(&dk0, padded * 8));
        SORTER_TRY(cudaMalloc(&dk1, padded * 8));
        SORTER_TRY(cudaMalloc(&dv0, padded * 4));
        SORTER_TRY(cudaMalloc(&dv1, padded * 4));
        SORTER_TRY(cudaMemset(dk0, 0, padded * 8));
        SORTER_TRY(cudaMemset(dk1, 0, padded * 8));
        SORTER_TRY(cudaMemset(dv0, 0, padded * 4));
        SORTER_TRY(cudaMemset(dv1, 0, padded * 4));
        SORTER_TRY(cudaMemcpy(dk0, keys, n * 8, cudaMemcpyHostToDevice));
        SORTER_TRY(cudaMemcpy(dv0, vals, n * 4, cudaMemcpyHostToDevice));
        if (int r = spl_sorter_sort(s, dk0, dv0, dk1, dv1, n, begin_bit, end_bit, nullptr)) return r;
        SORTER_TRY(cudaDeviceSynchronize());
        SORTER_TRY(cudaMemcpy(keys_out, dk1, n * 8, cudaMemcpyDeviceToHost));
        SORTER_TRY(cudaMemcpy(vals_out, dv1, n * 4, cudaMemcpyDeviceToHost));
        return SPL_OK;
    };
    rc = run();
    cudaFree(dk0); cudaFree(dk1); cudaFree(dv0); cudaFree(dv1);
    spl_sorter_destroy(s);
    return rc;
    });
}

int spl_write_ppm(const char* path, const void* host_fb, int width, int height, int fb_format)
{
    return guarded([&]() -> int {
    if (!path || !host_fb || width <= 0 || height <= 0) return set_error(SPL_ERR_ARG, "bad argument");
    std::vector<unsigned char> row((size_t)width * 3);  // before the file is opened: an allocation failure leaks nothing
    FILE* f = std::fopen(path, "wb");
    if (!f) return set_error(SPL_ERR_IO, std::string("failed to open ") + path);
    std::fprintf(f, "P6\n%d %d\n255\n", width, height);
    auto to8 = [](float v) { v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); return (unsigned char)(v * 255.0f + 0.5f); };
    for (int y = height - 1; y >= 0; y--)  // the framebuffer is bottom-up, images are top-down
    {
        for (int x = 0; x < width; x++)
        {
            float rgb[3];
            if (fb_format == SPL_FB_FP16)
            {
                const __half* p = static_cast<const __half*>(host_fb) + ((size_t)y * width + x) * 4;
                for (int c = 0; c < 3; c++) rgb[c] = __half2float(p[c]);
            }
            else
            {
                const float* p = static_cast<const float*>(host_fb) + ((size_t)y * width + x) * 4;
                for (int c = 0; c < 3; c++) rgb[c] = p[c];
            }
            for (int c = 0; c < 3; c++) row[(size_t)x * 3 + c] = to8(rgb[c]);
        }
        std::fwrite(row.data(), 1, row.size(), f);
    }
    std::fclose(f);
    return SPL_OK;
    });
}

void* spl_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess)
    {
        set_error(SPL_ERR_ALLOC, "cudaHostAlloc failed");
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void* spl_host_alloc_near(size_t bytes, int device)
{
    void* p = nullptr;
    guarded([&]() -> int {
        p = splat::host_alloc_near(bytes, device, nullptr);
        return p ? SPL_OK : set_error(SPL_ERR_ALLOC, "pinned host allocation failed");
    });
    return p;
}
void spl_host_free(void* p)
{
    guarded([&]() -> int { splat::host_free_any(p); return SPL_OK; });
}
int spl_device_numa_node(int device) { return splat::numa_node_of_device(device); }
int spl_bind_thread_near(int device)
{
    int node = -1, cpus = 0;
    int rc = -1;
    guarded([&]() -> int { rc = splat::bind_thread_near(device, &node, &cpus) ? node : -1; return SPL_OK; });
    return rc;
}

// ------------------------------------------------------------------ multi-GPU start-up
}  // extern "C"
struct spl_comm
{
    SplatComm c;
};
extern "C" {
int spl_comm_get_unique_id(unsigned char id[SPL_COMM_ID_BYTES])
{
    return guarded([&]() -> int {
    if (!id) return set_error(SPL_ERR_ARG, "null argument");
    std::string err;
    if (!SplatComm::GetUniqueId(id, &err)) return set_error(SPL_ERR_CUDA, err);
    return SPL_OK;
    });
}
spl_comm* spl_comm_create(const unsigned char id[SPL_COMM_ID_BYTES], int rank, int nranks, int device)
{
    spl_comm* out = nullptr;
    guarded([&]() -> int {
    if (!id) return set_error(SPL_ERR_ARG, "null argument");
    spl_comm* h = new spl_comm;
    if (!h->c.Init(id, rank, nranks, device))
    {
        set_error(SPL_ERR_CUDA, h->c.LastError());
        delete h;
        return SPL_ERR_CUDA;
    }
    out = h;
    return SPL_OK;
    });
    return out;
}
void spl_comm_destroy(spl_comm* h) { delete h; }
int spl_comm_nccl_version(void) { return SplatComm::NcclVersion(); }
int spl_renderer_init_broadcast(spl_renderer* h, spl_comm* comm, const spl_cloud* cloud_on_root, int root, int srgb, float* broadcast_ms)
{
    return guarded([&]() -> int {
    if (!h || !comm) return set_error(SPL_ERR_ARG, "null argument");
    return from_renderer(h, h->r.InitBroadcast(comm->c, cloud_on_root ? cloud_on_root->gc : nullptr, root, srgb != 0, broadcast_ms));
    });
}

}  // extern "C"
