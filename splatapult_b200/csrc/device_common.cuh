// device_common.cuh -- types shared by the sm_100a kernels of the splat hot path, and PTX helpers.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace splat
{

// ---------------------------------------------------------------------------------------------------
// pipeline geometry
// ---------------------------------------------------------------------------------------------------
// composite tile = 32 x 16 pixels: one CTA of 8 warps, each warp an 8x8 pixel block, TWO pixels per thread (packed f32x2
// arithmetic, composite.cu).  Wider tiles also mean fewer (gaussian, tile) instances to emit, sort and stage.
constexpr int kTileW = 32, kTileH = 16;
constexpr int kTileShiftX = 5, kTileShiftY = 4;
constexpr int kDepthBits = 32;       // low half of the sort key: the reference's u32 depth key
constexpr int kRadixBits = 8;
constexpr int kRadixSize = 1 << kRadixBits;
constexpr int kMaxSortPasses = 8;    // 64-bit key / 8

constexpr int kSortThreads = 256;    // onesweep CTA
constexpr int kSortKeysPerThread = 16;
constexpr int kSortTileKeys = kSortThreads * kSortKeysPerThread;  // 4096 keys = 32 KiB staged by one bulk copy
constexpr int kSortMinTileKeys = kSortThreads * 8;                // smallest tile variant (sizes the look-back array)
// look-back words carry a 5-bit generation instead of being cleared per sort: the owner of the status array bumps its
// generation per sort and zeroes the array whenever (generation & mask) wraps to 0 (every 31 sorts)
constexpr uint32_t kSortGenerationMask = 31;

constexpr uint64_t kMaxInstances = (1ull << 30) - 1;  // tile instances one frame may hold

// per-tile segment sort (binsort.cu): segments up to kSegSortSmallCap are sorted by a 256-thread CTA in 2 x 16 KiB
// of static shared memory, up to kSegSortLargeCap by a 1024-thread CTA in 2 x 96 KiB, longer ones through global memory
constexpr int kSegSortSmallCap = 2048;
constexpr int kSegSortLargeCap = 12288;

// ---------------------------------------------------------------------------------------------------
// per-frame constants (host computes them exactly like the reference's CPU side does:
// view = inverse(cameraMat), mvp = projMat * view   src/splatrenderer.cpp:161,175,327-328)
// ---------------------------------------------------------------------------------------------------
struct FrameConstants
{
    float view[16];
    float mvp[16];  // of the last Sort: cull + depth key only (presort_compute.glsl:42-56); the draw uses view and proj
    float proj[16];
    float eye[3];
    float W, H, zNear, zFar;
    int width, height, tilesX, tilesY;
    int srgb;
    // shared sort (the reference's VR path sorts for eye 0 and draws both eyes in that order, src/app.cpp:603-607):
    float zFarSort;   // far plane of the Sort call (= zFar unless a Render re-projects for another eye)
    int stereoPass;   // 0: Sort and Render see the same view.  1: Sort of a shared-sort frame (every gaussian that passes
                      // the pre-sort cull keeps its depth key, whether or not this eye draws it).  2: Render for
                      // another eye: view/proj/eye/zNear/zFar are that eye's, mvp/zFarSort still the Sort's; no keys.
};

// cloud packed at Init into coalescing-friendly planes (DESIGN.md "HBM layout")
struct PackedCloud
{
    const float4* posAlpha;  // xyz + alpha
    const float4* covA;      // cov3_col0.xyz, cov3_col1.x
    const float4* covB;      // cov3_col1.yz, cov3_col2.xy
    const float* covC;       // cov3_col2.z
    const float* dc[3];      // shMode 0: f_dc per channel
    const float4* sh[12];    // shMode 1: r_sh0,g_sh0,b_sh0 ; shMode 3: r_sh0..3,g_sh0..3,b_sh0..3
    uint32_t n;
    int shMode;              // 0 = DC only, 1 = degree 1, 3 = degree 3
};

// one projected gaussian: what the reference's vertex+geometry stages hand to the fragment stage
// (splat_vert.glsl:43-45 geom_color/geom_cov2/geom_p, splat_geom.glsl:42 cov2Dinv) plus the pre-sort key
struct __align__(16) ProjRecord
{
    float px, py;              // screen centre, GL window coords
    float i00, i01, i10, i11;  // inverse 2D covariance, inv[col][row]
    float r, g, b;             // unclamped radiance
    float alpha;
    uint32_t key;              // 0xFFFFFFFF - uint(depth/far * 2^32)
    uint32_t visible;
};
static_assert(sizeof(ProjRecord) == 48, "ProjRecord must be 48 bytes");

// device-resident counters of one frame: nothing here is read by the host on the frame path
struct FrameCounters
{
    unsigned long long totalInstances;  // D the frame needs (may exceed capacity)
    unsigned int usedInstances;         // D actually emitted and sorted
    unsigned int visible;               // V
    unsigned int overflow;
    unsigned int projectTicket;         // dynamic block ids of the projection kernel
    unsigned int sortTickets[kMaxSortPasses];
    unsigned int largeTiles;            // tiles whose segment exceeds kSegSortSmallCap (binned sort)
    unsigned int sortCount;             // elements of the per-gaussian depth sort (default LSD path): N
};

// ---------------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// 1-D bulk asynchronous copy global -> shared (TMA engine; SASS UBLKCP), completion on an mbarrier.
// dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smemDst, const void* gmemSrc, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smemDst)),
                 "l"(gmemSrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// order prior generic-proxy shared-memory accesses before subsequent async-proxy (bulk copy) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// shared-memory reduction without nvcc's automatic warp-aggregation loop (which costs ~one iteration per distinct
// address in the warp -- a pessimisation for histogram updates whose addresses are mostly distinct)
__device__ __forceinline__ void smem_red_add(uint32_t* p, uint32_t v)
{
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// MUFU.SQRT (max relative error 2^-23 over the normal range): for conservative bounds that are inflated anyway
__device__ __forceinline__ float sqrt_approx(float x)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Programmatic dependent launch (a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start before
// its predecessor in the stream has finished; both instructions are no-ops in a normally launched kernel):
//   pdl_wait()    -- returns when the preceding kernel has completed and its memory operations are visible.  Everything a
//                    kernel reads from ANY earlier kernel of the frame must come after it ...
//   pdl_trigger() -- ... and is called only AFTER the kernel's own pdl_wait(): the next kernel is launched once every CTA
//                    of this one has triggered (or exited), so the next kernel's code BEFORE its wait runs after this
//                    kernel's predecessor has completed, but concurrently with this kernel's tail.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// NOTE: data the preceding kernels produce must not be read through the read-only path (`const T* __restrict__` parameters,
// __ldg): those loads are LDG.CONSTANT, which the compiler and ptxas move freely -- also above this wait (seen in the SASS:
// the sorter's histogram read sat above ACQBULK and read the row before its producer had finished).  Use plain or
// ld.relaxed loads for such data.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

}  // namespace splat
