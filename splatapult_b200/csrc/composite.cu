// composite.cu -- tile ranges + C: per-tile front-to-back alpha compositing into a float framebuffer.
//
// Replaces the reference's draw: glDrawElements(GL_POINTS, sortCount) (src/splatrenderer.cpp:338) ->
// fixed-function rasteriser -> shader/splat_frag.glsl:18-42 (K7: g = exp(-1/2 d^T Sigma2^-1 d), w = alpha g,
// discard iff w <= 1/256, premultiplied output) -> ROP blend GL_ONE, GL_ONE_MINUS_SRC_ALPHA over a target
// cleared to (0,0,0,1) (K8, src/app.cpp:144-164), strictly back to front.
//
// Here one CTA owns one 16x16-pixel tile and walks the tile's sorted list.  The list is stored back-to-front
// (ascending reference key); the CTA walks it from its END, i.e. front to back, accumulating
//     C += T * w * rgb ;  T *= (1 - w)
// which is the closed form of the reference's "over" chain up to float reassociation, and
// lets a pixel stop once its transmittance T can no longer change the result by more than earlyOutT
// (a bounded optimisation, earlyOutT = 0 reproduces the reference's full chain).
// Pixel centres are at (x+0.5, y+0.5), row 0 is the bottom row (GL window coordinates).
#include <cuda_fp16.h>

#include "device_common.cuh"
#include "kernels.h"

namespace splat
{

constexpr int kCompThreads = kTileSize * kTileSize;  // 256: one thread per pixel
constexpr int kCompChunk = 256;                      // list entries staged per round

// ---------------------------------------------------------------------------------------------------
// tile ranges: [begin, end) of each tile's run in the sorted key array
// ---------------------------------------------------------------------------------------------------
__global__ void k_tile_ranges(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ nDevice, uint2* __restrict__ ranges,
                              uint32_t numTiles)
{
    const uint32_t n = *nDevice;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint32_t tile = (uint32_t)(keys[i] >> 32);
        if (tile >= numTiles) continue;  // cannot happen for keys this library emits
        if (i == 0 || (uint32_t)(keys[i - 1] >> 32) != tile) ranges[tile].x = i;
        if (i == n - 1 || (uint32_t)(keys[i + 1] >> 32) != tile) ranges[tile].y = i + 1;
    }
}

cudaError_t launch_tile_ranges(const unsigned long long* sortedKeys, const uint32_t* nDevice, uint64_t capacity, uint2* ranges, uint32_t numTiles,
                               int numSMs, cudaStream_t stream)
{
    cudaError_t e = cudaMemsetAsync(ranges, 0, (size_t)numTiles * sizeof(uint2), stream);
    if (e != cudaSuccess) return e;
    const uint32_t blocks = (uint32_t)max(1ull, min((unsigned long long)numSMs * 8ull, (unsigned long long)((capacity + 255) / 256)));
    k_tile_ranges<<<blocks, 256, 0, stream>>>(sortedKeys, nDevice, ranges, numTiles);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// composite
// ---------------------------------------------------------------------------------------------------
// staged entry, coefficients folded for exp2:  w = alpha * 2^(A dx^2 + B dx dy + C dy^2)
struct __align__(16) StagedSplat
{
    float cx, cy, A, B;
    float C, alpha, r, g;
    float b, pad0, pad1, pad2;
};

template <bool HALF>
__global__ void __launch_bounds__(kCompThreads)
k_composite(const uint2* __restrict__ ranges, const uint32_t* __restrict__ vals, const ProjRecord* __restrict__ rec, void* __restrict__ fb,
            int width, int height, int tilesX, float earlyOutT)
{
    __shared__ float4 sA[kCompChunk], sB[kCompChunk];
    __shared__ float sBlue[kCompChunk];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int tileX = tile % tilesX, tileY = tile / tilesX;
    // warp w covers an 8x4 pixel block: 2 blocks across, 4 up
    const int x = tileX * kTileSize + (warp & 1) * 8 + (lane & 7);
    const int y = tileY * kTileSize + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = x < width && y < height;
    const float fx = (float)x + 0.5f, fy = (float)y + 0.5f;  // gl_FragCoord.xy

    const uint2 range = ranges[tile];
    float T = 1.0f, cr = 0.0f, cg = 0.0f, cb = 0.0f;
    bool done = !inside;

    const float kLog2e = 1.4426950408889634f;
    for (uint32_t hi = range.y; hi > range.x;)
    {
        const uint32_t cnt = min((uint32_t)kCompChunk, hi - range.x);
        // stage entries hi-1, hi-2, ... (nearest first)
        if ((uint32_t)tid < cnt)
        {
            const uint32_t id = __ldg(vals + (hi - 1 - tid));
            const float4* p = reinterpret_cast<const float4*>(rec + id);
            const float4 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2);
            // q0 = px,py,i00,i01 ; q1 = i10,i11,r,g ; q2 = b,alpha,key,visible
            const float h = -0.5f * kLog2e;
            sA[tid] = make_float4(q0.x, q0.y, h * q0.z, h * (q0.w + q1.x));
            sB[tid] = make_float4(h * q1.y, q2.y, q1.z, q1.w);
            sBlue[tid] = q2.x;
        }
        __syncthreads();
        if (!__all_sync(0xffffffffu, done))
        {
            for (uint32_t j = 0; j < cnt; j++)
            {
                const float4 a = sA[j], b = sB[j];
                const float dx = fx - a.x, dy = fy - a.y;
                const float pw = dx * (a.z * dx + a.w * dy) + (b.x * dy) * dy;
                const float w = b.y * ex2_approx(pw);
                if (w > (1.0f / 256.0f) && !done)
                {
                    const float tw = T * w;
                    cr = fmaf(tw, b.z, cr);
                    cg = fmaf(tw, b.w, cg);
                    cb = fmaf(tw, sBlue[j], cb);
                    T = T - tw;
                    done = T < earlyOutT;
                }
            }
        }
        hi -= cnt;
        if (__syncthreads_and(done)) break;
    }

    if (inside)
    {
        const size_t pix = (size_t)y * width + x;
        if (HALF)
        {
            __half2 lo = __floats2half2_rn(cr, cg), hi2 = __floats2half2_rn(cb, 1.0f);
            uint2 v;
            v.x = *reinterpret_cast<uint32_t*>(&lo);
            v.y = *reinterpret_cast<uint32_t*>(&hi2);
            reinterpret_cast<uint2*>(fb)[pix] = v;
        }
        else
        {
            reinterpret_cast<float4*>(fb)[pix] = make_float4(cr, cg, cb, 1.0f);
        }
    }
}

cudaError_t launch_composite(const uint2* ranges, const uint32_t* sortedVals, const ProjRecord* rec, void* framebuffer, int fbFormat, int width,
                             int height, int tilesX, int tilesY, float earlyOutT, cudaStream_t stream)
{
    const int tiles = tilesX * tilesY;
    if (tiles <= 0) return cudaSuccess;
    if (fbFormat == 1)
        k_composite<true><<<tiles, kCompThreads, 0, stream>>>(ranges, sortedVals, rec, framebuffer, width, height, tilesX, earlyOutT);
    else
        k_composite<false><<<tiles, kCompThreads, 0, stream>>>(ranges, sortedVals, rec, framebuffer, width, height, tilesX, earlyOutT);
    return cudaGetLastError();
}

}  // namespace splat
