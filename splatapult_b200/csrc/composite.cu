// composite.cu -- tile ranges + C: per-tile front-to-back alpha compositing into a float framebuffer.
//
// Replaces the reference's draw: glDrawElements(GL_POINTS, sortCount) (src/splatrenderer.cpp:338) ->
// fixed-function rasteriser -> shader/splat_frag.glsl:18-42 (K7: g = exp(-1/2 d^T Sigma2^-1 d), w = alpha g,
// discard iff w <= 1/256, premultiplied output) -> ROP blend GL_ONE, GL_ONE_MINUS_SRC_ALPHA over a target
// cleared to (0,0,0,1) (K8, src/app.cpp:144-164), strictly back to front.
//
// One CTA owns one 16x16-pixel tile (8 warps, each an 8x4 pixel block, one pixel per thread) and walks the
// tile's sorted list.  The list is stored back-to-front (ascending reference key); the CTA walks it from its
// END, i.e. front to back, accumulating
//     C += T * w * rgb ;  T *= (1 - w)
// which is the closed form of the reference's "over" chain up to float reassociation, and
// lets a pixel stop once its transmittance T can no longer change the result by more than earlyOutT
// (a bounded optimisation, earlyOutT = 0 reproduces the reference's full chain).
//
// The kernel is instruction-issue bound (ncu: profiles/), not HBM bound, so the design minimises issued
// instructions per (splat, pixel):
//   * entries are staged 256 at a time into shared memory with their exp2-folded conic; the gather of the next
//     chunk's records is issued before the current chunk is evaluated (software pipeline, one barrier per chunk);
//   * while staging, each entry gets an 8-bit mask of the 8x4 sub-blocks its footprint box can touch; a warp
//     ballots the masks of 32 entries at a time and evaluates only its hits -- warp-uniform, no divergence.
// Pixel centres are at (x+0.5, y+0.5), row 0 is the bottom row (GL window coordinates).
#include <cuda_fp16.h>

#include "device_common.cuh"
#include "kernels.h"

namespace splat
{

constexpr int kCompThreads = kTileSize * kTileSize;  // 256: one thread per pixel
constexpr int kCompChunk = kCompThreads;             // list entries staged per round (one per thread)
constexpr int kSubW = 8, kSubH = 4;                  // a warp's pixel block
constexpr int kSubCols = kTileSize / kSubW;          // 2
constexpr int kSubRows = kTileSize / kSubH;          // 4

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
struct CompSmem
{
    float4 a[2][kCompChunk];    // cx, cy, A, B     with  w = 2^(A dx^2 + B dx dy + C dy^2 + L),  L = log2(alpha)
    float4 b[2][kCompChunk];    // C, L, r, g
    float4 c[2][kCompChunk];    // b, -, -, -       (same index stride as a/b: one address computation per entry)
    uint32_t mask[2][kCompChunk];  // bit (row*2+col) set: the footprint box touches that 8x4 sub-block
};

template <bool HALF>
__global__ void __launch_bounds__(kCompThreads, 6)
k_composite(const uint2* __restrict__ ranges, const uint32_t* __restrict__ vals, int idStride, const ProjRecord* __restrict__ rec,
            void* __restrict__ fb, int width, int height, int tilesX, float earlyOutT)
{
    __shared__ CompSmem s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int tileX = tile % tilesX, tileY = tile / tilesX;
    const int subCol = warp % kSubCols, subRow = warp / kSubCols;
    const int x = tileX * kTileSize + subCol * kSubW + (lane & (kSubW - 1));
    const int y = tileY * kTileSize + subRow * kSubH + (lane / kSubW);
    const bool inside = x < width && y < height;
    const float fx = (float)x + 0.5f, fy = (float)y + 0.5f;  // gl_FragCoord.xy
    const float tileX0 = (float)(tileX * kTileSize), tileY0 = (float)(tileY * kTileSize);

    const uint2 range = ranges[tile];
    // T == 0 marks a finished pixel: every later contribution is multiplied by it
    float T = inside ? 1.0f : 0.0f, cr = 0.0f, cg = 0.0f, cb = 0.0f;

    // software pipeline: the records of chunk k+1 are in flight while chunk k is evaluated
    float4 q0, q1, q2;
    bool have = false;
    auto fetch = [&](uint32_t hi) {
        const uint32_t cnt = min((uint32_t)kCompChunk, hi - range.x);
        have = (uint32_t)tid < cnt;
        if (have)
        {
            const uint32_t id = __ldg(vals + (size_t)(hi - 1 - tid) * idStride);  // nearest first
            const float4* p = reinterpret_cast<const float4*>(rec + id);
            q0 = __ldg(p);      // px, py, i00, i01
            q1 = __ldg(p + 1);  // i10, i11, r, g
            q2 = __ldg(p + 2);  // b, alpha, key, visible
        }
    };
    auto stage = [&](int buf) {
        if (!have) return;
        const float kLog2e = 1.4426950408889634f;
        const float h = -0.5f * kLog2e;
        s.a[buf][tid] = make_float4(q0.x, q0.y, h * q0.z, h * (q0.w + q1.x));
        s.b[buf][tid] = make_float4(h * q1.y, log2f(q2.y), q1.z, q1.w);
        s.c[buf][tid] = make_float4(q2.x, 0.0f, 0.0f, 0.0f);
        // footprint box inside this tile -> sub-block mask.  Same bound as the projection kernel's tile
        // rectangle: {alpha exp(-Q/2) > 1/256} lies within |dx| <= sqrt(r2 Sxx), |dy| <= sqrt(r2 Syy),
        // r2 = 2 ln(256 alpha); Sigma = (Sigma^-1)^-1; inflated so rounding cannot drop a contributing pixel.
        const float det = q0.z * q1.y - q0.w * q1.x;
        const float r2 = 2.0f * 0.6931471805599453f * (8.0f + __log2f(q2.y));
        const float rdet = __fdividef(r2, det);
        const float ex = sqrtf(fmaxf(rdet * q1.y, 0.0f)) * 1.002f + 0.02f;
        const float ey = sqrtf(fmaxf(rdet * q0.z, 0.0f)) * 1.002f + 0.02f;
        // pixel centres (x + 0.5) in [px - ex, px + ex], relative to the tile origin
        const float lx0 = q0.x - ex - 0.5f - tileX0, lx1 = q0.x + ex - 0.5f - tileX0;
        const float ly0 = q0.y - ey - 0.5f - tileY0, ly1 = q0.y + ey - 0.5f - tileY0;
        uint32_t m = 0xffu;
        if (lx0 == lx0 && lx1 == lx1 && ly0 == ly0 && ly1 == ly1)  // NaN extents: evaluate everywhere
        {
            const int px0 = (int)fmaxf(ceilf(lx0), 0.0f), px1 = (int)fminf(floorf(lx1), (float)(kTileSize - 1));
            const int py0 = (int)fmaxf(ceilf(ly0), 0.0f), py1 = (int)fminf(floorf(ly1), (float)(kTileSize - 1));
            m = 0u;
            if (px1 >= px0 && py1 >= py0)
            {
                const int c0 = px0 / kSubW, c1 = px1 / kSubW, r0 = py0 / kSubH, r1 = py1 / kSubH;
                const uint32_t colBits = ((1u << (c1 - c0 + 1)) - 1u) << c0;
                for (int r = r0; r <= r1; r++) m |= colBits << (r * kSubCols);
            }
        }
        s.mask[buf][tid] = m;
    };

    uint32_t hi = range.y;
    if (hi > range.x) fetch(hi);
    int buf = 0;
    while (hi > range.x)
    {
        const uint32_t cnt = min((uint32_t)kCompChunk, hi - range.x);
        stage(buf);
        // one barrier per chunk: publishes the staged chunk and tells whether every pixel of the tile is finished
        if (__syncthreads_and(T == 0.0f)) break;
        hi -= cnt;
        if (hi > range.x) fetch(hi);

        if (__any_sync(0xffffffffu, T != 0.0f))
        {
            for (uint32_t base = 0; base < cnt; base += 32)
            {
                const uint32_t m = (base + lane < cnt) ? s.mask[buf][base + lane] : 0u;
                uint32_t hits = __ballot_sync(0xffffffffu, (m >> warp) & 1u);
                while (hits)
                {
                    const uint32_t j = base + __ffs(hits) - 1;
                    hits &= hits - 1;
                    const float4 a = s.a[buf][j], b = s.b[buf][j];
                    const float blue = s.c[buf][j].x;
                    const float dx = fx - a.x, dy = fy - a.y;
                    const float pw = fmaf(dx, fmaf(a.z, dx, a.w * dy), fmaf(b.x * dy, dy, b.y));
                    const float w = ex2_approx(pw);
                    const float tw = (w > (1.0f / 256.0f)) ? T * w : 0.0f;  // discard iff w <= 1/256 (splat_frag.glsl:38)
                    cr = fmaf(tw, b.z, cr);
                    cg = fmaf(tw, b.w, cg);
                    cb = fmaf(tw, blue, cb);
                    T = T - tw;
                }
                // transmittance early-out, checked once per 32 list entries (T only decreases)
                if (T < earlyOutT) T = 0.0f;
            }
        }
        buf ^= 1;
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

cudaError_t launch_composite(const uint2* ranges, const uint32_t* sortedVals, int idStrideWords, const ProjRecord* rec, void* framebuffer,
                             int fbFormat, int width, int height, int tilesX, int tilesY, float earlyOutT, cudaStream_t stream)
{
    const int tiles = tilesX * tilesY;
    if (tiles <= 0) return cudaSuccess;
    if (fbFormat == 1)
        k_composite<true><<<tiles, kCompThreads, 0, stream>>>(ranges, sortedVals, idStrideWords, rec, framebuffer, width, height, tilesX, earlyOutT);
    else
        k_composite<false><<<tiles, kCompThreads, 0, stream>>>(ranges, sortedVals, idStrideWords, rec, framebuffer, width, height, tilesX, earlyOutT);
    return cudaGetLastError();
}

}  // namespace splat
