// composite.cu -- tile ranges + C: per-tile front-to-back alpha compositing into a float framebuffer.
//
// Replaces the reference's draw: glDrawElements(GL_POINTS, sortCount) (src/splatrenderer.cpp:338) ->
// fixed-function rasteriser -> shader/splat_frag.glsl:18-42 (K7: g = exp(-1/2 d^T Sigma2^-1 d), w = alpha g,
// discard iff w <= 1/256, premultiplied output) -> ROP blend GL_ONE, GL_ONE_MINUS_SRC_ALPHA over a target
// cleared to (0,0,0,1) (K8, src/app.cpp:144-164), strictly back to front.
//
// One CTA owns one 32x16-pixel tile: 8 warps, each an 8x8 pixel block, TWO pixels per thread (lane l: column l & 7, rows
// l >> 3 and (l >> 3) + 4 of the block).  It walks the tile's sorted list.  The list is stored back-to-front (ascending
// reference key); the CTA walks it from its END, i.e. front to back, accumulating
//     C += T * w * rgb ;  T *= (1 - w)
// which is the closed form of the reference's "over" chain up to float reassociation, and
// lets a pixel stop once its transmittance T can no longer change the result by more than earlyOutT
// (a bounded optimisation, earlyOutT = 0 reproduces the reference's full chain).
//
// The kernel is instruction-issue bound (ncu: profiles/), not HBM bound, so the design minimises issued
// instructions per (splat, pixel):
//   * entries are staged 256 at a time into shared memory as the coefficients of log2(weight) in the pixel's offset
//     (u, v) from the tile centre; the gather of the next chunk's records is issued before the current chunk is
//     evaluated (software pipeline, one barrier per chunk);
//   * while staging, each entry gets an 8-bit mask of the 8x8 blocks its discard ellipse reaches (exact); a warp
//     ballots the masks of 32 entries at a time and evaluates only its hits -- warp-uniform, no divergence;
//   * the two pixels of a thread are evaluated with Blackwell's packed FP32 pipe (fma.rn.f32x2 -> SASS FFMA2, whose
//     operands may each be a scalar broadcast): log2 w = [c0 + u (c1 + hA u)] + v (c2 + hB u + hC v) is 2 FFMA + 3 FFMA2 for both
//     pixels (they share u), the weight and the three colour accumulators 1 FMUL2 + 3 FFMA2 + 1 FADD2 -- a 24-instruction
//     hit body per 64 pixels where one pixel per thread took 21 per 32; the walk (FLO, BMSK, LOP3, LEA, 3 LDS, BRA) is
//     paid once per pair.  8x8 blocks are hit by 0.61x as many (entry, block) pairs as 8x4 ones (benchmarks/hit_model.py).
// Pixel centres are at (x+0.5, y+0.5), row 0 is the bottom row (GL window coordinates).
#include <cuda_fp16.h>

#include "device_common.cuh"
#include "kernels.h"

namespace splat
{

constexpr int kCompThreads = 256;                    // 8 warps; two pixels per thread
constexpr int kCompChunk = kCompThreads;             // list entries staged per round (one per thread)
constexpr int kSubW = 8, kSubH = 8;                  // a warp's pixel block
constexpr int kSubCols = kTileW / kSubW;             // 4
constexpr int kSubRows = kTileH / kSubH;             // 2
static_assert(kSubCols * kSubRows * 32 == kCompThreads, "one warp per 8x8 block");
static_assert(kSubCols * kSubRows <= 8, "the sub-block mask has 8 bits");

// ---- packed FP32 pairs (Blackwell: one FFMA2 / FMUL2 / FADD2 does two independent fp32 operations; every operand may be a
// scalar broadcast in the SASS encoding, which the compiler uses when both halves of a pair are the same register)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2 bcast2(float s) { return pack2(s, s); }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// ---------------------------------------------------------------------------------------------------
// tile ranges: [begin, end) of each tile's run in the sorted key array
// ---------------------------------------------------------------------------------------------------
template <typename KeyT>
__global__ void k_tile_ranges(const KeyT* __restrict__ keys, int keyShift, const uint32_t* __restrict__ nDevice, uint2* __restrict__ ranges,
                              uint32_t numTiles)
{
    const uint32_t n = *nDevice;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint32_t tile = (uint32_t)(keys[i] >> keyShift);
        if (tile >= numTiles) continue;  // cannot happen for keys this library emits
        if (i == 0 || (uint32_t)(keys[i - 1] >> keyShift) != tile) ranges[tile].x = i;
        if (i == n - 1 || (uint32_t)(keys[i + 1] >> keyShift) != tile) ranges[tile].y = i + 1;
    }
}

cudaError_t launch_tile_ranges(const void* sortedKeys, int keyBytes, const uint32_t* nDevice, uint64_t capacity, uint2* ranges, uint32_t numTiles,
                               int numSMs, cudaStream_t stream)
{
    cudaError_t e = cudaMemsetAsync(ranges, 0, (size_t)numTiles * sizeof(uint2), stream);
    if (e != cudaSuccess) return e;
    const uint32_t blocks = (uint32_t)max(1ull, min((unsigned long long)numSMs * 8ull, (unsigned long long)((capacity + 255) / 256)));
    if (keyBytes == 4) k_tile_ranges<uint32_t><<<blocks, 256, 0, stream>>>(static_cast<const uint32_t*>(sortedKeys), 0, nDevice, ranges, numTiles);
    else k_tile_ranges<unsigned long long><<<blocks, 256, 0, stream>>>(static_cast<const unsigned long long*>(sortedKeys), 32, nDevice, ranges, numTiles);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// composite
// ---------------------------------------------------------------------------------------------------
struct CompSmem
{
    float4 a[2][kCompChunk];    // c0, c1, c2, hA   with  log2 w = c0 + c1 u + c2 v + hA u^2 + hB uv + hC v^2  (stage())
    float4 b[2][kCompChunk];    // hB, hC, r, g
    float4 c[2][kCompChunk];    // b, -, -, -       (same index stride as a/b: one address computation per entry)
    uint32_t hits[2][kCompChunk / 32][8];  // per group of 32 entries and per warp: the entries whose ellipse reaches that warp's
                                           // 8x8 block (bit k = entry slot k of the group)
};

// 5 CTAs per SM (48 registers): measured 2-3 % faster than 6 CTAs at 40 registers with spills (171 vs 175 us at 1 M, 266 vs
// 275 at 6 M); two hits per loop trip (interleaved weight chains) was slower at either occupancy (181 / 189 us).
template <bool HALF>
__global__ void __launch_bounds__(kCompThreads, 5)
k_composite(const uint2* ranges, const uint32_t* vals, int idStride, const ProjRecord* rec,
            void* __restrict__ fb, int width, int height, int tilesX, float earlyOutT, int cullMode, uint32_t beginXor, FrameReport report)
{
    __shared__ CompSmem s;
    pdl_wait();  // tile ranges, sorted ids and records are the preceding kernels' output (no-op in a normal launch)

    // frame report (renderer.cu: pollReports): the instance count this frame needed, tagged with the frame's sequence number,
    // posted to a word of mapped host memory -- a fire-and-forget PCIe write, so the host learns about a capacity overflow on
    // the asynchronous paths without a synchronising read-back.  The counters are final: every producer ran before this kernel.
    if (report.hostWord && blockIdx.x == 0 && threadIdx.x == 0)
    {
        const unsigned long long total = report.counters->totalInstances;
        *reinterpret_cast<volatile unsigned long long*>(report.hostWord) = report.seqTag | (total < kReportCountMask ? total : kReportCountMask);
    }

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int tileX = tile % tilesX, tileY = tile / tilesX;
    const int subCol = warp % kSubCols, subRow = warp / kSubCols;
    // this thread's two pixels: same column, rows 4 apart inside the warp's 8x8 block
    const int x = tileX * kTileW + subCol * kSubW + (lane & (kSubW - 1));
    const int y0 = tileY * kTileH + subRow * kSubH + (lane / kSubW), y1 = y0 + kSubH / 2;
    const bool inside0 = x < width && y0 < height, inside1 = x < width && y1 < height;
    const float tileX0 = (float)(tileX * kTileW), tileY0 = (float)(tileY * kTileH);
    const float tileCx = tileX0 + 0.5f * kTileW, tileCy = tileY0 + 0.5f * kTileH;
    // gl_FragCoord.xy = (x + 0.5, y + 0.5), kept relative to the tile centre (|u| <= 15.5, |v| <= 7.5: exact in float)
    const float pu = ((float)x + 0.5f) - tileCx;
    const f32x2 pv = pack2(((float)y0 + 0.5f) - tileCy, ((float)y1 + 0.5f) - tileCy);

    uint2 range = ranges[tile];
    range.x ^= beginXor;  // ranges written by the last sort pass store ~begin (kernels.h: SortWorkspace::tileRanges)
    // T == 0 marks a finished pixel: every later contribution is multiplied by it
    f32x2 T = pack2(inside0 ? 1.0f : 0.0f, inside1 ? 1.0f : 0.0f);
    f32x2 cr = pack2(0.0f, 0.0f), cg = cr, cb = cr;
    auto live = [&]() {
        float t0, t1;
        unpack2(T, t0, t1);
        return t0 != 0.0f || t1 != 0.0f;
    };

    // software pipeline: the records of chunk k+1 are in flight while chunk k is evaluated
    float4 q0, q1, q2;
    bool have = false;
    auto fetch = [&](uint32_t hi) {
        const uint32_t cnt = min((uint32_t)kCompChunk, hi - range.x);
        have = (uint32_t)tid < cnt;
        if (have)
        {
            // plain loads (not __ldg / the read-only path): ptxas may move LDG.CONSTANT loads above pdl_wait()
            const uint32_t id = vals[(size_t)(hi - 1 - tid) * idStride];  // nearest first
            const float4* p = reinterpret_cast<const float4*>(rec + id);
            q0 = p[0];  // px, py, i00, i01
            q1 = p[1];  // i10, i11, r, g
            q2 = p[2];  // b, alpha, key, visible
        }
    };
    auto stage = [&](int buf) {
        // a warp evaluates its hits highest ballot bit first (FLO, no bit reversal): each group of 32 entries is
        // stored in reverse lane order, so the highest bit is the nearest entry
        const int sidx = tid ^ 31;
        uint32_t m = 0u;  // sub-block mask of this thread's entry: bit (row * kSubCols + col) = the warp that owns the block
        if (have)
        {
        // log2 w as a polynomial in the pixel's offset (u, v) from the TILE centre: with d0 = tile centre - splat centre,
        //   log2 w = hA (dx0+u)^2 + hB (dx0+u)(dy0+v) + hC (dy0+v)^2 + L = c0 + c1 u + c2 v + hA u^2 + hB uv + hC v^2
        // evaluated per pixel pair as [c0 + u (c1 + hA u)] + v (c2 + hB u + hC v): 2 FFMA + 3 FFMA2; the rounding error is that of the
        // direct form (same term magnitudes).
        const float kLog2e = 1.4426950408889634f;
        const float h = -0.5f * kLog2e;
        const float hA = h * q0.z, hB = h * (q0.w + q1.x), hC = h * q1.y;
        const float dx0 = tileCx - q0.x, dy0 = tileCy - q0.y;
        const float c0 = fmaf(dx0, fmaf(hA, dx0, hB * dy0), fmaf(hC * dy0, dy0, log2f(q2.y)));
        const float c1 = fmaf(2.0f * hA, dx0, hB * dy0);
        const float c2 = fmaf(2.0f * hC, dy0, hB * dx0);
        s.a[buf][sidx] = make_float4(c0, c1, c2, hA);
        s.b[buf][sidx] = make_float4(hB, hC, q1.z, q1.w);
        s.c[buf][sidx] = make_float4(q2.x, 0.0f, 0.0f, 0.0f);
        // Sub-block mask: which 8x8 pixel blocks can hold a pixel with alpha exp(-Q/2) > 1/256, i.e. a pixel centre
        // inside the ellipse Q(d) = A dx^2 + 2B dx dy + C dy^2 < r2, r2 = 2 ln(256 alpha) (inflated so rounding in
        // the evaluation below cannot put a contributing pixel outside).  A culled lane would have been discarded
        // anyway (tw = 0 leaves the accumulators bit-unchanged): the image does not depend on the mask.
        const float A = q0.z, B = 0.5f * (q0.w + q1.x), Cc = q1.y;
        // det = A C - B^2 by Kahan's difference of products: elongated splats cancel almost completely
        const float bb = B * B;
        const float det = fmaf(A, Cc, -bb) + fmaf(-B, B, bb);
        const float r2 = 2.0f * 0.6931471805599453f * (8.0f + __log2f(q2.y)) * 1.002f + 0.004f;
        m = (1u << (kSubCols * kSubRows)) - 1u;
        // not a bounded ellipse as stored (or NaN): evaluate everywhere
        if (cullMode != 0 && det > 0.0f && A > 0.0f && r2 == r2 && fabsf(q0.x) < 1e30f && fabsf(q0.y) < 1e30f)
        {
            m = 0u;
            if (r2 > 0.0f)
            {
                const float eps = 0.01f;
                const float rdet = __fdividef(r2, det);
                const float ex = sqrt_approx(rdet * Cc) * 1.001f, ey = sqrt_approx(rdet * A) * 1.001f;  // half extents
                const float ox = tileX0 - q0.x, oy = tileY0 - q0.y;  // tile origin relative to the splat centre
                if (cullMode == 1)
                {
                    // footprint box only
                    for (int r = 0; r < kSubRows; r++)
                    {
                        const float d0 = oy + (float)(r * kSubH) + 0.5f, d1 = d0 + (float)(kSubH - 1);
                        if (d0 - eps > ey || d1 + eps < -ey) continue;
                        for (int c = 0; c < kSubCols; c++)
                        {
                            const float c0 = ox + (float)(c * kSubW) + 0.5f, c1 = c0 + (float)(kSubW - 1);
                            if (c0 - eps <= ex && c1 + eps >= -ex) m |= 1u << (r * kSubCols + c);
                        }
                    }
                }
                else
                {
                    // exact: the ellipse cut by the row band dy in [d0, d1] is convex, so it meets a block iff its
                    // x-extent [xl, xr] over the band overlaps the block's columns.  x_right(dy) = -(B/A) dy +
                    // sqrt(r2 A - det dy^2)/A is concave with its maximum at the ellipse's rightmost point
                    // dy* = -(B/C) ex: clamp dy* into the band and evaluate; x_left mirrors it.
                    const float rA = __fdividef(1.0f, A);
                    const float slope = -B * rA;
                    const float tStar = -B * __fdividef(ex, Cc);
                    const float r2A = r2 * A * 1.002f;
#pragma unroll
                    for (int r = 0; r < kSubRows; r++)
                    {
                        const float d0 = oy + (float)(r * kSubH) + 0.5f - eps, d1 = d0 + (float)(kSubH - 1) + 2.0f * eps;
                        const float lo = fmaxf(d0, -ey), hi = fminf(d1, ey);
                        if (lo > hi) continue;
                        const float tr = fminf(fmaxf(tStar, lo), hi), tl = fminf(fmaxf(-tStar, lo), hi);
                        const float xr = fmaf(slope, tr, sqrt_approx(fmaxf(fmaf(-det * tr, tr, r2A), 0.0f)) * rA) + eps;
                        const float xl = fmaf(slope, tl, -sqrt_approx(fmaxf(fmaf(-det * tl, tl, r2A), 0.0f)) * rA) - eps;
#pragma unroll
                        for (int c = 0; c < kSubCols; c++)
                        {
                            const float c0 = ox + (float)(c * kSubW) + 0.5f, c1 = c0 + (float)(kSubW - 1);
                            if (c0 <= xr && c1 >= xl) m |= 1u << (r * kSubCols + c);
                        }
                    }
                }
            }
        }
        }
        // transpose the 32 entry masks of this warp's group into one hit word per evaluating warp: the walk below then reads
        // ONE word per group instead of balloting 32 masks in each of the 8 warps.  Slot k of the group was staged by lane
        // k ^ 31, hence the bit reversal.
        uint32_t mine = 0u;
#pragma unroll
        for (int wq = 0; wq < kSubCols * kSubRows; wq++)
        {
            const uint32_t word = __ballot_sync(0xffffffffu, (m >> wq) & 1u);
            if (lane == wq) mine = word;
        }
        if (lane < kSubCols * kSubRows) s.hits[buf][warp][lane] = __brev(mine);
    };

    const f32x2 pu2 = bcast2(pu);
    const float kDiscard = 1.0f / 256.0f;
    uint32_t hi = range.y;
    if (hi > range.x) fetch(hi);
    int buf = 0;
    while (hi > range.x)
    {
        const uint32_t cnt = min((uint32_t)kCompChunk, hi - range.x);
        stage(buf);
        // one barrier per chunk: publishes the staged chunk and tells whether every pixel of the tile is finished
        if (__syncthreads_and(!live())) break;
        hi -= cnt;
        if (hi > range.x) fetch(hi);

        if (__any_sync(0xffffffffu, live()))
        {
            // 32-bit shared addresses, advanced per group of 32 entries (keeps the walk out of 64-bit address math)
            uint32_t hitAddr = (uint32_t)__cvta_generic_to_shared(&s.hits[buf][0][warp]);
            uint32_t groupAddr = (uint32_t)__cvta_generic_to_shared(&s.a[buf][0]);
            const uint32_t hitEnd = hitAddr + ((cnt + 31u) >> 5) * 32u;
            for (; hitAddr < hitEnd; hitAddr += 32, groupAddr += 32 * 16)
            {
                uint32_t hits;  // slots past the list's end hold 0
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(hits) : "r"(hitAddr));
                // the index of the NEXT hit is found (FLO runs on the slow XU pipe) while the current one is evaluated
                uint32_t top;
                asm("bfind.u32 %0, %1;" : "=r"(top) : "r"(hits));  // index of the highest set bit (0xffffffff for 0: unused then)
                while (hits)
                {
                    // hit body, two pixels per thread: BMSK, LOP3, LEA, FLO for the walk; 3 LDS; 2 FFMA + 3 FFMA2 (log2 w of both
                    // pixels), 2 MUFU.EX2, 2 FSETP + 2 FSEL (discard), FMUL2, 3 FFMA2, FADD2; BRA
                    uint32_t below;
                    asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(below) : "r"(0), "r"(top));   // bits [0, top)
                    hits &= below;
                    const uint32_t addr = groupAddr + top * 16u;
                    asm("bfind.u32 %0, %1;" : "=r"(top) : "r"(hits));
                    float4 a, b;
                    float blue;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(addr));
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];"
                                 : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                                 : "r"(addr), "n"(sizeof(CompSmem::a)));
                    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(blue) : "r"(addr), "n"(sizeof(CompSmem::a) + sizeof(CompSmem::b)));
                    // log2 w = [c0 + u (c1 + hA u)] + v (c2 + hB u + hC v): the bracket is the same for both pixels (they share u) --
                    // two scalar FFMA --, the rest three FFMA2; the dependent chain is three deep (t2 -> t2 -> pw), the bracket runs
                    // beside it
                    const float s1 = fmaf(a.w, pu, a.y);
                    const float s0 = fmaf(pu, s1, a.x);
                    f32x2 t2 = fma2(bcast2(b.y), pv, bcast2(a.z));
                    t2 = fma2(bcast2(b.x), pu2, t2);
                    const f32x2 pw = fma2(pv, t2, bcast2(s0));
                    float pw0, pw1;
                    unpack2(pw, pw0, pw1);
                    float w0 = ex2_approx(pw0), w1 = ex2_approx(pw1);
                    w0 = (w0 > kDiscard) ? w0 : 0.0f;  // discard iff w <= 1/256 (splat_frag.glsl:38)
                    w1 = (w1 > kDiscard) ? w1 : 0.0f;
                    const f32x2 tw = mul2(T, pack2(w0, w1));
                    cr = fma2(bcast2(b.z), tw, cr);
                    cg = fma2(bcast2(b.w), tw, cg);
                    cb = fma2(bcast2(blue), tw, cb);
                    T = sub2(T, tw);
                }
                // transmittance early-out, checked once per 32 list entries (T only decreases)
                {
                    float t0, t1;
                    unpack2(T, t0, t1);
                    if (t0 < earlyOutT) t0 = 0.0f;
                    if (t1 < earlyOutT) t1 = 0.0f;
                    T = pack2(t0, t1);
                    // every pixel of this warp finished: leave the rest of the chunk (its hits would all be multiplied by
                    // T = 0).  benchmarks/earlyout_model.py: a fifth of the kernel's hit-loop trips were such trips.
                    if (!__any_sync(0xffffffffu, t0 != 0.0f || t1 != 0.0f)) break;
                }
            }
        }
        buf ^= 1;
    }

    float r0, r1, g0, g1, b0, b1;
    unpack2(cr, r0, r1);
    unpack2(cg, g0, g1);
    unpack2(cb, b0, b1);
    auto store = [&](int yy, float r, float g, float b) {
        const size_t pix = (size_t)yy * width + x;
        if (HALF)
        {
            __half2 lo = __floats2half2_rn(r, g), hi2 = __floats2half2_rn(b, 1.0f);
            uint2 v;
            v.x = *reinterpret_cast<uint32_t*>(&lo);
            v.y = *reinterpret_cast<uint32_t*>(&hi2);
            reinterpret_cast<uint2*>(fb)[pix] = v;
        }
        else
        {
            reinterpret_cast<float4*>(fb)[pix] = make_float4(r, g, b, 1.0f);
        }
    };
    if (inside0) store(y0, r0, g0, b0);
    if (inside1) store(y1, r1, g1, b1);
}

// ---------------------------------------------------------------------------------------------------
// present: float target -> RGBA8, the reference's RenderDesktop blit (src/app.cpp:1076-1082; desktop_frag.glsl:32-37
// with color = (1,1,1,1): out = texColor) into an 8-bit window: clamp to [0,1], x255, round to nearest.  NaN -> 0.
// ---------------------------------------------------------------------------------------------------
template <bool HALF>
__global__ void k_present_rgba8(const void* __restrict__ fb, uint32_t* __restrict__ out, size_t pixels)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pixels) return;
    float4 v;
    if (HALF)
    {
        const uint2 raw = reinterpret_cast<const uint2*>(fb)[i];
        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
        const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
        v = make_float4(lo.x, lo.y, hi.x, hi.y);
    }
    else
    {
        v = reinterpret_cast<const float4*>(fb)[i];
    }
    auto to8 = [](float c) -> uint32_t { return (uint32_t)(fminf(fmaxf(c, 0.0f), 1.0f) * 255.0f + 0.5f); };  // fmaxf(NaN, 0) = 0
    out[i] = to8(v.x) | (to8(v.y) << 8) | (to8(v.z) << 16) | (to8(v.w) << 24);
}

cudaError_t launch_present_rgba8(const void* framebuffer, int fbFormat, uint32_t* out, size_t pixels, cudaStream_t stream)
{
    if (!pixels) return cudaSuccess;
    const unsigned blocks = (unsigned)((pixels + 255) / 256);
    if (fbFormat == 1) k_present_rgba8<true><<<blocks, 256, 0, stream>>>(framebuffer, out, pixels);
    else k_present_rgba8<false><<<blocks, 256, 0, stream>>>(framebuffer, out, pixels);
    return cudaGetLastError();
}

cudaError_t launch_composite(const uint2* ranges, const uint32_t* sortedVals, int idStrideWords, const ProjRecord* rec, void* framebuffer,
                             int fbFormat, int width, int height, int tilesX, int tilesY, float earlyOutT, int cullMode, uint32_t rangeBeginXor,
                             const FrameReport& report, cudaStream_t stream, bool pdl)
{
    const int tiles = tilesX * tilesY;
    if (tiles <= 0) return cudaSuccess;
    if (fbFormat == 1)
        return launch_kernel(k_composite<true>, dim3(tiles), dim3(kCompThreads), 0, stream, pdl, ranges, sortedVals, idStrideWords, rec, framebuffer, width, height, tilesX,
                             earlyOutT, cullMode, rangeBeginXor, report);
    return launch_kernel(k_composite<false>, dim3(tiles), dim3(kCompThreads), 0, stream, pdl, ranges, sortedVals, idStrideWords, rec, framebuffer, width, height, tilesX,
                         earlyOutT, cullMode, rangeBeginXor, report);
}

}  // namespace splat
