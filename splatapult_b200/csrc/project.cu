// project.cu -- P: cull + project + SH radiance + 2D conic + (tile|depth) key emission, one fused pass.
//
// Replaces, in ONE kernel over the gaussians in storage order (fully coalesced, no index indirection):
//   K1 shader/presort_compute.glsl:31-57   frustum cull, depth key (the reference appends through a global
//                                          atomic and then stalls on a counter read-back, splatrenderer.cpp:195-204)
//   K5 shader/splat_vert.glsl:153-222      view transform, Jacobian, Sigma' = JW Sigma JW^T, +0.3 low-pass,
//                                          screen centre, SH radiance (:51-127)
//   K6 shader/splat_geom.glsl:22-54        2x2 inverse, guard-band cull.  The 3.5-sigma quad (:56-104) becomes
//                                          a conservative tile rectangle.
// Three consumers of the same per-gaussian function project_gaussian():
//   k_project_keys (default path)  one (u32 depthKey, id) pair + tile rectangle per gaussian; after the depth digits
//                                  are sorted on the gaussians, k_emit_sorted expands the sorted list into
//                                  (u32 tile id, id) instances in depth order (scan + look-back + coalesced emission)
//   k_project      (flat path)     projection + scan + emission of (tile << 32 | depthKey, id) in gaussian order
//   k_project_bin  (binned path)   side records + per-tile instance counts
// Key slots come from a single-pass decoupled look-back scan over the tile counts, so the emission order is
// deterministic (the reference's compaction order is not) and the instance count D never leaves the device.
//
// This translation unit is compiled with -fmad=false: every a*b+c below is two IEEE roundings in source
// order, divisions and square roots are IEEE (nvcc defaults -prec-div/-prec-sqrt=true, -ftz=false), so the
// projected records equal oracle/splat_oracle.c:project_one() bit for bit.
// Frame counters (visible, usedInstances, ...) take ONE atomic per CTA: a per-warp atomicAdd on one L2 address was,
// at N/32 serialised operations, the whole run time of the projection kernel until it was found.
#include <cstdlib>
#include <mutex>

#include "device_common.cuh"
#include "kernels.h"

namespace splat
{

constexpr int kProjThreads = 256;
constexpr unsigned long long kScanAggregate = 1ull << 62;
constexpr unsigned long long kScanPrefix = 2ull << 62;
constexpr unsigned long long kScanValueMask = (1ull << 62) - 1;

#define M4(m, c, r) ((m)[(c) * 4 + (r)])
#define M3(m, c, r) ((m)[(c) * 3 + (r)])

__device__ __forceinline__ void xform_point(const float* m, float x, float y, float z, float* o)
{
#pragma unroll
    for (int r = 0; r < 4; r++) o[r] = ((M4(m, 0, r) * x + M4(m, 1, r) * y) + M4(m, 2, r) * z) + M4(m, 3, r);
}
__device__ __forceinline__ void xform_vec4(const float* m, const float* v, float* o)
{
#pragma unroll
    for (int r = 0; r < 4; r++) o[r] = ((M4(m, 0, r) * v[0] + M4(m, 1, r) * v[1]) + M4(m, 2, r) * v[2]) + M4(m, 3, r) * v[3];
}
__device__ __forceinline__ void mat3_mul(const float* a, const float* b, float* out)
{
    float t[9];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int r = 0; r < 3; r++) t[c * 3 + r] = (M3(a, 0, r) * M3(b, c, 0) + M3(a, 1, r) * M3(b, c, 1)) + M3(a, 2, r) * M3(b, c, 2);
#pragma unroll
    for (int k = 0; k < 9; k++) out[k] = t[k];
}

__device__ __forceinline__ float srgb_to_linear(float s)
{
    if (s <= 0.04045f) return s / 12.92f;
    return powf((s + 0.055f) / 1.055f, 2.4f);
}

// splat_vert.glsl:51-127; sh planes hold vec4 blocks {r,g,b}_sh{0..3}
template <int SH_MODE>
__device__ __forceinline__ void radiance(float vx, float vy, float vz, float* rgb, const float* dcPre, const float4* shPre)
{
    float b[16];
    const float vx2 = vx * vx, vy2 = vy * vy, vz2 = vz * vz;
    b[0] = 0.28209479177387814f;
    const float k1 = 0.4886025119029199f;
    b[1] = -k1 * vy;
    b[2] = k1 * vz;
    b[3] = -k1 * vx;
    if (SH_MODE == 3)
    {
        const float k2 = 1.0925484305920792f, k3 = 0.31539156525252005f, k4 = 0.5462742152960396f;
        b[4] = k2 * vy * vx;
        b[5] = -k2 * vy * vz;
        b[6] = k3 * (3.0f * vz2 - 1.0f);
        b[7] = -k2 * vx * vz;
        b[8] = k4 * (vx2 - vy2);
        const float k5 = 0.5900435899266435f, k6 = 2.8906114426405543f, k7 = 0.4570457994644658f;
        const float k8 = 0.37317633259011546f, k9 = 1.4453057213202771f;
        b[9] = -k5 * vy * (3.0f * vx2 - vy2);
        b[10] = k6 * vy * vx * vz;
        b[11] = -k7 * vy * (5.0f * vz2 - 1.0f);
        b[12] = k8 * vz * (5.0f * vz2 - 3.0f);
        b[13] = -k7 * vx * (5.0f * vz2 - 1.0f);
        b[14] = k9 * vz * (vx2 - vy2);
        b[15] = -k5 * vx * (vx2 - 3.0f * vy2);
    }
#pragma unroll
    for (int c = 0; c < 3; c++)
    {
        float acc;
        if (SH_MODE == 0)
        {
            // degree-1 slots are exact zeros in this layout: b[k]*0 contributes +-0 and cannot change acc
            acc = b[0] * dcPre[c];  // prefetched with the position
        }
        else if (SH_MODE == 1)
        {
            const float4 s0 = shPre[c];
            acc = b[0] * s0.x;
            acc = acc + b[1] * s0.y;
            acc = acc + b[2] * s0.z;
            acc = acc + b[3] * s0.w;
        }
        else
        {
            float4 s = shPre[4 * c];
            acc = b[0] * s.x;
            acc = acc + b[1] * s.y;
            acc = acc + b[2] * s.z;
            acc = acc + b[3] * s.w;
#pragma unroll
            for (int j = 1; j < 4; j++)
            {
                s = shPre[4 * c + j];
                acc = acc + b[4 * j + 0] * s.x;
                acc = acc + b[4 * j + 1] * s.y;
                acc = acc + b[4 * j + 2] * s.z;
                acc = acc + b[4 * j + 3] * s.w;
            }
        }
        rgb[c] = 0.5f + acc;
    }
}

// K1 + K5 + K6 for gaussian i: writes its projected record, returns its pre-sort key and tile rectangle
struct Projected
{
    uint32_t count;   // tiles touched (0: culled, or too faint to reach the 1/256 threshold anywhere)
    uint32_t x0, y0;  // first tile column / row
    uint32_t w, h;    // rectangle size in tiles
    uint32_t refKey;  // 0xFFFFFFFF - uint(depth/far * 2^32)
    bool visible;
    bool listed;      // passed the pre-sort cull (K1 and depth < far): a member of the reference's sorted list
};

template <int SH_MODE>
__device__ __forceinline__ Projected project_gaussian(const PackedCloud& cloud, const FrameConstants& fc, ProjRecord* __restrict__ rec, uint32_t i,
                                                      bool inRange)
{
    uint32_t count = 0, rx0 = 0, ry0 = 0, rw = 0, rh = 0, refKey = 0;
    bool visible = false, listed = false;

    if (inRange)
    {
        const float4 pa = __ldg(cloud.posAlpha + i);
        const float x = pa.x, y = pa.y, z = pa.z, alpha = pa.w;
        // speculative: Sigma (and the DC colour) are fetched together with the position instead of after the cull
        // decision -- one memory round trip per gaussian instead of two (almost every gaussian survives the cull in
        // the benchmark scenes; a culled one costs 36-48 wasted bytes)
        const float4 ca = __ldg(cloud.covA + i), cb = __ldg(cloud.covB + i);
        const float cc = __ldg(cloud.covC + i);
        float dcPre[3] = {0.0f, 0.0f, 0.0f};
        if (SH_MODE == 0)
        {
#pragma unroll
            for (int c = 0; c < 3; c++) dcPre[c] = __ldg(cloud.dc[c] + i);
        }
        // same for the SH planes (48 / 192 bytes per gaussian): all loads of a gaussian are in flight together instead
        // of one dependent round trip after the other (volatile asm: the compiler would otherwise sink them to their uses)
        constexpr int kShPlanes = SH_MODE == 3 ? 12 : (SH_MODE == 1 ? 3 : 1);
        float4 shPre[kShPlanes];
        if (SH_MODE != 0)
        {
#pragma unroll
            for (int k = 0; k < kShPlanes; k++)
                asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];"
                             : "=f"(shPre[k].x), "=f"(shPre[k].y), "=f"(shPre[k].z), "=f"(shPre[k].w)
                             : "l"(cloud.sh[k] + i));
        }

        // ---- K1 presort_compute.glsl:42-56
        float p[4];
        xform_point(fc.mvp, x, y, z, p);
        const float depth = p[3];
        const float xx = p[0] / depth, yy = p[1] / depth;
        const float CLIP = 1.5f;
        bool keep = (depth > 0.0f && xx < CLIP && xx > -CLIP && yy < CLIP && yy > -CLIP);
        keep = keep && (depth < fc.zFarSort);  // uint() overflow is undefined in GLSL; the quad is far-clipped anyway
        if (keep)
        {
            const float q = (depth / fc.zFarSort) * 4294967296.0f;
            refKey = 0xFFFFFFFFu - __float2uint_rz(q);
        }
        listed = keep;
        // ---- K5/K6 transforms + guard band (splat_vert.glsl:157,200-201 ; splat_geom.glsl:46-54)
        float t[4], p4[4];
        xform_point(fc.view, x, y, z, t);
        xform_vec4(fc.proj, t, p4);
        const float ndcx = p4[0] / p4[3], ndcy = p4[1] / p4[3], ndcz = p4[2] / p4[3];
        if (ndcz < 0.25f || ndcx > 2.0f || ndcx < -2.0f || ndcy > 2.0f || ndcy < -2.0f) keep = false;

        float4 q2 = make_float4(0.0f, alpha, __uint_as_float(refKey), __uint_as_float(0u));
        if (keep)
        {
            const float WIDTH = fc.W, HEIGHT = fc.H, Z_NEAR = fc.zNear, Z_FAR = fc.zFar;
            const float SX = M4(fc.proj, 0, 0), SY = M4(fc.proj, 1, 1), WZ = M4(fc.proj, 3, 2);
            const float tzSq = t[2] * t[2];
            const float jsx = -(SX * WIDTH) / (2.0f * t[2]);
            const float jsy = -(SY * HEIGHT) / (2.0f * t[2]);
            const float jtx = (SX * t[0] * WIDTH) / (2.0f * tzSq);
            const float jty = (SY * t[1] * HEIGHT) / (2.0f * tzSq);
            const float jtz = ((Z_FAR - Z_NEAR) * WZ) / (2.0f * tzSq);
            float J[9] = {jsx, 0.0f, 0.0f, 0.0f, jsy, 0.0f, jtx, jty, jtz};  // columns (splat_vert.glsl:179-181)
            float Wm[9];
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int r = 0; r < 3; r++) M3(Wm, c, r) = M4(fc.view, c, r);
            float V[9] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w, cc};
            float JW[9], JWt[9], Vp[9];
            mat3_mul(J, Wm, JW);
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int r = 0; r < 3; r++) M3(JWt, c, r) = M3(JW, r, c);
            mat3_mul(JW, V, Vp);
            mat3_mul(Vp, JWt, Vp);
            const float m00 = M3(Vp, 0, 0) + 0.3f, m01 = M3(Vp, 0, 1), m10 = M3(Vp, 1, 0), m11 = M3(Vp, 1, 1) + 0.3f;
            const float det = m00 * m11 - m01 * m10;
            if ((det > 0.0f) && (det <= 3.4028234e38f))  // degenerate covariance: skipped (oracle does the same)
            {
                visible = true;
                const float i00 = m11 / det, i01 = -m01 / det, i10 = -m10 / det, i11 = m00 / det;
                const float gx = p4[0] / p4[3], gy = p4[1] / p4[3];
                const float px = 0.5f * (WIDTH + (gx * WIDTH) + (2.0f * 0.0f));
                const float py = 0.5f * (HEIGHT + (gy * HEIGHT) + (2.0f * 0.0f));
                const float dx = x - fc.eye[0], dy = y - fc.eye[1], dz = z - fc.eye[2];
                const float len = sqrtf((dx * dx + dy * dy) + dz * dz);
                float rgb[3];
                radiance<SH_MODE>(dx / len, dy / len, dz / len, rgb, dcPre, shPre);
                if (fc.srgb)
                {
#pragma unroll
                    for (int c = 0; c < 3; c++) rgb[c] = srgb_to_linear(rgb[c]);
                }
                float4* out = reinterpret_cast<float4*>(rec + i);
                out[0] = make_float4(px, py, i00, i01);
                out[1] = make_float4(i10, i11, rgb[0], rgb[1]);
                q2 = make_float4(rgb[2], alpha, __uint_as_float(refKey), __uint_as_float(1u));

                // ---- tile rectangle (not in the reference: stands in for the 3.5-sigma quad + rasteriser).
                // A pixel contributes iff alpha*exp(-Q/2) > 1/256 (splat_frag.glsl:38-41)  <=>  Q < 2 ln(256 alpha);
                // the axis-aligned extent of {Q <= r2} is sqrt(r2 * Sigma_xx), sqrt(r2 * Sigma_yy).  Inflated so
                // rounding can never drop a contributing pixel centre.
                const float r2 = 2.0f * logf(256.0f * alpha);
                if (r2 > 0.0f)
                {
                    const float ex = sqrtf(r2 * m00) * 1.001f + 0.01f;
                    const float ey = sqrtf(r2 * m11) * 1.001f + 0.01f;
                    // pixel centres x+0.5 in [px-ex, px+ex]
                    const float fx0 = fmaxf(ceilf(px - ex - 0.5f), 0.0f), fx1 = fminf(floorf(px + ex - 0.5f), (float)(fc.width - 1));
                    const float fy0 = fmaxf(ceilf(py - ey - 0.5f), 0.0f), fy1 = fminf(floorf(py + ey - 0.5f), (float)(fc.height - 1));
                    if (fx1 >= fx0 && fy1 >= fy0)  // false for NaN too
                    {
                        const uint32_t tx0 = (uint32_t)fx0 >> kTileShiftX, tx1 = (uint32_t)fx1 >> kTileShiftX;
                        const uint32_t ty0 = (uint32_t)fy0 >> kTileShiftY, ty1 = (uint32_t)fy1 >> kTileShiftY;
                        const uint32_t w = tx1 - tx0 + 1, h = ty1 - ty0 + 1;
                        count = w * h;
                        rx0 = tx0; ry0 = ty0; rw = w; rh = h;
                    }
                }
            }
        }
        reinterpret_cast<float4*>(rec + i)[2] = q2;  // {b, alpha, key, visible}: always written
    }

    Projected out;
    out.count = count; out.x0 = rx0; out.y0 = ry0; out.w = rw; out.h = rh; out.refKey = refKey; out.visible = visible;
    out.listed = listed;
    return out;
}

// "Flat" LSD path: projects gaussian i, scans the tile counts (decoupled look-back) and emits its keys in index order;
// the histogram it leaves for the sorter is that of the keys' LOWEST digit (weighted by the tile count).
template <int SH_MODE>
__global__ void __launch_bounds__(kProjThreads)
k_project(const PackedCloud cloud, const __grid_constant__ FrameConstants fc, ProjRecord* __restrict__ rec,
          unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals, FrameCounters* __restrict__ ctr,
          unsigned long long* __restrict__ scanState, uint32_t capacity, uint32_t* __restrict__ digit0Hist)
{
    __shared__ uint32_t sBlockId;
    __shared__ uint32_t sWarpTotal[kProjThreads / 32];
    __shared__ unsigned long long sBlockBase;
    __shared__ uint32_t sDigit0[kRadixSize];  // histogram of the emitted keys' next sort digit (saves the sorter a pass)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sBlockId = atomicAdd(&ctr->projectTicket, 1u);  // dynamic id: predecessors are resident or done
    sDigit0[tid] = 0;
    __syncthreads();
    const uint32_t vbid = sBlockId;
    uint32_t i = vbid * kProjThreads + tid;  // gaussian id
    uint32_t count = 0, refKey = 0, rect = 0;
    bool visible = false;
    {
        const Projected pg = project_gaussian<SH_MODE>(cloud, fc, rec, i, i < cloud.n);
        count = pg.count;
        refKey = pg.refKey;
        rect = pg.x0 | (pg.y0 << 10) | ((pg.w ? pg.w - 1 : 0) << 20);
        visible = pg.visible;
    }

    // ---- block-wide exclusive scan of the tile counts
    uint32_t incl = count;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) sWarpTotal[warp] = incl;
    const uint32_t nVisibleWarp = __popc(__ballot_sync(0xffffffffu, visible));
    __syncthreads();
    uint32_t warpBase = 0, blockTotal = 0;
#pragma unroll
    for (int w = 0; w < kProjThreads / 32; w++)
    {
        const uint32_t t = sWarpTotal[w];
        if (w < warp) warpBase += t;
        blockTotal += t;
    }

    // ---- decoupled look-back over block aggregates (warp 0)
    if (warp == 0)
    {
        unsigned long long excl = 0;
        if (lane == 0) st_relaxed_u64(scanState + vbid, (vbid == 0 ? kScanPrefix : kScanAggregate) | blockTotal);
        if (vbid > 0)
        {
            int base = (int)vbid - 1;
            for (;;)
            {
                const int j = base - lane;
                unsigned long long s = kScanPrefix;  // virtual predecessor of block 0: prefix 0
                if (j >= 0)
                {
                    do { s = ld_relaxed_u64(scanState + j); } while ((s >> 62) == 0);
                }
                const bool isPrefix = (s >> 62) == 2;
                const uint32_t pmask = __ballot_sync(0xffffffffu, isPrefix);
                const int firstP = pmask ? (__ffs(pmask) - 1) : 32;
                unsigned long long v = (lane <= firstP) ? (s & kScanValueMask) : 0ull;
#pragma unroll
                for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                excl += v;
                if (pmask) break;
                base -= 32;
            }
            if (lane == 0) st_relaxed_u64(scanState + vbid, kScanPrefix | (excl + blockTotal));
        }
        if (lane == 0)
        {
            sBlockBase = excl;
            if (vbid == gridDim.x - 1) ctr->totalInstances = excl + blockTotal;
        }
    }
    __syncthreads();
    // frame counters: one atomic per CTA (same-address atomics serialise in L2), through sWarpTotal, which is free now
    if (lane == 0) sWarpTotal[warp] = nVisibleWarp;
    __syncthreads();
    if (tid == 0)
    {
        uint32_t vis = 0;
#pragma unroll
        for (int w = 0; w < kProjThreads / 32; w++) vis += sWarpTotal[w];
        if (vis) atomicAdd(&ctr->visible, vis);
    }
    __syncthreads();

    // ---- warp-cooperative, coalesced key emission
    const unsigned long long gOff64 = sBlockBase + warpBase + (incl - count);
    const bool fits = gOff64 + count <= (unsigned long long)capacity;
    const uint32_t cntE = fits ? count : 0u;
    if (!fits) atomicOr(&ctr->overflow, 1u);
    const uint32_t gOff = fits ? (uint32_t)gOff64 : 0u;
    // flat path: all cntE keys of this gaussian share their low 32 bits, one weighted update covers them
    if (cntE && digit0Hist) smem_red_add(&sDigit0[refKey & (kRadixSize - 1)], cntE);

    uint32_t inclE = cntE;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inclE, d);
        if (lane >= d) inclE += v;
    }
    const uint32_t wexE = inclE - cntE;
    const uint32_t T = __shfl_sync(0xffffffffu, inclE, 31);

    uint32_t usedEnd = cntE ? gOff + cntE : 0u;
#pragma unroll
    for (int d = 16; d; d >>= 1) usedEnd = max(usedEnd, __shfl_xor_sync(0xffffffffu, usedEnd, d));
    if (lane == 0) sWarpTotal[warp] = usedEnd;

    const uint32_t tilesX = (uint32_t)fc.tilesX;
    for (uint32_t j0 = 0; j0 < T; j0 += 32)
    {
        const uint32_t j = j0 + lane;
        int owner = 0;
#pragma unroll
        for (int s = 16; s; s >>= 1)
        {
            const int cand = owner + s;
            const uint32_t v = __shfl_sync(0xffffffffu, wexE, cand);
            if (v <= j) owner = cand;
        }
        const uint32_t oBase = __shfl_sync(0xffffffffu, wexE, owner);
        const uint32_t oOff = __shfl_sync(0xffffffffu, gOff, owner);
        const uint32_t oRect = __shfl_sync(0xffffffffu, rect, owner);
        const uint32_t oKey = __shfl_sync(0xffffffffu, refKey, owner);
        const uint32_t oId = __shfl_sync(0xffffffffu, i, owner);
        if (j < T)
        {
            const uint32_t k = j - oBase;
            const uint32_t w = (oRect >> 20) + 1, x0 = oRect & 1023u, y0 = (oRect >> 10) & 1023u;
            const uint32_t row = k / w, col = k - row * w;
            const uint32_t tile = (y0 + row) * tilesX + (x0 + col);
            keys[oOff + k] = ((unsigned long long)tile << 32) | oKey;
            vals[oOff + k] = oId;
        }
    }
    __syncthreads();
    if (tid == 0)
    {
        uint32_t m = 0;
#pragma unroll
        for (int w = 0; w < kProjThreads / 32; w++) m = max(m, sWarpTotal[w]);
        if (m) atomicMax(&ctr->usedInstances, m);
    }
    if (digit0Hist)
    {
        const uint32_t c = sDigit0[tid];
        if (c) atomicAdd(digit0Hist + tid, c);
    }
}

// ---------------------------------------------------------------------------------------------------
// default LSD path, step 2: expand the depth-sorted gaussian list into (tile id, gaussian id) instances, in depth
// order -- the instances do not carry the depth key: the tile-digit passes that follow are stable.  Same scan + look-back + warp-cooperative emission as k_project, but every thread carries
// kEmitRounds entries (entry j = cta * 1024 + round * 256 + tid, so loads stay coalesced and the emission order is
// the list order): a quarter of the CTAs means a quarter of the links in the look-back chain.  With one entry per
// thread the walkers (32 predecessors per L2 round trip) were slower than the rate at which new CTAs start, so the
// walks kept growing and 7 of 8 warps sat on the barrier behind warp 0 (ncu: barrier stall 21 warps per issue).
// ---------------------------------------------------------------------------------------------------
constexpr int kEmitRounds = 4;

__global__ void __launch_bounds__(kProjThreads)
k_emit_sorted(const uint32_t* sortedVals, const uint2* side,
              uint32_t nSorted, uint32_t tilesX, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
              FrameCounters* __restrict__ ctr, unsigned long long* __restrict__ scanState, uint32_t capacity, uint32_t* __restrict__ tileDigitHist)
{
    constexpr int kWarps = kProjThreads / 32;
    __shared__ uint32_t sBlockId;
    __shared__ uint32_t sWarpTotal[kEmitRounds][kWarps];
    __shared__ unsigned long long sBlockBase;
    __shared__ uint32_t sDigit0[2][kRadixSize];  // histograms of both tile digits (bits 0-7 and 8-15 of the tile id), one count per instance

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sBlockId = atomicAdd(&ctr->projectTicket, 1u);  // dynamic id: predecessors are resident or done
    sDigit0[0][tid] = 0;
    sDigit0[1][tid] = 0;
    pdl_wait();     // the depth-sorted list is the last depth pass's output
    pdl_trigger();  // the first tile pass may be scheduled as this grid retires
    __syncthreads();
    const uint32_t vbid = sBlockId;

    uint32_t count[kEmitRounds], rect[kEmitRounds], id[kEmitRounds], incl[kEmitRounds];
#pragma unroll
    for (int r = 0; r < kEmitRounds; r++)
    {
        const uint32_t j = (vbid * kEmitRounds + r) * kProjThreads + tid;
        count[r] = 0; rect[r] = 0; id[r] = 0;
        if (j < nSorted)
        {
            // plain loads, not the read-only path (__ldg): ptxas is free to move LDG.CONSTANT loads above pdl_wait()
            id[r] = sortedVals[j];
            const uint2 sd = side[id[r]];
            const uint32_t w = sd.y & 0xffffu, h = sd.y >> 16;
            count[r] = w * h;
            rect[r] = (sd.x & 0xffffu) | ((sd.x >> 16) << 10) | ((w ? w - 1 : 0) << 20);
        }
    }
#pragma unroll
    for (int r = 0; r < kEmitRounds; r++)
    {
        uint32_t v = count[r];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const uint32_t u = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += u;
        }
        incl[r] = v;
        if (lane == 31) sWarpTotal[r][warp] = v;
    }
    __syncthreads();
    uint32_t base[kEmitRounds];  // CTA-relative offset of this warp's first instance of round r
    uint32_t blockTotal = 0;
#pragma unroll
    for (int r = 0; r < kEmitRounds; r++)
    {
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kWarps; w++)
        {
            const uint32_t t = sWarpTotal[r][w];
            if (w < warp) before += t;
            total += t;
        }
        base[r] = blockTotal + before;
        blockTotal += total;
    }

    // ---- decoupled look-back over CTA aggregates (warp 0)
    if (warp == 0)
    {
        unsigned long long excl = 0;
        if (lane == 0) st_relaxed_u64(scanState + vbid, (vbid == 0 ? kScanPrefix : kScanAggregate) | blockTotal);
        if (vbid > 0)
        {
            int b0 = (int)vbid - 1;
            for (;;)
            {
                const int j = b0 - lane;
                unsigned long long s = kScanPrefix;  // virtual predecessor of block 0: prefix 0
                if (j >= 0)
                {
                    do { s = ld_relaxed_u64(scanState + j); } while ((s >> 62) == 0);
                }
                const bool isPrefix = (s >> 62) == 2;
                const uint32_t pmask = __ballot_sync(0xffffffffu, isPrefix);
                const int firstP = pmask ? (__ffs(pmask) - 1) : 32;
                unsigned long long v = (lane <= firstP) ? (s & kScanValueMask) : 0ull;
#pragma unroll
                for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                excl += v;
                if (pmask) break;
                b0 -= 32;
            }
            if (lane == 0) st_relaxed_u64(scanState + vbid, kScanPrefix | (excl + blockTotal));
        }
        if (lane == 0)
        {
            sBlockBase = excl;
            if (vbid == gridDim.x - 1) ctr->totalInstances = excl + blockTotal;
        }
    }
    __syncthreads();

    // ---- warp-cooperative, coalesced emission, one round at a time
    // (measured alternatives, round 2: every lane writing its own entry's instances serially -- a third of the instructions,
    // but scattered partial-sector stores: 103 vs 99 us for sort+emit at 1 M, 302 vs 316 at 6 M; the same with the warp's run
    // staged in shared memory and copied out coalesced: 110 / 323 us.  The kernel's time is its scan chain, not the emission.)
    const unsigned long long blockBase = sBlockBase;
    uint32_t usedEnd = 0;
#pragma unroll
    for (int r = 0; r < kEmitRounds; r++)
    {
        const unsigned long long gOff64 = blockBase + base[r] + (incl[r] - count[r]);
        const bool fits = gOff64 + count[r] <= (unsigned long long)capacity;
        const uint32_t cntE = fits ? count[r] : 0u;
        if (!fits) atomicOr(&ctr->overflow, 1u);
        const uint32_t gOff = fits ? (uint32_t)gOff64 : 0u;
        if (cntE) usedEnd = max(usedEnd, gOff + cntE);

        uint32_t inclE = cntE;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inclE, d);
            if (lane >= d) inclE += v;
        }
        const uint32_t wexE = inclE - cntE;
        const uint32_t T = __shfl_sync(0xffffffffu, inclE, 31);
        for (uint32_t j0 = 0; j0 < T; j0 += 32)
        {
            const uint32_t j = j0 + lane;
            int owner = 0;
#pragma unroll
            for (int s = 16; s; s >>= 1)
            {
                const int cand = owner + s;
                const uint32_t v = __shfl_sync(0xffffffffu, wexE, cand);
                if (v <= j) owner = cand;
            }
            const uint32_t oBase = __shfl_sync(0xffffffffu, wexE, owner);
            const uint32_t oOff = __shfl_sync(0xffffffffu, gOff, owner);
            const uint32_t oRect = __shfl_sync(0xffffffffu, rect[r], owner);
            const uint32_t oId = __shfl_sync(0xffffffffu, id[r], owner);
            if (j < T)
            {
                const uint32_t k = j - oBase;
                const uint32_t w = (oRect >> 20) + 1, x0 = oRect & 1023u, y0 = (oRect >> 10) & 1023u;
                const uint32_t row = k / w, col = k - row * w;
                const uint32_t tile = (y0 + row) * tilesX + (x0 + col);
                keys[oOff + k] = tile;
                vals[oOff + k] = oId;
                if (tileDigitHist)
                {
                    smem_red_add(&sDigit0[0][tile & (kRadixSize - 1)], 1u);
                    smem_red_add(&sDigit0[1][(tile >> kRadixBits) & (kRadixSize - 1)], 1u);
                }
            }
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) usedEnd = max(usedEnd, __shfl_xor_sync(0xffffffffu, usedEnd, d));
    // one atomic per CTA on the frame counter, not one per warp: same-address atomics serialise in L2
    if (lane == 0) sWarpTotal[0][warp] = usedEnd;  // sWarpTotal is free again: every warp passed the barrier after the scan
    __syncthreads();
    if (tid == 0)
    {
        uint32_t m = 0;
#pragma unroll
        for (int w = 0; w < kWarps; w++) m = max(m, sWarpTotal[0][w]);
        if (m) atomicMax(&ctr->usedInstances, m);
    }
    if (tileDigitHist)
    {
        const uint32_t c0 = sDigit0[0][tid], c1 = sDigit0[1][tid];
        if (c0) atomicAdd(tileDigitHist + tid, c0);
        if (c1) atomicAdd(tileDigitHist + kRadixSize + tid, c1);
    }
}

// ---------------------------------------------------------------------------------------------------
// default LSD path, step 1: project every gaussian and leave ONE (depthKey, id) pair per gaussian for the sorter
// plus an 8-byte side record with its tile rectangle.  Gaussians that touch no tile get depthKey 0xFFFFFFFF (they
// sort to the end and expand to nothing).  No scan, no atomics besides the per-CTA histogram flush.
// ---------------------------------------------------------------------------------------------------
template <int SH_MODE>
__global__ void __launch_bounds__(kProjThreads)
k_project_keys(const PackedCloud cloud, const __grid_constant__ FrameConstants fc, ProjRecord* __restrict__ rec,
               uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, uint2* __restrict__ side, FrameCounters* __restrict__ ctr,
               uint32_t* __restrict__ digit0Hist)
{
    // grid-stride over chunks of 256 gaussians: the CTA keeps its digit histogram and its visible count across chunks
    // and flushes them once.  (With one CTA per chunk and one atomicAdd per warp on ctr->visible, that single L2
    // address took N/32 serialised atomics: the kernel's time was their time -- 58 us at 1 M, 553 us at 6 M gaussians.)
    // histograms of ALL FOUR depth digits: the keys are in registers here, so the depth passes of the sorter neither need a
    // histogram pass nor accumulate the next digit's histogram while they scatter
    __shared__ uint32_t sDigit[kDepthBits / kRadixBits][kRadixSize];
    const int tid = threadIdx.x;
    pdl_trigger();  // first kernel of the frame's chain (launched normally, after the control-block memset): the first depth
                    // pass may be scheduled as this grid's CTAs retire; it waits for this grid before it reads anything
#pragma unroll
    for (int p = 0; p < kDepthBits / kRadixBits; p++) sDigit[p][tid] = 0;
    __syncthreads();
    uint32_t visibleMine = 0;
    const uint32_t numChunks = (cloud.n + kProjThreads - 1) / kProjThreads;
    for (uint32_t chunk = blockIdx.x; chunk < numChunks; chunk += gridDim.x)
    {
        const uint32_t i = chunk * kProjThreads + tid;
        const bool inRange = i < cloud.n;
        const Projected pg = project_gaussian<SH_MODE>(cloud, fc, rec, i, inRange);
        if (inRange)
        {
            side[i] = pg.count ? make_uint2(pg.x0 | (pg.y0 << 16), pg.w | (pg.h << 16)) : make_uint2(0u, 0u);
            if (fc.stereoPass != 2)
            {
                // shared sort: the order must hold for an eye that may draw what this one does not
                const bool sorted = fc.stereoPass == 1 ? pg.listed : pg.count != 0;
                const uint32_t key = sorted ? pg.refKey : 0xFFFFFFFFu;
                keys[i] = key;
                vals[i] = i;
#pragma unroll
                for (int p = 0; p < kDepthBits / kRadixBits; p++) smem_red_add(&sDigit[p][(key >> (p * kRadixBits)) & (kRadixSize - 1)], 1u);
            }
        }
        visibleMine += pg.visible ? 1u : 0u;
    }
    if (fc.stereoPass != 2 && blockIdx.x == 0 && tid == 0) ctr->sortCount = cloud.n;
    // CTA total of the visible counts: warp reduce, then one atomic per warp into shared memory, one global atomic
    __shared__ uint32_t sVisible;
    if (tid == 0) sVisible = 0;
    __syncthreads();  // also: every thread's histogram updates are done
#pragma unroll
    for (int d = 16; d; d >>= 1) visibleMine += __shfl_xor_sync(0xffffffffu, visibleMine, d);
    if ((tid & 31) == 0 && visibleMine) atomicAdd(&sVisible, visibleMine);
    __syncthreads();
    if (tid == 0 && sVisible) atomicAdd(&ctr->visible, sVisible);
#pragma unroll
    for (int p = 0; p < kDepthBits / kRadixBits; p++)
    {
        const uint32_t c = sDigit[p][tid];
        if (c) atomicAdd(digit0Hist + p * kRadixSize + tid, c);  // all zero in a render-only pass
    }
}

// ---------------------------------------------------------------------------------------------------
// binned (MSD-by-tile) variant: same projection, but instead of emitting keys it leaves a 16-byte side record
// {x0|y0<<16, w|h<<16, depthKey, count} per gaussian and counts the instances of every tile with one RED per
// instance; key slots are handed out later, per tile (binsort.cu).
// ---------------------------------------------------------------------------------------------------
template <int SH_MODE>
__global__ void __launch_bounds__(kProjThreads)
k_project_bin(const PackedCloud cloud, const __grid_constant__ FrameConstants fc, ProjRecord* __restrict__ rec, uint4* __restrict__ side,
              uint32_t* __restrict__ tileCount, FrameCounters* __restrict__ ctr)
{
    const int lane = threadIdx.x & 31;
    const uint32_t i = blockIdx.x * kProjThreads + threadIdx.x;
    const bool inRange = i < cloud.n;
    const Projected pg = project_gaussian<SH_MODE>(cloud, fc, rec, i, inRange);
    if (inRange) side[i] = make_uint4(pg.x0 | (pg.y0 << 16), pg.w | (pg.h << 16), pg.refKey, pg.count);

    const uint32_t nVisibleWarp = __popc(__ballot_sync(0xffffffffu, pg.visible));
    uint32_t incl = pg.count;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    const uint32_t wex = incl - pg.count;
    const uint32_t T = __shfl_sync(0xffffffffu, incl, 31);
    // frame counters: one pair of atomics per CTA, not per warp (same-address atomics serialise in L2)
    __shared__ uint32_t sVis[kProjThreads / 32], sTot[kProjThreads / 32];
    if (lane == 0)
    {
        sVis[threadIdx.x >> 5] = nVisibleWarp;
        sTot[threadIdx.x >> 5] = T;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        uint32_t vis = 0;
        unsigned long long tot = 0;
#pragma unroll
        for (int w = 0; w < kProjThreads / 32; w++)
        {
            vis += sVis[w];
            tot += sTot[w];
        }
        if (vis) atomicAdd(&ctr->visible, vis);
        if (tot) atomicAdd(&ctr->totalInstances, tot);
    }
    const uint32_t rectXY = pg.x0 | (pg.y0 << 16);
    const uint32_t tilesX = (uint32_t)fc.tilesX;
    for (uint32_t j0 = 0; j0 < T; j0 += 32)
    {
        const uint32_t j = j0 + lane;
        int owner = 0;
#pragma unroll
        for (int s = 16; s; s >>= 1)
        {
            const int cand = owner + s;
            const uint32_t v = __shfl_sync(0xffffffffu, wex, cand);
            if (v <= j) owner = cand;
        }
        const uint32_t oBase = __shfl_sync(0xffffffffu, wex, owner);
        const uint32_t oXY = __shfl_sync(0xffffffffu, rectXY, owner);
        const uint32_t oW = __shfl_sync(0xffffffffu, pg.w, owner);
        if (j < T)
        {
            const uint32_t k = j - oBase;
            const uint32_t row = k / oW, col = k - row * oW;
            const uint32_t tile = ((oXY >> 16) + row) * tilesX + ((oXY & 0xffffu) + col);
            atomicAdd(tileCount + tile, 1u);  // result unused: compiles to RED
        }
    }
}

// scatter: every instance takes the next free slot of its tile's segment (order inside a segment is arbitrary
// here; the per-tile sort orders by (depthKey, id), which is total, so the final lists are deterministic)
__global__ void __launch_bounds__(kProjThreads)
k_bin_instances(const uint4* __restrict__ side, uint32_t n, int tilesX, uint32_t* __restrict__ tileCursor, unsigned long long* __restrict__ pairs,
                uint32_t capacity)
{
    const int lane = threadIdx.x & 31;
    const uint32_t i = blockIdx.x * kProjThreads + threadIdx.x;
    uint4 sd = make_uint4(0, 0, 0, 0);
    if (i < n) sd = __ldg(side + i);
    const uint32_t count = sd.w;
    uint32_t incl = count;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    const uint32_t wex = incl - count;
    const uint32_t T = __shfl_sync(0xffffffffu, incl, 31);
    for (uint32_t j0 = 0; j0 < T; j0 += 32)
    {
        const uint32_t j = j0 + lane;
        int owner = 0;
#pragma unroll
        for (int s = 16; s; s >>= 1)
        {
            const int cand = owner + s;
            const uint32_t v = __shfl_sync(0xffffffffu, wex, cand);
            if (v <= j) owner = cand;
        }
        const uint32_t oBase = __shfl_sync(0xffffffffu, wex, owner);
        const uint32_t oXY = __shfl_sync(0xffffffffu, sd.x, owner);
        const uint32_t oW = __shfl_sync(0xffffffffu, sd.y, owner) & 0xffffu;
        const uint32_t oKey = __shfl_sync(0xffffffffu, sd.z, owner);
        const uint32_t oId = __shfl_sync(0xffffffffu, i, owner);
        if (j < T)
        {
            const uint32_t k = j - oBase;
            const uint32_t row = k / oW, col = k - row * oW;
            const uint32_t tile = ((oXY >> 16) + row) * (uint32_t)tilesX + ((oXY & 0xffffu) + col);
            const uint32_t slot = atomicAdd(tileCursor + tile, 1u);
            if (slot < capacity) pairs[slot] = ((unsigned long long)oKey << 32) | oId;
        }
    }
}

cudaError_t launch_project_bin(const PackedCloud& cloud, const FrameConstants& fc, ProjRecord* rec, uint4* side, uint32_t* tileCount,
                               FrameCounters* ctr, cudaStream_t stream)
{
    const uint32_t blocks = project_num_blocks(cloud.n);
    if (blocks == 0) return cudaSuccess;
    switch (cloud.shMode)
    {
    case 0: k_project_bin<0><<<blocks, kProjThreads, 0, stream>>>(cloud, fc, rec, side, tileCount, ctr); break;
    case 1: k_project_bin<1><<<blocks, kProjThreads, 0, stream>>>(cloud, fc, rec, side, tileCount, ctr); break;
    default: k_project_bin<3><<<blocks, kProjThreads, 0, stream>>>(cloud, fc, rec, side, tileCount, ctr); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_bin_instances(const uint4* side, uint32_t n, int tilesX, uint32_t* tileCursor, unsigned long long* pairs, uint32_t capacity,
                                 cudaStream_t stream)
{
    const uint32_t blocks = project_num_blocks(n);
    if (blocks == 0) return cudaSuccess;
    k_bin_instances<<<blocks, kProjThreads, 0, stream>>>(side, n, tilesX, tileCursor, pairs, capacity);
    return cudaGetLastError();
}

cudaError_t launch_project(const PackedCloud& cloud, const FrameConstants& fc, ProjRecord* rec, unsigned long long* keys, uint32_t* vals,
                           FrameCounters* ctr, unsigned long long* scanState, uint32_t capacity, uint32_t* digit0Hist, cudaStream_t stream)
{
    const uint32_t blocks = project_num_blocks(cloud.n);
    if (blocks == 0) return cudaSuccess;
    switch (cloud.shMode)
    {
    case 0: k_project<0><<<blocks, kProjThreads, 0, stream>>>(cloud, fc, rec, keys, vals, ctr, scanState, capacity, digit0Hist); break;
    case 1: k_project<1><<<blocks, kProjThreads, 0, stream>>>(cloud, fc, rec, keys, vals, ctr, scanState, capacity, digit0Hist); break;
    default: k_project<3><<<blocks, kProjThreads, 0, stream>>>(cloud, fc, rec, keys, vals, ctr, scanState, capacity, digit0Hist); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_project_keys(const PackedCloud& cloud, const FrameConstants& fc, ProjRecord* rec, uint32_t* keys, uint32_t* vals,
                                uint2* side, FrameCounters* ctr, uint32_t* digit0Hist, cudaStream_t stream)
{
    uint32_t blocks = project_num_blocks(cloud.n);
    if (blocks == 0) return cudaSuccess;
    // persistent, grid-stride: as many CTAs as are resident at once (a multiple of the SM count), times `waves`
    // (SPL_PROJECT_WAVES overrides, for tuning)
    // per device (the SM count and the occupancy are the device's), filled once under a lock: renderers on several GPUs or
    // threads of one process share this table
    static std::mutex cacheMutex;
    static int ctasPerWaveCache[64][4] = {};
    const int mode = cloud.shMode == 0 ? 0 : (cloud.shMode == 1 ? 1 : 3);
    int dev = 0;
    cudaGetDevice(&dev);
    int ctasPerWave[4] = {0, 0, 0, 0};
    {
        std::lock_guard<std::mutex> lock(cacheMutex);
        int& slot = ctasPerWaveCache[dev & 63][mode];
        if (!slot)
        {
            int numSMs = 0, occ = 0;
            cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, dev);
            if (mode == 0) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_project_keys<0>, kProjThreads, 0);
            else if (mode == 1) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_project_keys<1>, kProjThreads, 0);
            else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_project_keys<3>, kProjThreads, 0);
            slot = (numSMs > 0 ? numSMs : 148) * (occ > 0 ? occ : 4);
        }
        ctasPerWave[mode] = slot;
    }
    // measured: one wave is best for the light kernels (34 us vs 40 / 43 us with 2 / 4 waves at 1 M gaussians, deg 0); the
    // degree-3 kernel (89 registers, 2 CTAs per SM) balances its tail better with 4 (302 vs 311 us at 6 M)
    static const int wavesEnv = getenv("SPL_PROJECT_WAVES") ? atoi(getenv("SPL_PROJECT_WAVES")) : 0;
    const int waves = wavesEnv > 0 ? wavesEnv : (mode == 3 ? 4 : 1);
    blocks = min(blocks, (uint32_t)(ctasPerWave[mode] * waves));
    switch (cloud.shMode)
    {
    case 0: k_project_keys<0><<<blocks, kProjThreads, 0, stream>>>(cloud, fc, rec, keys, vals, side, ctr, digit0Hist); break;
    case 1: k_project_keys<1><<<blocks, kProjThreads, 0, stream>>>(cloud, fc, rec, keys, vals, side, ctr, digit0Hist); break;
    default: k_project_keys<3><<<blocks, kProjThreads, 0, stream>>>(cloud, fc, rec, keys, vals, side, ctr, digit0Hist); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_emit_sorted(const uint32_t* sortedVals, const uint2* side, uint32_t n,
                               const FrameConstants& fc, uint32_t* keys, uint32_t* vals, FrameCounters* ctr,
                               unsigned long long* scanState, uint32_t capacity, uint32_t* tileDigitHist, cudaStream_t stream, bool pdl)
{
    const uint32_t perCta = kProjThreads * kEmitRounds;
    const uint32_t blocks = (n + perCta - 1) / perCta;  // <= project_num_blocks(n): scanState is large enough
    if (blocks == 0) return cudaSuccess;
    return launch_kernel(k_emit_sorted, dim3(blocks), dim3(kProjThreads), 0, stream, pdl, sortedVals, side, n, (uint32_t)fc.tilesX, keys, vals, ctr, scanState,
                         capacity, tileDigitHist);
}

uint32_t project_num_blocks(uint32_t n) { return (n + kProjThreads - 1) / kProjThreads; }

}  // namespace splat
