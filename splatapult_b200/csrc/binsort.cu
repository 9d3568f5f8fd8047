// binsort.cu -- S (frame path): MSD radix sort of the (tile | depthKey) instance keys.
//
// Same contract as the reference's sort stage (src/splatrenderer.cpp:206-264: stable ascending key order, i.e.
// back to front, payload = gaussian index) applied to the packed (tile << 32 | depthKey) key, but organised
// most-significant-digit first so that the instance list crosses HBM once instead of once per 8-bit digit:
//   1. the top digit IS the tile id.  The projection kernel has already counted the instances of every tile
//      (one RED each); k_tile_scan turns the counts into segment offsets = the tile ranges the compositor needs;
//   2. k_bin_instances (project.cu) scatters every instance into its tile's segment -- a counting sort on the
//      whole tile id in ONE pass (8160 bins at 1080p); the segment array (34 MB at 1M gaussians) stays L2-resident;
//   3. the remaining 32 depth bits are sorted per segment in SHARED MEMORY by k_segment_sort_*: the same stable
//      8-bit LSD scheme as the device-wide sorter (per-warp digit counters, warp match/popc ranks, 4 passes), the
//      segment ping-ponging between two shared buffers; runs of equal depth keys are then ordered by gaussian id,
//      which is exactly the order a stable sort of the id-ordered emission would give -- so both frame paths
//      produce bit-identical lists although the scatter in step 2 fills segments in arbitrary order.
//      Segments too long for shared memory (> 12288 instances in one 16x16 tile) ping-pong through global memory.
// The stand-alone LSD onesweep sorter (radix_sort.cu) remains selectable (SPL_OPT_SORT_MODE) and is what
// spl_sorter_* exposes.
#include <mutex>

#include "device_common.cuh"
#include "kernels.h"

namespace splat
{

constexpr int kScanThreads = 1024;

__global__ void __launch_bounds__(kScanThreads)
k_tile_scan(const uint32_t* __restrict__ tileCount, uint32_t numTiles, uint2* __restrict__ ranges, uint32_t* __restrict__ tileCursor,
            uint32_t* __restrict__ largeList, FrameCounters* __restrict__ ctr, uint32_t capacity)
{
    __shared__ unsigned long long sWarp[kScanThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t per = (numTiles + kScanThreads - 1) / kScanThreads;
    const uint32_t t0 = min(numTiles, tid * per), t1 = min(numTiles, t0 + per);
    unsigned long long sum = 0;
    for (uint32_t t = t0; t < t1; t++) sum += tileCount[t];
    unsigned long long incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) sWarp[warp] = incl;
    __syncthreads();
    unsigned long long base = 0, total = 0;
    for (int w = 0; w < kScanThreads / 32; w++)
    {
        const unsigned long long v = sWarp[w];
        if (w < warp) base += v;
        total += v;
    }
    unsigned long long off = base + incl - sum;
    for (uint32_t t = t0; t < t1; t++)
    {
        const uint32_t c = tileCount[t];
        const uint32_t b = (uint32_t)min(off, (unsigned long long)capacity);
        const uint32_t e = (uint32_t)min(off + c, (unsigned long long)capacity);
        ranges[t] = make_uint2(b, e);
        tileCursor[t] = (uint32_t)min(off, 0xffffffffull);
        if (e - b > (uint32_t)kSegSortSmallCap) largeList[atomicAdd(&ctr->largeTiles, 1u)] = t;
        off += c;
    }
    if (tid == 0)
    {
        ctr->usedInstances = (uint32_t)min(total, (unsigned long long)capacity);
        if (total > capacity) ctr->overflow = 1u;
    }
}

// ---------------------------------------------------------------------------------------------------
// per-segment sort: stable LSD radix on the high word (depthKey) of u64 (depthKey << 32 | id), then id order
// inside runs of equal depthKey.  NWARPS warps; item k of the segment belongs to warp (k / 32) % NWARPS ... in
// the "warp-striped" arrangement used throughout: warp w owns items [w*per, (w+1)*per), lane-interleaved, so the
// (item, lane) visiting order is the memory order and ranks are stable.
// ---------------------------------------------------------------------------------------------------
template <int NWARPS>
struct SegCounters
{
    uint32_t warpCount[NWARPS][kRadixSize];
    uint32_t scanTmp[NWARPS];
};

// one stable counting pass src -> dst on digit (hi32 >> shift) & 255.  Ranks are computed twice (count phase,
// scatter phase) instead of being kept per item, so segments of any length need no per-item registers.
template <int NWARPS>
__device__ __forceinline__ void segment_radix_pass(const unsigned long long* src, unsigned long long* dst, uint32_t n, int shift,
                                                   SegCounters<NWARPS>& c, int tid)
{
    constexpr int NT = NWARPS * 32;
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t perWarp = ((n + NT - 1) / NT) * 32;  // items per warp, multiple of 32
    const uint32_t wBegin = warp * perWarp;

    for (int d = tid; d < NWARPS * kRadixSize; d += NT) (&c.warpCount[0][0])[d] = 0;
    __syncthreads();
    // ---- count
    for (uint32_t k = 0; k < perWarp; k += 32)
    {
        const uint32_t idx = wBegin + k + lane;
        const bool ok = idx < n;
        const uint32_t dgt = ok ? ((uint32_t)(src[idx] >> (32 + shift)) & 255u) : (0x10000u | lane);
        const uint32_t peers = __match_any_sync(0xffffffffu, dgt);
        if (ok && (peers & lanemask_lt()) == 0) c.warpCount[warp][dgt] += __popc(peers);  // one lane per distinct digit
        __syncwarp();
    }
    __syncthreads();
    // ---- digit totals -> exclusive bases (thread d < 256 owns digit d)
    uint32_t total = 0;
    if (tid < kRadixSize)
    {
#pragma unroll 4
        for (int w = 0; w < NWARPS; w++)
        {
            const uint32_t v = c.warpCount[w][tid];
            c.warpCount[w][tid] = total;
            total += v;
        }
    }
    uint32_t incl = total;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (tid < kRadixSize && lane == 31) c.scanTmp[warp] = incl;
    __syncthreads();
    if (tid < kRadixSize)
    {
        uint32_t base = 0;
        for (int w = 0; w < kRadixSize / 32; w++)
            if (w < warp) base += c.scanTmp[w];
        const uint32_t excl = base + incl - total;
        for (int w = 0; w < NWARPS; w++) c.warpCount[w][tid] += excl;
    }
    __syncthreads();
    // ---- scatter (same visiting order => same ranks)
    for (uint32_t k = 0; k < perWarp; k += 32)
    {
        const uint32_t idx = wBegin + k + lane;
        const bool ok = idx < n;
        const unsigned long long v = ok ? src[idx] : 0ull;
        const uint32_t dgt = ok ? ((uint32_t)(v >> (32 + shift)) & 255u) : (0x10000u | lane);
        const uint32_t peers = __match_any_sync(0xffffffffu, dgt);
        const uint32_t before = __popc(peers & lanemask_lt());
        uint32_t prev = 0;
        if (ok) prev = c.warpCount[warp][dgt];
        __syncwarp();
        if (ok)
        {
            dst[prev + before] = v;
            if (before == 0) c.warpCount[warp][dgt] = prev + __popc(peers);
        }
        __syncwarp();
    }
    __syncthreads();
}

// runs of equal depthKey are put in ascending id order (= the order of a stable sort of id-ordered input)
__device__ __forceinline__ void segment_fix_ties(unsigned long long* data, uint32_t n, int tid, int nthreads)
{
    for (uint32_t i = tid; i + 1 < n; i += nthreads)
    {
        const uint32_t key = (uint32_t)(data[i] >> 32);
        if ((i == 0 || (uint32_t)(data[i - 1] >> 32) != key) && (uint32_t)(data[i + 1] >> 32) == key)
        {
            uint32_t end = i + 2;
            while (end < n && (uint32_t)(data[end] >> 32) == key) end++;
            for (uint32_t a = i + 1; a < end; a++)  // insertion sort; runs are short
            {
                const unsigned long long v = data[a];
                uint32_t b = a;
                while (b > i && data[b - 1] > v) { data[b] = data[b - 1]; b--; }
                data[b] = v;
            }
        }
    }
}

template <int NWARPS>
__device__ __forceinline__ void segment_sort(unsigned long long* a, unsigned long long* b, uint32_t n, SegCounters<NWARPS>& c, int tid)
{
    // 4 x 8 bits of depthKey; an even number of passes ends in `a`
    segment_radix_pass<NWARPS>(a, b, n, 0, c, tid);
    segment_radix_pass<NWARPS>(b, a, n, 8, c, tid);
    segment_radix_pass<NWARPS>(a, b, n, 16, c, tid);
    segment_radix_pass<NWARPS>(b, a, n, 24, c, tid);
    segment_fix_ties(a, n, tid, NWARPS * 32);
    __syncthreads();
}

constexpr int kSegSmallWarps = 8;
constexpr int kSegSmallThreads = kSegSmallWarps * 32;

struct SegSmallSmem
{
    unsigned long long buf[2][kSegSortSmallCap];
    SegCounters<kSegSmallWarps> c;
};

__global__ void __launch_bounds__(kSegSmallThreads)
k_segment_sort_small(unsigned long long* __restrict__ pairs, const uint2* __restrict__ ranges)
{
    __shared__ SegSmallSmem s;
    const uint2 r = ranges[blockIdx.x];
    const uint32_t n = r.y - r.x;
    if (n < 2 || n > (uint32_t)kSegSortSmallCap) return;
    for (uint32_t i = threadIdx.x; i < n; i += kSegSmallThreads) s.buf[0][i] = pairs[r.x + i];
    __syncthreads();
    segment_sort<kSegSmallWarps>(s.buf[0], s.buf[1], n, s.c, threadIdx.x);
    for (uint32_t i = threadIdx.x; i < n; i += kSegSmallThreads) pairs[r.x + i] = s.buf[0][i];
}

constexpr int kSegLargeWarps = 32;
constexpr int kSegLargeThreads = kSegLargeWarps * 32;

__global__ void __launch_bounds__(kSegLargeThreads)
k_segment_sort_large(unsigned long long* __restrict__ pairs, unsigned long long* __restrict__ scratch, const uint2* __restrict__ ranges,
                     const uint32_t* __restrict__ largeList, const FrameCounters* __restrict__ ctr)
{
    extern __shared__ __align__(16) unsigned char sLargeRaw[];
    SegCounters<kSegLargeWarps>& c = *reinterpret_cast<SegCounters<kSegLargeWarps>*>(sLargeRaw);
    unsigned long long* bufA = reinterpret_cast<unsigned long long*>(sLargeRaw + ((sizeof(SegCounters<kSegLargeWarps>) + 15) & ~size_t(15)));
    unsigned long long* bufB = bufA + kSegSortLargeCap;
    const uint32_t nLarge = ctr->largeTiles;
    for (uint32_t li = blockIdx.x; li < nLarge; li += gridDim.x)
    {
        const uint2 r = ranges[largeList[li]];
        const uint32_t n = r.y - r.x;
        if (n <= (uint32_t)kSegSortLargeCap)
        {
            for (uint32_t i = threadIdx.x; i < n; i += kSegLargeThreads) bufA[i] = pairs[r.x + i];
            __syncthreads();
            segment_sort<kSegLargeWarps>(bufA, bufB, n, c, threadIdx.x);
            for (uint32_t i = threadIdx.x; i < n; i += kSegLargeThreads) pairs[r.x + i] = bufA[i];
        }
        else
        {
            // ping-pong through global memory; `scratch` mirrors the pairs array index for index
            segment_sort<kSegLargeWarps>(pairs + r.x, scratch + r.x, n, c, threadIdx.x);
        }
        __syncthreads();
    }
}

static constexpr size_t kSegLargeSmemBytes = ((sizeof(SegCounters<kSegLargeWarps>) + 15) & ~size_t(15)) + 2 * (size_t)kSegSortLargeCap * sizeof(unsigned long long);

cudaError_t launch_tile_scan(const uint32_t* tileCount, uint32_t numTiles, uint2* ranges, uint32_t* tileCursor, uint32_t* largeList,
                             FrameCounters* ctr, uint32_t capacity, cudaStream_t stream)
{
    k_tile_scan<<<1, kScanThreads, 0, stream>>>(tileCount, numTiles, ranges, tileCursor, largeList, ctr, capacity);
    return cudaGetLastError();
}

cudaError_t launch_segment_sort(unsigned long long* pairs, unsigned long long* scratch, const uint2* ranges, uint32_t numTiles,
                                const uint32_t* largeList, const FrameCounters* ctr, int numSMs, cudaStream_t stream)
{
    static std::mutex attrMutex;  // cudaFuncSetAttribute is per device: one table, filled under a lock
    static bool attrSet[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(attrMutex);
    if (!attrSet[dev & 63])
    {
        cudaError_t e = cudaFuncSetAttribute(k_segment_sort_large, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSegLargeSmemBytes);
        if (e != cudaSuccess) return e;
        attrSet[dev & 63] = true;
    }
    if (numTiles == 0) return cudaSuccess;
    k_segment_sort_small<<<numTiles, kSegSmallThreads, 0, stream>>>(pairs, ranges);
    k_segment_sort_large<<<numSMs, kSegLargeThreads, kSegLargeSmemBytes, stream>>>(pairs, scratch, ranges, largeList, ctr);
    return cudaGetLastError();
}

}  // namespace splat
