// radix_sort.cu -- S: device-wide stable LSD radix sort of (u64 key, u32 payload) pairs, 8-bit digits.
//
// Stands in for the reference's GLSL sorter: 4 x {multi_radixsort_histograms.glsl:29-54 (K2),
// multi_radixsort.glsl:53-138 (K3)} dispatched from src/splatrenderer.cpp:206-264 (and its rgc::radix_sort
// fallback, src/radix_sort.hpp:340-485).  Same contract -- stable, ascending, payload carried along -- on the
// wider (tile << 32 | depthKey) key, so one global sort yields every tile's back-to-front list.
//
// Differences in kind, not only in speed:
//   * the reference re-reduces ALL workgroup histograms in EVERY workgroup (O(WG^2) reads) and only
//     ceil(n/262144) of its workgroups see data; here every digit's global histogram is built once, up front,
//     in one pass over the keys (S1), and each scatter pass (S2) resolves its tile's offsets with a
//     decoupled look-back over per-tile status words: one read + one write of the data per pass.
//   * key tiles (4096 keys = 32 KiB, + 16 KiB payload) are staged into shared memory by the TMA engine
//     (cp.async.bulk + mbarrier, SASS UBLKCP) instead of per-thread loads; ranks come from warp match /
//     popc ("warp-scan scatter"); the tile is reordered in shared memory so global stores are digit-run coalesced.
//   * element counts are read from device memory: no host round trip anywhere.
#include "device_common.cuh"
#include "kernels.h"

namespace splat
{

constexpr int kHistThreads = 256;
constexpr int kWarpsPerSortCta = kSortThreads / 32;

__device__ __forceinline__ uint32_t digit_of(unsigned long long key, int shift, uint32_t mask) { return (uint32_t)(key >> shift) & mask; }
__host__ __device__ __forceinline__ uint32_t digit_mask(int shift, int endBit) { const int bits = endBit - shift; return bits >= kRadixBits ? (uint32_t)(kRadixSize - 1) : ((1u << bits) - 1u); }

// ---------------------------------------------------------------------------------------------------
// S1: all digit histograms in one pass; also clears the look-back status words this sort will use
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kHistThreads)
k_sort_histogram(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ nDevice, int beginBit, int endBit, int passes,
                 uint32_t* __restrict__ histograms, uint32_t* __restrict__ status, uint32_t maxTiles)
{
    extern __shared__ __align__(128) unsigned char smemRaw[];
    unsigned long long* sKeys[2] = {reinterpret_cast<unsigned long long*>(smemRaw),
                                    reinterpret_cast<unsigned long long*>(smemRaw) + kSortTileKeys};
    uint32_t* sHist = reinterpret_cast<uint32_t*>(smemRaw + 2 * kSortTileKeys * sizeof(unsigned long long));  // [passes][256]
    __shared__ __align__(8) uint64_t sBar[2];

    const int tid = threadIdx.x;
    const uint32_t n = *nDevice;
    const uint32_t numTiles = (n + kSortTileKeys - 1) / kSortTileKeys;

    // clear status rows [pass][0..numTiles) x 256
    {
        const uint32_t wordsPerPass = numTiles * kRadixSize;
        for (int p = 0; p < passes; p++)
        {
            uint4* row = reinterpret_cast<uint4*>(status + (size_t)p * maxTiles * kRadixSize);
            for (uint32_t w = blockIdx.x * kHistThreads + tid; w < wordsPerPass / 4; w += gridDim.x * kHistThreads) row[w] = make_uint4(0, 0, 0, 0);
        }
    }

    for (int k = tid; k < passes * kRadixSize; k += kHistThreads) sHist[k] = 0;
    if (tid == 0)
    {
        mbar_init(&sBar[0], 1);
        mbar_init(&sBar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](uint32_t tile, int buf) {
        const uint32_t cnt = min((uint32_t)kSortTileKeys, n - tile * kSortTileKeys);
        const uint32_t bytes = (cnt * 8u + 15u) & ~15u;
        fence_proxy_async();
        mbar_expect_tx(&sBar[buf], bytes);
        bulk_g2s(sKeys[buf], keys + (size_t)tile * kSortTileKeys, bytes, &sBar[buf]);
    };

    uint32_t phase[2] = {0, 0};
    int buf = 0;
    if (tid == 0 && blockIdx.x < numTiles) issue(blockIdx.x, 0);
    for (uint32_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x)
    {
        const uint32_t next = tile + gridDim.x;
        if (tid == 0 && next < numTiles) issue(next, buf ^ 1);
        mbar_wait(&sBar[buf], phase[buf]);
        phase[buf] ^= 1;
        const uint32_t cnt = min((uint32_t)kSortTileKeys, n - tile * kSortTileKeys);
#pragma unroll 4
        for (int k = 0; k < kSortKeysPerThread; k++)
        {
            const uint32_t idx = k * kHistThreads + tid;
            const bool ok = idx < cnt;
            const unsigned long long key = ok ? sKeys[buf][idx] : 0ull;
            const uint32_t active = __ballot_sync(0xffffffffu, ok);
            for (int p = 0; p < passes; p++)
            {
                const int sh = beginBit + p * kRadixBits;
                const uint32_t d = digit_of(key, sh, digit_mask(sh, endBit));
                if (p < 3)
                {
                    if (ok) atomicAdd(&sHist[p * kRadixSize + d], 1u);
                }
                else if (ok)
                {
                    // high digits (top depth byte, tile id) are nearly constant across a warp: aggregate
                    const uint32_t peers = __match_any_sync(active, d);
                    if ((peers & lanemask_lt()) == 0) atomicAdd(&sHist[p * kRadixSize + d], (uint32_t)__popc(peers));
                }
            }
        }
        __syncthreads();  // all reads of sKeys[buf] done before it is refilled two tiles from now
        buf ^= 1;
    }
    __syncthreads();
    for (int k = tid; k < passes * kRadixSize; k += kHistThreads)
    {
        const uint32_t c = sHist[k];
        if (c) atomicAdd(&histograms[k], c);
    }
}

// ---------------------------------------------------------------------------------------------------
// S2: one scatter pass (onesweep): rank in tile, look back for the tile's global digit offsets, scatter
// ---------------------------------------------------------------------------------------------------
struct SortSmem
{
    unsigned long long keys[kSortTileKeys];           // TMA destination; reused for the digit-ordered tile
    uint32_t vals[kSortTileKeys];                     // TMA destination; reused for the digit-ordered payloads
    uint32_t warpCount[kWarpsPerSortCta][kRadixSize]; // per-warp digit counters -> exclusive prefix over warps
    uint32_t ctaExcl[kRadixSize];                     // first local slot of each digit
    uint32_t globalBase[kRadixSize];                  // global index of local slot 0 of each digit (minus ctaExcl)
    uint32_t scanTmp[kWarpsPerSortCta];
    uint32_t tile;
    uint64_t bar;
};

__global__ void __launch_bounds__(kSortThreads, 3)
k_sort_onesweep(const unsigned long long* __restrict__ keysIn, const uint32_t* __restrict__ valsIn, unsigned long long* __restrict__ keysOut,
                uint32_t* __restrict__ valsOut, const uint32_t* __restrict__ nDevice, int shift, uint32_t mask, const uint32_t* __restrict__ histogram,
                uint32_t* __restrict__ status, uint32_t* __restrict__ ticket)
{
    extern __shared__ __align__(128) unsigned char smemRaw[];
    SortSmem& s = *reinterpret_cast<SortSmem*>(smemRaw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = *nDevice;
    const uint32_t numTiles = (n + kSortTileKeys - 1) / kSortTileKeys;

    // global exclusive scan of this pass's digit histogram: thread d owns digit d
    uint32_t digitStart;
    {
        const uint32_t h = histogram[tid];
        uint32_t incl = h;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) s.scanTmp[warp] = incl;
        if (tid == 0)
        {
            mbar_init(&s.bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
        uint32_t base = 0;
#pragma unroll
        for (int w = 0; w < kWarpsPerSortCta; w++)
            if (w < warp) base += s.scanTmp[w];
        digitStart = base + incl - h;
        __syncthreads();
    }

    uint32_t parity = 0;
    for (;;)
    {
        if (tid == 0) s.tile = atomicAdd(ticket, 1u);
        // clear per-warp counters for this tile
#pragma unroll
        for (int w = 0; w < kWarpsPerSortCta; w++) s.warpCount[w][tid] = 0;
        __syncthreads();
        const uint32_t tile = s.tile;
        if (tile >= numTiles) break;
        const uint32_t nValid = min((uint32_t)kSortTileKeys, n - tile * kSortTileKeys);

        if (tid == 0)
        {
            const uint32_t kb = (nValid * 8u + 15u) & ~15u, vb = (nValid * 4u + 15u) & ~15u;
            fence_proxy_async();
            mbar_expect_tx(&s.bar, kb + vb);
            bulk_g2s(s.keys, keysIn + (size_t)tile * kSortTileKeys, kb, &s.bar);
            bulk_g2s(s.vals, valsIn + (size_t)tile * kSortTileKeys, vb, &s.bar);
        }
        mbar_wait(&s.bar, parity);
        parity ^= 1;

        // ---- rank: warp w owns items [w*512, (w+1)*512), item j of a lane = w*512 + j*32 + lane (memory order)
        unsigned long long key[kSortKeysPerThread];
        uint16_t pos[kSortKeysPerThread];
#pragma unroll
        for (int j = 0; j < kSortKeysPerThread; j++)
        {
            const uint32_t idx = warp * (kSortKeysPerThread * 32) + j * 32 + lane;
            key[j] = idx < nValid ? s.keys[idx] : ~0ull;  // padding sorts last inside the tile and is never written
        }
#pragma unroll
        for (int j = 0; j < kSortKeysPerThread; j++)
        {
            const uint32_t d = digit_of(key[j], shift, mask);
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            const uint32_t before = __popc(peers & lanemask_lt());
            const uint32_t prev = s.warpCount[warp][d];
            __syncwarp();
            if (before == 0) s.warpCount[warp][d] = prev + __popc(peers);
            __syncwarp();
            pos[j] = (uint16_t)(prev + before);
        }
        __syncthreads();

        // ---- per-digit totals (thread d), publish, look back, local/global bases
        {
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < kWarpsPerSortCta; w++)
            {
                const uint32_t c = s.warpCount[w][tid];
                s.warpCount[w][tid] = run;
                run += c;
            }
            const uint32_t cnt = run;
            uint32_t* myStatus = status + (size_t)tile * kRadixSize + tid;
            st_relaxed_u32(myStatus, (tile == 0 ? kFlagPrefix : kFlagAggregate) | cnt);

            // exclusive scan of cnt over the 256 digits
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            if (lane == 31) s.scanTmp[warp] = incl;
            __syncthreads();
            uint32_t base = 0;
#pragma unroll
            for (int w = 0; w < kWarpsPerSortCta; w++)
                if (w < warp) base += s.scanTmp[w];
            const uint32_t excl = base + incl - cnt;

            uint32_t prior = 0;
            if (tile > 0)
            {
                for (int t = (int)tile - 1; t >= 0; t--)
                {
                    const uint32_t* st = status + (size_t)t * kRadixSize + tid;
                    uint32_t v;
                    do { v = ld_relaxed_u32(st); } while ((v & kFlagMask) == 0);
                    prior += v & kValueMask;
                    if ((v & kFlagMask) == kFlagPrefix) break;
                }
                st_relaxed_u32(myStatus, kFlagPrefix | (prior + cnt));
            }
            s.ctaExcl[tid] = excl;
            s.globalBase[tid] = digitStart + prior - excl;
        }
        __syncthreads();

        // ---- reorder the tile in shared memory: keys
#pragma unroll
        for (int j = 0; j < kSortKeysPerThread; j++)
        {
            const uint32_t d = digit_of(key[j], shift, mask);
            const uint32_t lp = s.ctaExcl[d] + s.warpCount[warp][d] + pos[j];
            pos[j] = (uint16_t)lp;
        }
        // payloads: staged (memory order) -> registers, then both to their sorted slots
        uint32_t val[kSortKeysPerThread];
#pragma unroll
        for (int j = 0; j < kSortKeysPerThread; j++) val[j] = s.vals[warp * (kSortKeysPerThread * 32) + j * 32 + lane];
        __syncthreads();  // every staged key/payload is in registers
#pragma unroll
        for (int j = 0; j < kSortKeysPerThread; j++)
        {
            s.keys[pos[j]] = key[j];
            s.vals[pos[j]] = val[j];
        }
        __syncthreads();

        // ---- coalesced scatter: slot i of digit d goes to globalBase[d] + i
#pragma unroll 4
        for (int k = 0; k < kSortKeysPerThread; k++)
        {
            const uint32_t i = k * kSortThreads + tid;
            if (i < nValid)
            {
                const unsigned long long kk = s.keys[i];
                const uint32_t g = s.globalBase[digit_of(kk, shift, mask)] + i;
                keysOut[g] = kk;
                valsOut[g] = s.vals[i];
            }
        }
        __syncthreads();  // shared tile fully consumed before the next bulk copy lands
    }
}

size_t sort_status_bytes(uint64_t capacity, int passes)
{
    const uint64_t maxTiles = (capacity + kSortTileKeys - 1) / kSortTileKeys + 1;
    return (size_t)passes * maxTiles * kRadixSize * sizeof(uint32_t);
}

static cudaError_t ensure_sort_attrs(size_t histSmem, size_t sortSmem)
{
    static bool attrSet[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attrSet[dev & 63])
    {
        cudaError_t e = cudaFuncSetAttribute(k_sort_histogram, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)histSmem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_sort_onesweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sortSmem);
        if (e != cudaSuccess) return e;
        attrSet[dev & 63] = true;
    }
    return cudaSuccess;
}

static constexpr size_t kHistSmemBytes = 2 * kSortTileKeys * sizeof(unsigned long long) + (size_t)kMaxSortPasses * kRadixSize * sizeof(uint32_t);

cudaError_t launch_sort_histogram(const unsigned long long* keysA, const uint32_t* nDevice, uint64_t capacity, int beginBit, int endBit,
                                  const SortWorkspace& ws, int numSMs, cudaStream_t stream)
{
    const int passes = (endBit - beginBit + kRadixBits - 1) / kRadixBits;
    if (passes < 1 || passes > kMaxSortPasses) return cudaErrorInvalidValue;
    cudaError_t e = ensure_sort_attrs(kHistSmemBytes, sizeof(SortSmem));
    if (e != cudaSuccess) return e;
    const uint32_t maxTiles = (uint32_t)((capacity + kSortTileKeys - 1) / kSortTileKeys + 1);
    const uint32_t capTiles = (uint32_t)((capacity + kSortTileKeys - 1) / kSortTileKeys);
    const uint32_t histGrid = max(1u, min((uint32_t)numSMs * 3u, capTiles));
    k_sort_histogram<<<histGrid, kHistThreads, kHistSmemBytes, stream>>>(keysA, nDevice, beginBit, endBit, passes, ws.histograms, ws.status, maxTiles);
    return cudaGetLastError();
}

cudaError_t launch_sort_passes(unsigned long long* keysA, uint32_t* valsA, unsigned long long* keysB, uint32_t* valsB, const uint32_t* nDevice,
                               uint64_t capacity, int beginBit, int endBit, const SortWorkspace& ws, int numSMs, cudaStream_t stream)
{
    const int passes = (endBit - beginBit + kRadixBits - 1) / kRadixBits;
    if (passes < 1 || passes > kMaxSortPasses) return cudaErrorInvalidValue;
    cudaError_t e = ensure_sort_attrs(kHistSmemBytes, sizeof(SortSmem));
    if (e != cudaSuccess) return e;
    const uint32_t maxTiles = (uint32_t)((capacity + kSortTileKeys - 1) / kSortTileKeys + 1);
    const uint32_t capTiles = (uint32_t)((capacity + kSortTileKeys - 1) / kSortTileKeys);
    const uint32_t sortGrid = max(1u, min((uint32_t)numSMs * 3u, capTiles));
    for (int p = 0; p < passes; p++)
    {
        const unsigned long long* kin = (p & 1) ? keysB : keysA;
        const uint32_t* vin = (p & 1) ? valsB : valsA;
        unsigned long long* kout = (p & 1) ? keysA : keysB;
        uint32_t* vout = (p & 1) ? valsA : valsB;
        const int shift = beginBit + p * kRadixBits;
        k_sort_onesweep<<<sortGrid, kSortThreads, sizeof(SortSmem), stream>>>(kin, vin, kout, vout, nDevice, shift, digit_mask(shift, endBit),
                                                                              ws.histograms + p * kRadixSize,
                                                                              ws.status + (size_t)p * maxTiles * kRadixSize, ws.tickets + p);
    }
    return cudaGetLastError();
}

cudaError_t launch_radix_sort(unsigned long long* keysA, uint32_t* valsA, unsigned long long* keysB, uint32_t* valsB, const uint32_t* nDevice,
                              uint64_t capacity, int beginBit, int endBit, const SortWorkspace& ws, int numSMs, cudaStream_t stream, int* launches)
{
    cudaError_t e = launch_sort_histogram(keysA, nDevice, capacity, beginBit, endBit, ws, numSMs, stream);
    if (e != cudaSuccess) return e;
    e = launch_sort_passes(keysA, valsA, keysB, valsB, nDevice, capacity, beginBit, endBit, ws, numSMs, stream);
    if (launches) *launches += 1 + (endBit - beginBit + kRadixBits - 1) / kRadixBits;
    return e;
}

}  // namespace splat
