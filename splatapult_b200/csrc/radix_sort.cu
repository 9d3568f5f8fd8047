// radix_sort.cu -- S: device-wide stable LSD radix sort of (u64 key, u32 payload) pairs, 8-bit digits.
//
// Stands in for the reference's GLSL sorter: 4 x {multi_radixsort_histograms.glsl:29-54 (K2),
// multi_radixsort.glsl:53-138 (K3)} dispatched from src/splatrenderer.cpp:206-264 (and its rgc::radix_sort
// fallback, src/radix_sort.hpp:340-485).  Same contract -- stable, ascending, payload carried along -- on the
// wider (tile << 32 | depthKey) key, so one global sort yields every tile's back-to-front list.
//
// Differences in kind, not only in speed:
//   * the reference re-reduces ALL workgroup histograms in EVERY workgroup (O(WG^2) reads) and only
//     ceil(n/262144) of its workgroups see data.  Here each scatter pass (S2) resolves its tile's offsets with a
//     decoupled look-back over per-tile status words: one read + one write of the data per pass;
//   * digit histograms never cost a pass over the data: pass p accumulates the histogram of digit p+1 while the
//     keys are in registers, and digit 0's comes from the producer (the projection kernel, weighted by each
//     gaussian's tile count; k_sort_histogram for the stand-alone sorter);
//   * key tiles are staged into shared memory by the TMA engine (cp.async.bulk + mbarrier, SASS UBLKCP) instead
//     of per-thread loads; ranks come from per-warp peer masks (atomic OR into shared memory, or MATCH.ANY when the
//     digit is concentrated) + popc ("warp-scan scatter"); the tile is reordered in shared memory so global stores
//     are digit-run coalesced;
//   * look-back words are 32 bits wide and carry a 5-bit generation, so nothing is cleared between sorts (the owner
//     zeroes the array when the generation wraps); element counts are read from device memory: no host round trip;
//   * the last pass of a frame sort also writes every tile's [begin, end) range (SortWorkspace::tileRanges).
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>
#include "device_common.cuh"
#include "kernels.h"

namespace splat
{

constexpr int kHistThreads = 256;

template <typename KeyT>
__device__ __forceinline__ uint32_t digit_of(KeyT key, int shift, uint32_t mask) { return (uint32_t)(key >> shift) & mask; }
__host__ __device__ __forceinline__ uint32_t digit_mask(int shift, int endBit)
{
    const int bits = endBit - shift;
    return bits >= kRadixBits ? (uint32_t)(kRadixSize - 1) : ((1u << bits) - 1u);
}

// below this effective number of digit values (1 / sum of squared bin fractions) a pass ranks with MATCH.ANY
constexpr int kMatchAnyMaxValues = 4;

// Look-back status.  The walk over predecessor tiles is L2-bandwidth bound (a tile reads ~150 predecessor rows of 256
// words against 48 KB of key data), so the word every walker reads is 32 bits:
//   lo[tile][digit] = [31:30] flag | [29] has-hi | [28:24] generation (kSortGenerationMask) | [23:0] count, low 24 bits
//   hi[tile][digit] = count >> 24   -- written (before lo, which is then stored with release) and read (after an acquire
//                                      re-read of lo) only for a PREFIX of 2^24 or more: one digit value with more
//                                      than 16.7 M keys in front of the tile.  Everything else is relaxed 32-bit traffic.
constexpr uint32_t kStAggregate = 1u << 30;
constexpr uint32_t kStPrefix = 2u << 30;
constexpr uint32_t kStHasHi = 1u << 29;
constexpr uint32_t kStCountMask = (1u << 24) - 1u;
__device__ __forceinline__ uint32_t st_pack(uint32_t flag, uint32_t gen, uint32_t count)
{
    return flag | ((gen & kSortGenerationMask) << 24) | (count & kStCountMask);
}
__device__ __forceinline__ bool st_valid(uint32_t w, uint32_t gen) { return (w >> 30) != 0 && ((w >> 24) & kSortGenerationMask) == (gen & kSortGenerationMask); }

// Group aggregates.  With all tiles of a pass resident at once (1 M keys = 245 tiles = one wave) nobody has a prefix when the
// walks start, so tile t used to sum t predecessor rows: O(tiles^2) L2 reads and ~30 dependent round trips per tile.  Every
// kLookGroup consecutive tiles therefore also add their digit counts into ONE 64-bit word per (group, digit):
//   groupAgg[group][digit] += (1 << kGroupArrivalShift) | count          (a single RED.64: arrivals and sum move together,
//                                                                         so no fence or separate "done" counter is needed)
// A walker sums the status words of its own group's predecessors (at most kLookGroup - 1), then one word per earlier group,
// short-cutting as soon as the last tile of a group has published its inclusive prefix.  The words are zeroed per sort by
// the owner of the workspace (they accumulate, so they cannot carry a generation).
constexpr int kLookGroup = 16;
constexpr int kGroupArrivalShift = 40;
constexpr unsigned long long kGroupCountMask = (1ull << kGroupArrivalShift) - 1ull;

// ---------------------------------------------------------------------------------------------------
// S1 (stand-alone sorter only): histogram of the FIRST digit; TMA-staged, double-buffered
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kHistThreads)
k_sort_histogram(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ nDevice, int shift, uint32_t mask,
                 uint32_t* __restrict__ histogram)
{
    extern __shared__ __align__(128) unsigned char smemRaw[];
    unsigned long long* sKeys[2] = {reinterpret_cast<unsigned long long*>(smemRaw),
                                    reinterpret_cast<unsigned long long*>(smemRaw) + kSortTileKeys};
    __shared__ uint32_t sHist[kRadixSize];
    __shared__ __align__(8) uint64_t sBar[2];

    const int tid = threadIdx.x;
    const uint32_t n = *nDevice;
    const uint32_t numTiles = (n + kSortTileKeys - 1) / kSortTileKeys;

    sHist[tid] = 0;
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
        for (int k = 0; k < kSortTileKeys / kHistThreads; k++)
        {
            const uint32_t idx = k * kHistThreads + tid;
            if (idx < cnt) smem_red_add(&sHist[digit_of(sKeys[buf][idx], shift, mask)], 1u);
        }
        __syncthreads();  // all reads of sKeys[buf] done before it is refilled two tiles from now
        buf ^= 1;
    }
    __syncthreads();
    const uint32_t c = sHist[tid];
    if (c) atomicAdd(&histogram[tid], c);
}

// ---------------------------------------------------------------------------------------------------
// S2: one scatter pass (onesweep): rank in tile, look back for the tile's global digit offsets, scatter;
//     also accumulates the NEXT pass's digit histogram
// ---------------------------------------------------------------------------------------------------
// KeyT = unsigned long long (tile << 32 | depthKey, and the stand-alone sorter) or uint32_t (the frame's depth keys on
// the gaussians and bare tile ids on the instances: 8 bytes per element instead of 12, 4 CTAs per SM instead of 3)
template <int KPT, typename KeyT>
struct SortSmem
{
    static constexpr int kTile = kSortThreads * KPT;
    static constexpr int kWarps = kSortThreads / 32;
    KeyT keys[kTile];                             // TMA destination; reused for the digit-ordered tile
    uint32_t vals[kTile];                         // TMA destination; reused for the digit-ordered payloads
    uint32_t warpCount[kWarps][kRadixSize];       // per-warp digit counters -> exclusive prefix over warps
    uint32_t matchMask[kWarps][kRadixSize];       // atomic-OR matching: lanes of the warp holding each digit (zero between rows)
    uint32_t ctaExcl[kRadixSize];                 // first local slot of each digit
    uint32_t globalBase[kRadixSize];              // global index of local slot 0 of each digit (minus ctaExcl)
    uint32_t nextHist[kRadixSize];                // histogram of the next pass's digit over this CTA's tiles
    uint32_t scanTmp[kWarps];
    float sqTmp[kWarps];
    uint32_t tile;
    uint64_t bar;
};

// SPL_SORT_TIMELINE (diagnostic builds only: `SPL_BUILD_DEFINES=SPL_SORT_TIMELINE python -m splatapult_b200.build`):
// thread 0 of every CTA stamps %globaltimer at each phase boundary of each tile; with the environment variable
// SPL_SORT_TIMELINE=<path> set at run time launch_passes_t dumps the stamps; benchmarks/tl_analyze.py reads them.
#ifdef SPL_SORT_TIMELINE
#define TLREC(k)                                                          \
    do                                                                    \
    {                                                                     \
        if (tl && tid == 0)                                               \
        {                                                                 \
            unsigned long long t_;                                        \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));        \
            tl[(size_t)tile * 8 + (k)] = t_;                              \
        }                                                                 \
    } while (0)
#else
#define TLREC(k) ((void)0)
#endif

template <int KPT, typename KeyT>
__global__ void __launch_bounds__(kSortThreads, KPT == 16 ? (sizeof(KeyT) == 4 ? 4 : 3) : 5)
k_sort_onesweep(const KeyT* __restrict__ keysIn, const uint32_t* __restrict__ valsIn, KeyT* __restrict__ keysOut,
                uint32_t* __restrict__ valsOut, const uint32_t* nDevice, int shift, uint32_t mask, int nextShift, uint32_t nextMask,
                const uint32_t* histogram, uint32_t* __restrict__ histogramNext, unsigned long long* __restrict__ status,
                uint32_t* __restrict__ ticket, uint32_t gen, unsigned long long* tl, uint2* __restrict__ tileRanges, uint32_t numRangeTiles,
                uint32_t statusTiles, int rangeShift, unsigned long long* __restrict__ groupAgg, int histFromPredecessor)
{
    using Smem = SortSmem<KPT, KeyT>;
    constexpr int kTile = Smem::kTile;
    constexpr int kWarps = Smem::kWarps;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    Smem& s = *reinterpret_cast<Smem*>(smemRaw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool wantNext = histogramNext != nullptr;
    uint32_t* const statusLo = reinterpret_cast<uint32_t*>(status);
    uint32_t* const statusHi = statusLo + (size_t)statusTiles * kRadixSize;

    // Programmatic dependent launch: this grid may be running while the pass before it drains.  The element count and a
    // histogram row left by the frame's FIRST kernels (all digits known up front) predate that pass and may be read now;
    // a row the preceding kernel itself produces, and the keys, only after pdl_wait().
    if (histFromPredecessor) pdl_wait();
    // (strong loads, not the read-only path: ptxas moves LDG.CONSTANT loads above the wait)
    const uint32_t n = ld_relaxed_u32(nDevice);
    const uint32_t numTiles = (n + kTile - 1) / kTile;

    // global exclusive scan of this pass's digit histogram: thread d owns digit d
    uint32_t digitStart;
    bool concentrated;
    bool digitUsed;  // no key of this pass has digit `tid`: nobody needs its offsets, so it takes no part in the look-back
    {
        const uint32_t h = ld_relaxed_u32(histogram + tid);
        digitUsed = h != 0;
        uint32_t incl = h;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        // sum of squared bin fractions = 1 / (effective number of digit values): decides the matching method below
        float sq = (float)h * (float)h;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, d);
        if (lane == 31) s.scanTmp[warp] = incl;
        if (lane == 0) s.sqTmp[warp] = sq;
        s.nextHist[tid] = 0;
#pragma unroll
        for (int w = 0; w < kWarps; w++) s.matchMask[w][tid] = 0;
        if (tid == 0)
        {
            mbar_init(&s.bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
        uint32_t base = 0;
        float sqAll = 0.0f;
#pragma unroll
        for (int w = 0; w < kWarps; w++)
        {
            if (w < warp) base += s.scanTmp[w];
            sqAll += s.sqTmp[w];
        }
        digitStart = base + incl - h;
        concentrated = sqAll * (float)kMatchAnyMaxValues > (float)n * (float)n;  // effective values < kMatchAnyMaxValues
        __syncthreads();
    }
    if (!histFromPredecessor) pdl_wait();  // the keys are the preceding pass's output
    pdl_trigger();                          // after the wait: see device_common.cuh

    uint32_t parity = 0;
    for (;;)
    {
        if (tid == 0) s.tile = atomicAdd(ticket, 1u);
#pragma unroll
        for (int w = 0; w < kWarps; w++) s.warpCount[w][tid] = 0;
        __syncthreads();
        const uint32_t tile = s.tile;
        if (tile >= numTiles) break;
        const uint32_t nValid = min((uint32_t)kTile, n - tile * kTile);
        TLREC(0);

        if (tid == 0)
        {
            const uint32_t kb = (nValid * (uint32_t)sizeof(KeyT) + 15u) & ~15u, vb = (nValid * 4u + 15u) & ~15u;
            fence_proxy_async();
            mbar_expect_tx(&s.bar, kb + vb);
            bulk_g2s(s.keys, keysIn + (size_t)tile * kTile, kb, &s.bar);
            bulk_g2s(s.vals, valsIn + (size_t)tile * kTile, vb, &s.bar);
        }
        mbar_wait(&s.bar, parity);
        parity ^= 1;
        TLREC(1);

        // ---- rank: warp w owns items [w*KPT*32, (w+1)*KPT*32), item j of a lane = base + j*32 + lane (memory order)
        KeyT key[KPT];
        uint16_t pos[KPT];
#pragma unroll
        for (int j = 0; j < KPT; j++)
        {
            const uint32_t idx = warp * (KPT * 32) + j * 32 + lane;
            key[j] = idx < nValid ? s.keys[idx] : (KeyT)~(KeyT)0;  // padding sorts last inside the tile and is never written
        }
        // peers = lanes of the warp holding the same digit.  Two ways to get it, chosen per pass from the digit
        // histogram (CTA-uniform): MATCH.ANY iterates over the distinct values in the warp (measured per 4096-key
        // tile: 1 us when the digit is constant, 7-10 us when it is spread); an atomic OR of the lane bit into a
        // per-warp, per-digit mask word is conflict-free on spread digits (2.4 us) but serialises when all lanes hit
        // one word (3.6 us).
        if (concentrated)
        {
#pragma unroll
            for (int j = 0; j < KPT; j++)
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
        }
        else
        {
#pragma unroll
            for (int j = 0; j < KPT; j++)
            {
                const uint32_t d = digit_of(key[j], shift, mask);
                atomicOr(&s.matchMask[warp][d], 1u << lane);
                __syncwarp();
                const uint32_t peers = s.matchMask[warp][d];
                const uint32_t before = __popc(peers & lanemask_lt());
                const uint32_t prev = s.warpCount[warp][d];
                __syncwarp();
                if (before == 0)
                {
                    s.warpCount[warp][d] = prev + __popc(peers);
                    s.matchMask[warp][d] = 0u;  // ready for the next row
                }
                __syncwarp();
                pos[j] = (uint16_t)(prev + before);
            }
        }
        TLREC(6);
        if (wantNext)
        {
            // next pass's digit histogram, while the keys are in registers (shared-memory REDs: 0.6-0.8 us per tile;
            // aggregating them per warp first costs more than it saves)
#pragma unroll
            for (int j = 0; j < KPT; j++)
            {
                const uint32_t idx = warp * (KPT * 32) + j * 32 + lane;
                if (idx < nValid) smem_red_add(&s.nextHist[digit_of(key[j], nextShift, nextMask)], 1u);
            }
        }
        __syncthreads();
        TLREC(2);

        // ---- per-digit totals (thread d), publish, look back, local/global bases
        {
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < kWarps; w++)
            {
                const uint32_t c = s.warpCount[w][tid];
                s.warpCount[w][tid] = run;
                run += c;
            }
            const uint32_t cnt = run;
            uint32_t* myLo = statusLo + (size_t)tile * kRadixSize + tid;
            uint32_t* myHi = statusHi + (size_t)tile * kRadixSize + tid;
            const uint32_t grp = tile / kLookGroup;
            if (digitUsed)
            {
                if (tile == 0)
                {
                    st_relaxed_u32(myLo, st_pack(kStPrefix, gen, cnt));  // cnt <= tile size: no high part
                }
                else
                {
                    st_relaxed_u32(myLo, st_pack(kStAggregate, gen, cnt));
                }
                // arrival + count into the group's word (result unused: compiles to RED.64); also when cnt == 0
                if (groupAgg) atomicAdd(groupAgg + (size_t)grp * kRadixSize + tid, (1ull << kGroupArrivalShift) | (unsigned long long)cnt);
            }

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
            for (int w = 0; w < kWarps; w++)
                if (w < warp) base += s.scanTmp[w];
            const uint32_t excl = base + incl - cnt;

            uint32_t prior = 0;
            if (tile > 0 && digitUsed)
            {
                // windowed look-back: kLookWindow predecessors are polled with independent loads in flight, so
                // the walk costs one L2 round trip per window instead of one per tile
                constexpr int kLookWindow = 8;
                // phase 1: the predecessors inside this tile's own group (all of them when there are no group aggregates)
                const int floorTile = groupAgg ? (int)(grp * kLookGroup) : 0;
                int t = (int)tile - 1;
                bool done = false, hasHi = false;
                while (!done)
                {
                    uint32_t v[kLookWindow];
#pragma unroll
                    for (int i = 0; i < kLookWindow; i++)
                        v[i] = (t - i >= floorTile) ? ld_relaxed_u32(statusLo + (size_t)(t - i) * kRadixSize + tid) : st_pack(kStPrefix, gen, 0u);
                    int used = 0;
#pragma unroll
                    for (int i = 0; i < kLookWindow; i++)
                    {
                        if (!done && used == i)
                        {
                            if (st_valid(v[i], gen))
                            {
                                prior += v[i] & kStCountMask;
                                used = i + 1;
                                done = (v[i] >> 30) == 2;
                                hasHi = done && (v[i] & kStHasHi) != 0u;
                            }
                        }
                    }
                    t -= used;  // re-poll from the first word that was not yet published
                }
                int prefixTile = t + 1;  // the walk ended on the prefix word of this tile (floorTile - 1: the virtual one)
                if (prefixTile < floorTile && floorTile > 0)
                {
                    // phase 2: one word per earlier group (all of them complete groups of kLookGroup tiles), newest first;
                    // the inclusive prefix of a group's LAST tile, once published, ends the walk
                    constexpr int kGroupWindow = 4;
                    int gg = (int)grp - 1;
                    done = false;
                    hasHi = false;
                    while (!done)
                    {
                        unsigned long long a[kGroupWindow];
                        uint32_t p[kGroupWindow];
#pragma unroll
                        for (int i = 0; i < kGroupWindow; i++)
                        {
                            const int gi = gg - i;
                            a[i] = gi >= 0 ? ld_relaxed_u64(groupAgg + (size_t)gi * kRadixSize + tid) : 0ull;
                            p[i] = gi >= 0 ? ld_relaxed_u32(statusLo + ((size_t)gi * kLookGroup + (kLookGroup - 1)) * kRadixSize + tid) : 0u;
                        }
                        int used = 0;
#pragma unroll
                        for (int i = 0; i < kGroupWindow; i++)
                        {
                            if (!done && used == i)
                            {
                                const int gi = gg - i;
                                if (gi < 0)
                                {
                                    done = true;  // before group 0: prefix 0
                                }
                                else if (st_valid(p[i], gen) && (p[i] >> 30) == 2)
                                {
                                    prior += p[i] & kStCountMask;
                                    hasHi = (p[i] & kStHasHi) != 0u;
                                    prefixTile = gi * kLookGroup + (kLookGroup - 1);
                                    done = true;
                                }
                                else if ((a[i] >> kGroupArrivalShift) == (unsigned long long)kLookGroup)
                                {
                                    prior += (uint32_t)(a[i] & kGroupCountMask);
                                    used = i + 1;
                                }
                            }
                        }
                        gg -= used;  // re-poll from the first group that was not yet complete
                    }
                }
                // if the prefix the walk ended on did not fit 24 bits its publisher stored hi, then lo with release: re-read lo
                // with acquire, then hi
                if (hasHi)
                {
                    (void)ld_acquire_u32(statusLo + (size_t)prefixTile * kRadixSize + tid);
                    prior += ld_relaxed_u32(statusHi + (size_t)prefixTile * kRadixSize + tid) << 24;
                }
                const uint32_t total = prior + cnt;
                if (total > kStCountMask)
                {
                    st_relaxed_u32(myHi, total >> 24);
                    st_release_u32(myLo, st_pack(kStPrefix | kStHasHi, gen, total));
                }
                else
                {
                    st_relaxed_u32(myLo, st_pack(kStPrefix, gen, total));
                }
            }
            s.ctaExcl[tid] = excl;
            s.globalBase[tid] = digitStart + prior - excl;
        }
        __syncthreads();
        TLREC(3);

        // ---- reorder the tile in shared memory
#pragma unroll
        for (int j = 0; j < KPT; j++)
        {
            const uint32_t d = digit_of(key[j], shift, mask);
            pos[j] = (uint16_t)(s.ctaExcl[d] + s.warpCount[warp][d] + pos[j]);
        }
        uint32_t val[KPT];
#pragma unroll
        for (int j = 0; j < KPT; j++) val[j] = s.vals[warp * (KPT * 32) + j * 32 + lane];
        __syncthreads();  // every staged key/payload is in registers
#pragma unroll
        for (int j = 0; j < KPT; j++)
        {
            s.keys[pos[j]] = key[j];
            s.vals[pos[j]] = val[j];
        }
        __syncthreads();
        TLREC(4);

        // ---- coalesced scatter: slot i of digit d goes to globalBase[d] + i
#pragma unroll 4
        for (int k = 0; k < KPT; k++)
        {
            const uint32_t i = k * kSortThreads + tid;
            if (i < nValid)
            {
                const KeyT kk = s.keys[i];
                const uint32_t g = s.globalBase[digit_of(kk, shift, mask)] + i;
                keysOut[g] = kk;
                valsOut[g] = s.vals[i];
                if (tileRanges)
                {
                    // last pass: neighbouring slots with the same high word (tile id) are neighbours in the output too,
                    // so only the first / last slot of each tile-id group in this CTA's tile can be a range end point
                    const uint32_t t = (uint32_t)(kk >> rangeShift);  // tile id: high word of a u64 key, a u32 key itself
                    const uint32_t tp = i > 0 ? (uint32_t)(s.keys[i - 1] >> rangeShift) : ~t;
                    const uint32_t tn = i + 1 < nValid ? (uint32_t)(s.keys[i + 1] >> rangeShift) : ~t;
                    if (t < numRangeTiles)
                    {
                        if (tp != t) atomicMax(&tileRanges[t].x, ~g);
                        if (tn != t) atomicMax(&tileRanges[t].y, g + 1u);
                    }
                }
            }
        }
        __syncthreads();  // shared tile fully consumed before the next bulk copy lands
        TLREC(5);
    }
    if (wantNext)
    {
        const uint32_t c = s.nextHist[tid];
        if (c) atomicAdd(&histogramNext[tid], c);
    }
}

size_t sort_group_words(uint64_t capacity, int keysPerThread)
{
    const uint64_t tileKeys = (uint64_t)kSortThreads * (keysPerThread == 8 ? 8 : 16);
    const uint64_t tiles = (capacity + tileKeys - 1) / tileKeys;
    return (size_t)((tiles + kLookGroup - 1) / kLookGroup + 1) * kRadixSize;
}

size_t sort_status_bytes(uint64_t capacity, int passes)
{
    const uint64_t maxTiles = (capacity + kSortMinTileKeys - 1) / kSortMinTileKeys + 1;
    return (size_t)passes * maxTiles * kRadixSize * sizeof(unsigned long long);
}

template <int KPT, typename KeyT>
static cudaError_t launch_passes_t(KeyT* keysA, uint32_t* valsA, KeyT* keysB, uint32_t* valsB, const uint32_t* nDevice,
                                   uint64_t capacity, int beginBit, int endBit, const SortWorkspace& ws, int numSMs, cudaStream_t stream)
{
    using Smem = SortSmem<KPT, KeyT>;
    constexpr int kRangeShift = sizeof(KeyT) == 8 ? 32 : 0;  // where the tile id sits in a key
    {
        // cudaFuncSetAttribute is per device; one table per kernel instantiation, filled under a lock
        static std::mutex attrMutex;
        static bool attrSet[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lock(attrMutex);
        if (!attrSet[dev & 63])
        {
            cudaError_t e = cudaFuncSetAttribute(k_sort_onesweep<KPT, KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
            if (e != cudaSuccess) return e;
            attrSet[dev & 63] = true;
        }
    }
    const int passes = (endBit - beginBit + kRadixBits - 1) / kRadixBits;
    const uint64_t maxTiles = (capacity + kSortMinTileKeys - 1) / kSortMinTileKeys + 1;
    const uint32_t capTiles = (uint32_t)((capacity + Smem::kTile - 1) / Smem::kTile);
    const uint32_t ctasPerSM = KPT == 16 ? (sizeof(KeyT) == 4 ? 4u : 3u) : 5u;
    const uint32_t grid = max(1u, min((uint32_t)numSMs * ctasPerSM, capTiles));
    unsigned long long* dTl = nullptr;
    size_t tlWords = 0;
#ifdef SPL_SORT_TIMELINE
    static unsigned long long* dTlStatic = nullptr;
    const char* tlPath = getenv("SPL_SORT_TIMELINE");
    tlWords = (size_t)maxTiles * 8;
    if (tlPath && !dTlStatic) cudaMalloc(&dTlStatic, kMaxSortPasses * tlWords * 8);
    if (tlPath) cudaMemsetAsync(dTlStatic, 0, kMaxSortPasses * tlWords * 8, stream);
    dTl = tlPath ? dTlStatic : nullptr;
#endif
    for (int p = 0; p < passes; p++)
    {
        const KeyT* kin = (p & 1) ? keysB : keysA;
        const uint32_t* vin = (p & 1) ? valsB : valsA;
        KeyT* kout = (p & 1) ? keysA : keysB;
        uint32_t* vout = (p & 1) ? valsA : valsB;
        const int shift = beginBit + p * kRadixBits;
        const bool last = p == passes - 1;
        const bool wantNext = !last && !ws.allHistogramsReady;
        const int nshift = last ? 0 : shift + kRadixBits;
        // a histogram row written by the kernel right before this pass: the producer's for pass 0, the previous pass's when the
        // passes accumulate each other's histograms
        const int histFromPredecessor = (p == 0 || !ws.allHistogramsReady) ? 1 : 0;
        cudaError_t le = launch_kernel(
            k_sort_onesweep<KPT, KeyT>, dim3(grid), dim3(kSortThreads), sizeof(Smem), stream, ws.pdl,
            kin, vin, kout, vout, nDevice, shift, digit_mask(shift, endBit), nshift, last ? 0u : digit_mask(nshift, endBit),
            (const uint32_t*)(ws.histograms + (ws.passOffset + p) * kRadixSize), wantNext ? ws.histograms + (ws.passOffset + p + 1) * kRadixSize : (uint32_t*)nullptr,
            ws.status + (size_t)(ws.passOffset + p) * maxTiles * kRadixSize, ws.tickets + ws.passOffset + p, ws.generation, dTl ? dTl + p * tlWords : (unsigned long long*)nullptr,
            last && shift >= kRangeShift ? ws.tileRanges : (uint2*)nullptr, ws.numRangeTiles,  // the digit must be part of the tile id
            (uint32_t)maxTiles, kRangeShift, ws.groupAgg ? ws.groupAgg + (size_t)p * ws.groupWordsPerPass : (unsigned long long*)nullptr, histFromPredecessor);
        if (le != cudaSuccess) return le;
    }
#ifdef SPL_SORT_TIMELINE
    if (tlPath)
    {
        cudaStreamSynchronize(stream);
        uint32_t nHost = 0;
        cudaMemcpy(&nHost, nDevice, 4, cudaMemcpyDeviceToHost);
        const uint32_t tiles = (nHost + Smem::kTile - 1) / Smem::kTile;
        std::vector<unsigned long long> h((size_t)passes * tiles * 8);
        for (int p = 0; p < passes; p++) cudaMemcpy(h.data() + (size_t)p * tiles * 8, dTl + p * tlWords, (size_t)tiles * 64, cudaMemcpyDeviceToHost);
        char name[512];
        snprintf(name, sizeof name, "%s.b%d", tlPath, beginBit);
        if (FILE* f = fopen(name, "wb")) { uint32_t hdr[4] = {(uint32_t)passes, tiles, nHost, (uint32_t)grid}; fwrite(hdr, 4, 4, f); fwrite(h.data(), 8, h.size(), f); fclose(f); }
    }
#endif
    return cudaGetLastError();
}

cudaError_t launch_sort_histogram(const unsigned long long* keysA, const uint32_t* nDevice, uint64_t capacity, int beginBit, int endBit,
                                  const SortWorkspace& ws, int numSMs, cudaStream_t stream)
{
    const size_t smem = 2 * kSortTileKeys * sizeof(unsigned long long);
    {
        static std::mutex attrMutex;
        static bool attrSet[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lock(attrMutex);
        if (!attrSet[dev & 63])
        {
            cudaError_t e = cudaFuncSetAttribute(k_sort_histogram, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            attrSet[dev & 63] = true;
        }
    }
    const uint32_t capTiles = (uint32_t)((capacity + kSortTileKeys - 1) / kSortTileKeys);
    const uint32_t grid = max(1u, min((uint32_t)numSMs * 3u, capTiles));
    k_sort_histogram<<<grid, kHistThreads, smem, stream>>>(keysA, nDevice, beginBit, digit_mask(beginBit, endBit),
                                                           ws.histograms + ws.passOffset * kRadixSize);
    return cudaGetLastError();
}

cudaError_t launch_sort_passes(unsigned long long* keysA, uint32_t* valsA, unsigned long long* keysB, uint32_t* valsB, const uint32_t* nDevice,
                               uint64_t capacity, int beginBit, int endBit, const SortWorkspace& ws, int numSMs, cudaStream_t stream)
{
    const int passes = (endBit - beginBit + kRadixBits - 1) / kRadixBits;
    if (passes < 1 || passes > kMaxSortPasses) return cudaErrorInvalidValue;
    if (ws.keysPerThread == 8) return launch_passes_t<8, unsigned long long>(keysA, valsA, keysB, valsB, nDevice, capacity, beginBit, endBit, ws, numSMs, stream);
    return launch_passes_t<16, unsigned long long>(keysA, valsA, keysB, valsB, nDevice, capacity, beginBit, endBit, ws, numSMs, stream);
}

cudaError_t launch_sort_passes32(uint32_t* keysA, uint32_t* valsA, uint32_t* keysB, uint32_t* valsB, const uint32_t* nDevice, uint64_t capacity,
                                 int beginBit, int endBit, const SortWorkspace& ws, int numSMs, cudaStream_t stream)
{
    const int passes = (endBit - beginBit + kRadixBits - 1) / kRadixBits;
    if (passes < 1 || passes > 4 || endBit > 32) return cudaErrorInvalidValue;
    if (ws.keysPerThread == 8) return launch_passes_t<8, uint32_t>(keysA, valsA, keysB, valsB, nDevice, capacity, beginBit, endBit, ws, numSMs, stream);
    return launch_passes_t<16, uint32_t>(keysA, valsA, keysB, valsB, nDevice, capacity, beginBit, endBit, ws, numSMs, stream);
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
