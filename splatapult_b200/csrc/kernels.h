// kernels.h -- host-callable launchers of the sm_100a kernels (internal to libsplat_b200.so).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

#include "device_common.cuh"

namespace splat
{

// kernel launch, optionally as a programmatic dependent of the preceding kernel in the stream (device_common.cuh: pdl_wait)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, Args&&... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- pack.cu
size_t packed_cloud_bytes(uint32_t n, int shMode);
cudaError_t detect_sh_mode(const void* deviceAos, uint32_t n, uint32_t strideBytes, bool hasFullSH, unsigned int* dFlagScratch,
                           cudaStream_t stream, int* shModeOut);
cudaError_t pack_cloud(const void* deviceAos, uint32_t n, uint32_t strideBytes, int shMode, void* planes, PackedCloud* out, cudaStream_t stream);

// ---- project.cu
uint32_t project_num_blocks(uint32_t n);
// digit0Hist: 256 global counters that receive the histogram of the emitted keys' lowest digit (may be null)
// (launch_project_keys: 4 x 256 counters, the histograms of all four depth digits)
cudaError_t launch_project(const PackedCloud& cloud, const FrameConstants& fc, ProjRecord* rec, unsigned long long* keys, uint32_t* vals,
                           FrameCounters* ctr, unsigned long long* scanState, uint32_t capacity, uint32_t* digit0Hist, cudaStream_t stream);

// default LSD path: (1) project + one (u32 depthKey, id) pair and an 8-byte tile rectangle per gaussian; after the
// depth digits have been sorted on those N pairs, (2) expand the sorted id list into (u32 tile id, id) instances in
// depth order (scan + coalesced emission) and leave the histogram of the first tile digit in tileDigitHist
cudaError_t launch_project_keys(const PackedCloud& cloud, const FrameConstants& fc, ProjRecord* rec, uint32_t* keys, uint32_t* vals,
                                uint2* side, FrameCounters* ctr, uint32_t* digit0Hist, cudaStream_t stream);
// tileDigitHist: 2 x 256 counters, the histograms of BOTH tile digits (bits 0-7 and 8-15 of the tile id).  pdl: launch as a
// programmatic dependent of the preceding kernel (the last depth pass)
cudaError_t launch_emit_sorted(const uint32_t* sortedVals, const uint2* side, uint32_t n,
                               const FrameConstants& fc, uint32_t* keys, uint32_t* vals, FrameCounters* ctr,
                               unsigned long long* scanState, uint32_t capacity, uint32_t* tileDigitHist, cudaStream_t stream, bool pdl = false);

cudaError_t launch_project_bin(const PackedCloud& cloud, const FrameConstants& fc, ProjRecord* rec, uint4* side, uint32_t* tileCount,
                               FrameCounters* ctr, cudaStream_t stream);
cudaError_t launch_bin_instances(const uint4* side, uint32_t n, int tilesX, uint32_t* tileCursor, unsigned long long* pairs, uint32_t capacity,
                                 cudaStream_t stream);

// ---- binsort.cu : MSD-by-tile frame sort (tile scan + per-segment shared-memory sort)
cudaError_t launch_tile_scan(const uint32_t* tileCount, uint32_t numTiles, uint2* ranges, uint32_t* tileCursor, uint32_t* largeList,
                             FrameCounters* ctr, uint32_t capacity, cudaStream_t stream);
// scratch: a second array of the same size, used only by segments too long for shared memory
cudaError_t launch_segment_sort(unsigned long long* pairs, unsigned long long* scratch, const uint2* ranges, uint32_t numTiles,
                                const uint32_t* largeList, const FrameCounters* ctr, int numSMs, cudaStream_t stream);

// ---- radix_sort.cu : device-wide LSD radix sort of (u64 key, u32 value) pairs, 8-bit digits, onesweep style
struct SortWorkspace
{
    uint32_t* histograms;         // [kMaxSortPasses][256] global digit counts; a sort's first row is filled by its producer,
                                  // row p+1 by pass p
    unsigned long long* status;   // look-back words, sort_status_bytes(capacity, kMaxSortPasses) bytes (never cleared per sort)
    uint32_t* tickets;            // [kMaxSortPasses] dynamic tile ids (zeroed per frame)
    uint64_t capacity;            // keys the status array was sized for
    uint32_t generation;          // distinguishes this frame's look-back words from stale ones (16 bits used)
    int keysPerThread;            // 16 (4096-key tiles, 3 CTAs/SM) or 8 (2048-key tiles, 5 CTAs/SM)
    int passOffset;               // first histogram row / ticket / status region this sort uses (0..kMaxSortPasses-1):
                                  // lets two sorts of one frame (depth digits on N, tile digits on D) share the workspace
    // optional: the LAST pass also writes each tile's range in the sorted output (tile = key >> 32), fusing the
    // boundary-detection kernel into the scatter.  Encoding: ranges[t] = {~begin, end}, zero-initialised by the caller
    // (so an untouched tile reads as begin = 0xFFFFFFFF, end = 0: empty); updated with atomicMax.
    uint2* tileRanges = nullptr;
    uint32_t numRangeTiles = 0;
    // the producer left the histograms of EVERY digit of this sort (k_project_keys does for the four depth digits): no pass
    // accumulates the next digit's histogram
    bool allHistogramsReady = false;
    // optional group aggregates of the look-back (radix_sort.cu): pass p uses words [p * groupWordsPerPass, ...), each pass
    // needs sort_group_words(keys it may see, keysPerThread) 64-bit words, ZEROED by the caller before every sort
    unsigned long long* groupAgg = nullptr;
    size_t groupWordsPerPass = 0;
    // launch every pass as a programmatic dependent of the kernel before it (device_common.cuh: pdl_wait); the first pass
    // then assumes that its histogram row comes from that kernel, later passes only when allHistogramsReady is false
    bool pdl = false;
};
size_t sort_group_words(uint64_t capacity, int keysPerThread);
constexpr uint32_t kFusedRangeBeginXor = 0xFFFFFFFFu;  // begin = ranges[t].x ^ this
size_t sort_status_bytes(uint64_t capacity, int passes);
// keys/vals ping-pong between (A) and (B); passes = ceil((endBit-beginBit)/8); the sorted result is in A when
// `passes` is even, else in B.
// nDevice points at the element count in device memory (no host round trip).
cudaError_t launch_radix_sort(unsigned long long* keysA, uint32_t* valsA, unsigned long long* keysB, uint32_t* valsB,
                              const uint32_t* nDevice, uint64_t capacity, int beginBit, int endBit, const SortWorkspace& ws,
                              int numSMs, cudaStream_t stream, int* launches);

cudaError_t launch_sort_histogram(const unsigned long long* keysA, const uint32_t* nDevice, uint64_t capacity, int beginBit, int endBit,
                                  const SortWorkspace& ws, int numSMs, cudaStream_t stream);
cudaError_t launch_sort_passes(unsigned long long* keysA, uint32_t* valsA, unsigned long long* keysB, uint32_t* valsB, const uint32_t* nDevice,
                               uint64_t capacity, int beginBit, int endBit, const SortWorkspace& ws, int numSMs, cudaStream_t stream);

// same passes on (u32 key, u32 value) pairs, endBit <= 32: 8 bytes per element, 4 CTAs per SM.  The frame's default path
// sorts u32 depth keys on the gaussians and bare u32 tile ids on the instances (the depth order survives the tile passes
// because they are stable, so the instances need not carry the depth key).  With ws.tileRanges the tile id is the key.
cudaError_t launch_sort_passes32(uint32_t* keysA, uint32_t* valsA, uint32_t* keysB, uint32_t* valsB, const uint32_t* nDevice, uint64_t capacity,
                                 int beginBit, int endBit, const SortWorkspace& ws, int numSMs, cudaStream_t stream);

// ---- composite.cu
// keyShift: where the tile id sits in a key (32: tile << 32 | depthKey, u64; 0 with keyBytes 4: bare u32 tile ids)
cudaError_t launch_tile_ranges(const void* sortedKeys, int keyBytes, const uint32_t* nDevice, uint64_t capacity, uint2* ranges,
                               uint32_t numTiles, int numSMs, cudaStream_t stream);
// ids: gaussian id of list entry k is ids[k * idStrideWords] (1: u32 payload array of the LSD sort, 2: low word of the
// binned sort's u64 (depthKey << 32 | id) pairs)
// report: where the frame's instance count is posted for the host (mapped pinned memory; hostWord null = no report)
struct FrameReport
{
    const FrameCounters* counters = nullptr;
    unsigned long long* hostWord = nullptr;  // device pointer of a mapped host word
    unsigned long long seqTag = 0;           // frame sequence number << kReportSeqShift
};
constexpr int kReportSeqShift = 40;
constexpr unsigned long long kReportCountMask = (1ull << kReportSeqShift) - 1;
cudaError_t launch_composite(const uint2* ranges, const uint32_t* ids, int idStrideWords, const ProjRecord* rec, void* framebuffer, int fbFormat,
                             int width, int height, int tilesX, int tilesY, float earlyOutT, int cullMode, uint32_t rangeBeginXor,
                             const FrameReport& report, cudaStream_t stream, bool pdl = false);

// float target -> packed RGBA8 (R in the low byte), same pixel order as the target
cudaError_t launch_present_rgba8(const void* framebuffer, int fbFormat, uint32_t* out, size_t pixels, cudaStream_t stream);

}  // namespace splat
