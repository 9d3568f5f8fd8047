// renderer.cu -- SplatRenderer: owns the device buffers and enqueues the frame (default: project_keys -> 4 depth-digit
// passes on the gaussians -> emit -> tile-digit passes on the instances, the last one writing the tile ranges -> composite).
//
// Call protocol and argument meaning follow the reference (src/splatrenderer.h:23-67, src/splatrenderer.cpp:50-343):
//   Init   : src/splatrenderer.cpp:50-151   (upload the cloud once; no reference to host data afterwards)
//   Sort   : src/splatrenderer.cpp:153-312  (pre-sort + radix sort; here also projection/SH and tile ranges)
//   Render : src/splatrenderer.cpp:315-343  (the draw; here the per-tile compositor into an explicit target)
// Unlike the reference there is no host synchronisation on the frame path and no GL state: the render target,
// its clear colour (0,0,0,1) and the "over" blend (src/app.cpp:144-164) are part of Render itself.
#include "splatapult_b200/splatrenderer.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: a no-op unless a profiler injects itself

#include "kernels.h"
#include "plydecode.h"
#include "splat_b200.h"
#include "splatapult_b200/ply.h"
#include "splatapult_b200/splatcomm.h"

namespace splat
{

#define CU_TRY(expr)                                                                                 \
    do                                                                                               \
    {                                                                                                \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess) return fail(SPL_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

// host-side range, closed on every exit path (the CU_TRY macro returns early)
struct NvtxScope
{
    explicit NvtxScope(const char* name) { nvtxRangePushA(name); }
    ~NvtxScope() { nvtxRangePop(); }
    NvtxScope(const NvtxScope&) = delete;
    NvtxScope& operator=(const NvtxScope&) = delete;
};

struct RendererImpl
{
    // configuration
    int device = 0;
    int fbFormat = SPL_FB_FP32;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    int numSMs = 148;
    double earlyOutT = 1e-4;
    int subblockCull = 2;
    bool profile = false;
    double instancesPerGaussian = 16.0;
    int sortMode = SPL_SORT_LSD;
    bool graphOption = true;           // SPL_OPT_GRAPH: replay a Sort's device work as one CUDA graph
    cudaGraphExec_t sortExec = nullptr;
    uint64_t warmFrames = 0;           // frames enqueued directly since the last configuration change
    uint64_t sortExecShape = 0;        // graphShape() the executable graph was instantiated for
    // SPL_OPT_PDL: which kernels of a frame are launched as programmatic dependents of the kernel before them
    // (bit 0: depth passes, bit 1: emission, bit 2: tile passes, bit 3: compositor); stage events between the kernels
    // (profile mode) would break the chain
    // Default 5, measured with the graph on (1 M deg 0 / 6 M deg 3 frames, ms): none 0.350 / 1.008; depth passes 0.338 / 1.000;
    // emission 0.352 / 1.073 (its CTAs take their scan tickets before the wait and then sit on SM slots); tile passes
    // 0.347 / 1.004; compositor 0.350 / 1.008; depth + tile passes 0.334 / 0.997.
    int pdlMask = 5;
    bool usePdl(int bit) const { return ((pdlMask >> bit) & 1) != 0 && !profile; }

    // cloud
    bool initialized = false;
    bool srgb = false;
    uint32_t n = 0;
    void* dPlanes = nullptr;
    PackedCloud cloud{};
    ProjRecord* dRec = nullptr;

    // per-frame buffers
    uint64_t capacity = 0;
    unsigned long long *dKeysA = nullptr, *dKeysB = nullptr;
    uint32_t *dValsA = nullptr, *dValsB = nullptr;
    void* dControl = nullptr;
    size_t controlBytes = 0;           // fixed part: counters | histograms | scan state | tile ranges (cleared every frame)
    size_t controlAllocBytes = 0;      // + the sorter's group-aggregate words, of which a frame clears only what its passes use
    unsigned long long* dGroups = nullptr;  // inside dControl, right after the fixed part
    FrameCounters* dCtr = nullptr;
    uint32_t* dHist = nullptr;
    unsigned long long* dScan = nullptr;
    unsigned long long* dStatus = nullptr;
    uint32_t sortGeneration = 0;
    const unsigned long long* dSortedKeys = nullptr;  // where the last LSD sort left its result ...
    int sortedKeyBytes = 8;                           // ... as u64 (tile << 32 | depthKey) or, default path, u32 tile ids
    const uint32_t* dSortedVals = nullptr;
    int sortKptOption = 0;  // SPL_OPT_SORT_TILE: 0 = automatic (below), 8 = 2048-key tiles, 16 = 4096-key tiles
    // keys per thread of a sort over at most `keys` elements.  The passes are latency bound per tile; when 4096-key tiles do
    // not fill one wave of CTAs (1 M gaussians = 245 tiles on 592 slots) 2048-key tiles finish a pass sooner (measured:
    // depth passes 104.6 -> 97.4 us at 1 M), with more tiles than slots the larger tile wins (6 M: 312 vs 337 us)
    int kptFor(uint64_t keys) const
    {
        if (sortKptOption) return sortKptOption;
        return (keys + kSortTileKeys - 1) / kSortTileKeys < (uint64_t)numSMs * 4 ? 8 : 16;
    }
    int depthKpt() const { return kptFor(n); }
    // what decides WHICH nodes a captured Sort consists of (kernel variants, pass counts): cudaGraphExecUpdate only takes
    // parameter changes, so an executable graph of another shape is dropped, not updated
    uint64_t graphShape() const
    {
        uint64_t h = 1469598103934665603ull;
        auto mix = [&](uint64_t v) { h = (h ^ v) * 1099511628211ull; };
        mix(sortMode); mix(fc.stereoPass); mix(sortPasses); mix(sortBits); mix(pdlMask); mix(sortKptOption); mix(depthKpt());
        mix(n); mix(cloud.shMode); mix(rangesTiles);
        return h;
    }
    int instKpt() const { return sortKptOption ? sortKptOption : 16; }  // the instance count is only known on the device
    uint2* dRanges = nullptr;
    uint32_t* dTileCtl = nullptr;  // tileCount | tileCursor | largeList, rangesTiles words each
    uint4* dSide = nullptr;        // per-gaussian {rect, depthKey, count} of the binned sort
    uint32_t rangesTiles = 0;
    void* dFb = nullptr;             // target of the last Render (= dFbs[fbCur])
    void* dFbs[2] = {nullptr, nullptr};
    void* dPresent[2] = {nullptr, nullptr};  // RGBA8 images of the two targets (allocated on first ReadPresented8)
    size_t presentBytes = 0;
    int fbNext = 0, fbCur = 0;
    cudaStream_t copyStream = nullptr;  // framebuffer D->H copies overlap the next frame
    cudaEvent_t evRendered[2] = {}, evCopied[2] = {};
    bool copyPending[2] = {false, false};
    size_t fbBytes = 0;
    int fbW = 0, fbH = 0;

    // last frame
    bool haveSort = false, haveRender = false;
    FrameConstants fc{};
    float lastArgs[16 + 16 + 4 + 2] = {};   // arguments of the last Sort
    float drawArgs[16 + 16 + 4 + 2] = {};   // arguments the current tile lists were built for (= lastArgs unless a
                                            // shared-sort Render re-projected them for another eye)
    bool sharedSort = false;                // SPL_OPT_SHARED_SORT: Render for other arguments keeps the last Sort's order
    uint32_t* dOrder = nullptr;             // shared sort: the depth-sorted gaussian ids of the last Sort (n words)
    int sortPasses = 0, sortBits = 0;
    int frameSortMode = SPL_SORT_LSD;  // sort path the last Sort() used
    bool rangesFused = false;          // dRanges holds {~begin, end} written by the last sort pass (kernels.h)
    uint32_t launches = 0;
    cudaEvent_t ev[7] = {};
    bool evCreated = false;

    // frame reports: the compositor posts {frames composited, instances the frame needed} into mapped host memory, so a
    // capacity overflow on the ASYNCHRONOUS paths (Render + GetFramebuffer, ReadFramebufferAsync) is noticed without a
    // synchronising read-back: see pollReports()
    static constexpr int kReportRing = 64;
    unsigned long long* hReport = nullptr;  // [kReportRing] ring of (seq << kReportSeqShift | instances needed)
    unsigned long long* dReport = nullptr;  // device address of hReport
    uint64_t frameSeq = 0;                  // frames composited so far (the newest frame's sequence number)
    uint64_t checkedSeq = 0;                // newest frame whose report has been looked at
    uint64_t seqCapacity[kReportRing] = {}; // capacity frame (seq % kReportRing) ran with
    uint64_t growTo = 0;                    // capacity the next Sort must reserve (an earlier frame overflowed)

    int status = 0;
    std::string error;

    int fail(int code, const std::string& msg)
    {
        status = code;
        error = msg;
        return code;
    }
    int ok()
    {
        status = 0;
        return 0;
    }

    ~RendererImpl() { release(); }

    void freeFrameBuffers()
    {
        cudaFree(dKeysA); cudaFree(dKeysB); cudaFree(dValsA); cudaFree(dValsB); cudaFree(dStatus);
        dKeysA = dKeysB = nullptr; dValsA = dValsB = nullptr; dStatus = nullptr;
        capacity = 0;
    }
    void release()
    {
        if (!initialized && !stream && !dFb && !hReport) return;
        cudaSetDevice(device);
        if (stream) cudaStreamSynchronize(stream);
        if (copyStream)
        {
            cudaStreamSynchronize(copyStream);
            for (int k = 0; k < 2; k++) { cudaEventDestroy(evRendered[k]); cudaEventDestroy(evCopied[k]); }
            cudaStreamDestroy(copyStream);
            copyStream = nullptr;
        }
        freeFrameBuffers();
        cudaFree(dPlanes); cudaFree(dRec); cudaFree(dControl); cudaFree(dFbs[0]); cudaFree(dFbs[1]); cudaFree(dTileCtl); cudaFree(dSide);  // dRanges lives in dControl
        cudaFree(dPresent[0]); cudaFree(dPresent[1]);
        dPresent[0] = dPresent[1] = nullptr;
        presentBytes = 0;
        cudaFree(dOrder);
        dOrder = nullptr;
        dPlanes = nullptr; dRec = nullptr; dControl = nullptr; dRanges = nullptr; dFb = dFbs[0] = dFbs[1] = nullptr; dTileCtl = nullptr; dSide = nullptr;
        if (evCreated)
            for (auto& e : ev) cudaEventDestroy(e);
        evCreated = false;
        if (ownStream && stream) cudaStreamDestroy(stream);
        stream = nullptr;
        ownStream = false;
        initialized = false;
        if (sortExec) cudaGraphExecDestroy(sortExec);
        sortExec = nullptr;
        if (hReport) cudaFreeHost(hReport);
        hReport = dReport = nullptr;
    }

    int begin()
    {
        cudaError_t e = cudaSetDevice(device);
        if (e != cudaSuccess) return fail(SPL_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e) + " (no CPU fallback exists)");
        if (!stream)
        {
            CU_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
            ownStream = true;
        }
        if (!hReport)
        {
            CU_TRY(cudaHostAlloc(reinterpret_cast<void**>(&hReport), kReportRing * sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable));
            std::memset(hReport, 0, kReportRing * sizeof(unsigned long long));
            CU_TRY(cudaHostGetDevicePointer(reinterpret_cast<void**>(&dReport), hReport, 0));
        }
        return 0;
    }

    // Look at the frame reports that have landed since the last call.  Returns the number of frames found INCOMPLETE (they
    // needed more tile instances than the capacity they ran with; the nearest splats are the ones missing) and writes a
    // message naming them; schedules the capacity growth for the next Sort.  Costs a few host loads, no CUDA call.
    int pollReports(std::string* msg)
    {
        if (!hReport || frameSeq == checkedSeq) return 0;
        int incomplete = 0;
        uint64_t need = 0, firstBad = 0, cap = 0;
        // reports land in frame order; walk from the oldest unchecked frame that can still be in the ring
        uint64_t s0 = checkedSeq + 1;
        if (frameSeq - checkedSeq > (uint64_t)kReportRing) s0 = frameSeq - kReportRing + 1;
        for (uint64_t q = s0; q <= frameSeq; q++)
        {
            const unsigned long long v = *reinterpret_cast<volatile unsigned long long*>(hReport + q % kReportRing);
            if ((v >> kReportSeqShift) != (q & ((1ull << (64 - kReportSeqShift)) - 1))) break;  // not landed yet
            checkedSeq = q;
            const uint64_t total = v & kReportCountMask;
            if (total > seqCapacity[q % kReportRing])
            {
                if (!incomplete) { firstBad = q; cap = seqCapacity[q % kReportRing]; }
                incomplete++;
                need = std::max<uint64_t>(need, total);
            }
        }
        if (incomplete)
        {
            growTo = std::max<uint64_t>(growTo, need + need / 4);
            if (msg)
                *msg = std::to_string(incomplete) + " frame(s) (first: frame " + std::to_string(firstBad) + ") needed up to " + std::to_string(need) +
                       " tile instances but ran with a capacity of " + std::to_string(cap) +
                       ": those images are incomplete (nearest splats missing); the capacity is raised before the next Sort";
        }
        return incomplete;
    }
    // a frame the caller is about to replace (blocking read-back re-renders it): its report must not be raised again
    void forgetReports() { checkedSeq = frameSeq; }

    int reserve(uint64_t want)
    {
        want = std::min<uint64_t>(std::max<uint64_t>(want, kSortTileKeys), kMaxInstances);
        if (want <= capacity) return ok();
        CU_TRY(cudaStreamSynchronize(stream));
        freeFrameBuffers();
        const uint64_t padded = want + kSortTileKeys;  // bulk copies may read up to a tile past the last key
        CU_TRY(cudaMalloc(&dKeysA, padded * 8));
        CU_TRY(cudaMalloc(&dKeysB, padded * 8));
        CU_TRY(cudaMalloc(&dValsA, padded * 4));
        CU_TRY(cudaMalloc(&dValsB, padded * 4));
        CU_TRY(cudaMalloc(&dStatus, sort_status_bytes(want, kMaxSortPasses)));
        CU_TRY(cudaMemsetAsync(dStatus, 0, sort_status_bytes(want, kMaxSortPasses), stream));
        sortGeneration = 0;
        // the padding is read (never used) by the last tile's bulk copy: keep it defined
        CU_TRY(cudaMemsetAsync(dKeysA, 0, padded * 8, stream));
        CU_TRY(cudaMemsetAsync(dKeysB, 0, padded * 8, stream));
        CU_TRY(cudaMemsetAsync(dValsA, 0, padded * 4, stream));
        CU_TRY(cudaMemsetAsync(dValsB, 0, padded * 4, stream));
        capacity = want;
        if (dControl) return allocControl(rangesTiles);  // the group-aggregate words are sized by the capacity
        return ok();
    }

    // control block: counters | digit histograms | projection scan state | tile ranges -- everything one memset per
    // frame has to clear.  (Re)allocated when the cloud or the tile count changes; the caller has synchronised.
    int allocControl(uint32_t tiles)
    {
        cudaFree(dControl);
        dControl = nullptr;
        const size_t scanBytes = (size_t)std::max<uint32_t>(project_num_blocks(n), 1) * sizeof(unsigned long long);
        const size_t histBytes = kMaxSortPasses * kRadixSize * sizeof(uint32_t);
        controlBytes = (256 + histBytes + scanBytes + (size_t)tiles * sizeof(uint2) + 15) & ~(size_t)15;
        // group aggregates: four depth passes over at most n keys, then the instance passes over at most `capacity` keys
        // (sized for the small sort tile, 8 keys per thread: SPL_OPT_SORT_TILE may change after Init)
        const size_t groupWords = 4 * sort_group_words(std::max<uint32_t>(n, 1), 8) + (size_t)kMaxSortPasses * sort_group_words(std::max<uint64_t>(capacity, 1), 8);
        controlAllocBytes = controlBytes + groupWords * sizeof(unsigned long long);
        CU_TRY(cudaMalloc(&dControl, controlAllocBytes));
        uint8_t* base = static_cast<uint8_t*>(dControl);
        dCtr = reinterpret_cast<FrameCounters*>(base);
        dHist = reinterpret_cast<uint32_t*>(base + 256);
        dScan = reinterpret_cast<unsigned long long*>(base + 256 + histBytes);
        dRanges = reinterpret_cast<uint2*>(base + 256 + histBytes + scanBytes);
        dGroups = reinterpret_cast<unsigned long long*>(base + controlBytes);
        CU_TRY(cudaMemsetAsync(dControl, 0, controlAllocBytes, stream));
        return 0;
    }
    // bytes of the control block a frame with `depthPasses` passes over the gaussians and `instPasses` over the instances clears
    size_t frameControlBytes(int depthPasses, int instPasses) const
    {
        return controlBytes + ((size_t)depthPasses * sort_group_words(std::max<uint32_t>(n, 1), depthKpt()) +
                               (size_t)instPasses * sort_group_words(std::max<uint64_t>(capacity, 1), instKpt())) * sizeof(unsigned long long);
    }

    int initFromDevice(const void* dAos, size_t count, size_t stride, bool fullSH, bool srgbIn)
    {
        if (int rc = begin()) return rc;
        if (stride != 100 && stride != 244) return fail(SPL_ERR_ARG, "stride must be 100 (base) or 244 (full SH)");
        if (fullSH && stride != 244) return fail(SPL_ERR_ARG, "full SH needs stride 244");
        if (count >= (1ull << 31)) return fail(SPL_ERR_ARG, "too many gaussians");
        cudaDeviceProp prop{};
        CU_TRY(cudaGetDeviceProperties(&prop, device));
        numSMs = prop.multiProcessorCount;
        if (prop.major < 10) return fail(SPL_ERR_CUDA, "libsplat_b200 is built for sm_100a only (found sm_" + std::to_string(prop.major * 10 + prop.minor) + ")");

        // drop any previous cloud
        CU_TRY(cudaStreamSynchronize(stream));
        cudaFree(dPlanes); cudaFree(dRec); cudaFree(dControl); cudaFree(dSide); cudaFree(dOrder);
        dOrder = nullptr;
        dPlanes = nullptr; dRec = nullptr; dControl = nullptr; dRanges = nullptr; dSide = nullptr;

        n = (uint32_t)count;
        srgb = srgbIn;
        if (int rc = allocControl(rangesTiles)) return rc;

        int shMode = 0;
        CU_TRY(detect_sh_mode(dAos, n, (uint32_t)stride, fullSH, reinterpret_cast<unsigned int*>(dControl), stream, &shMode));
        CU_TRY(cudaMemsetAsync(dControl, 0, 256, stream));
        CU_TRY(cudaMalloc(&dPlanes, packed_cloud_bytes(n, shMode)));
        CU_TRY(pack_cloud(dAos, n, (uint32_t)stride, shMode, dPlanes, &cloud, stream));
        CU_TRY(cudaMalloc(&dRec, std::max<size_t>(n, 1) * sizeof(ProjRecord)));
        CU_TRY(cudaMemsetAsync(dRec, 0, std::max<size_t>(n, 1) * sizeof(ProjRecord), stream));
        CU_TRY(cudaMalloc(&dSide, std::max<size_t>(n, 1) * sizeof(uint4)));
        if (int rc = reserve((uint64_t)std::ceil(std::max(1.0, instancesPerGaussian) * std::max<uint32_t>(n, 1)))) return rc;
        CU_TRY(cudaStreamSynchronize(stream));
        initialized = true;
        haveSort = haveRender = false;
        return ok();
    }

    int initFromHost(const GaussianCloud& gc, bool srgbIn)
    {
        if (int rc = begin()) return rc;
        const size_t bytes = gc.GetTotalSize();
        void* dAos = nullptr;
        CU_TRY(cudaMalloc(&dAos, std::max<size_t>(bytes, 16)));
        cudaError_t e = cudaMemcpyAsync(dAos, gc.GetRawDataPtr(), bytes, cudaMemcpyHostToDevice, stream);
        int rc = 0;
        if (e != cudaSuccess) rc = fail(SPL_ERR_CUDA, std::string("H2D cloud copy: ") + cudaGetErrorString(e));
        else rc = initFromDevice(dAos, gc.GetNumGaussians(), gc.GetStride(), gc.HasFullSH() && gc.GetStride() == 244, srgbIn);
        cudaStreamSynchronize(stream);
        cudaFree(dAos);
        return rc;
    }

    // multi-GPU start-up: metadata and the AoS travel by ncclBroadcast (NVLink 5 / NVSwitch), device memory to device memory
    int initBroadcast(SplatComm& comm, const GaussianCloud* cloudOnRoot, int root, bool srgbIn, float* broadcastMs)
    {
        if (int rc = begin()) return rc;
        if (!comm.Ok()) return fail(SPL_ERR_ARG, "InitBroadcast: communicator not initialised");
        if (comm.Device() != device) return fail(SPL_ERR_ARG, "InitBroadcast: the renderer is configured for device " + std::to_string(device) + ", the communicator for " + std::to_string(comm.Device()));
        const bool isRoot = comm.Rank() == root;
        if (isRoot && !cloudOnRoot) return fail(SPL_ERR_ARG, "InitBroadcast: the root rank must pass the cloud");
        unsigned long long meta[4] = {0, 0, 0, 0};
        if (isRoot)
        {
            meta[0] = cloudOnRoot->GetNumGaussians();
            meta[1] = cloudOnRoot->GetStride();
            meta[2] = (cloudOnRoot->HasFullSH() && cloudOnRoot->GetStride() == 244) ? 1 : 0;
            meta[3] = 0x53504c4154ull;  // "SPLAT"
        }
        unsigned long long* dMeta = nullptr;
        CU_TRY(cudaMalloc(&dMeta, sizeof meta));
        int rc = 0;
        void* dAos = nullptr;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        auto run = [&]() -> int {
            CU_TRY(cudaMemcpyAsync(dMeta, meta, sizeof meta, cudaMemcpyHostToDevice, stream));
            if (!comm.Broadcast(dMeta, sizeof meta, root, stream)) return fail(SPL_ERR_CUDA, comm.LastError());
            CU_TRY(cudaMemcpyAsync(meta, dMeta, sizeof meta, cudaMemcpyDeviceToHost, stream));
            CU_TRY(cudaStreamSynchronize(stream));
            if (meta[3] != 0x53504c4154ull) return fail(SPL_ERR_ARG, "InitBroadcast: the root rank did not provide a cloud");
            const size_t count = (size_t)meta[0], stride = (size_t)meta[1];
            if (count >= (1ull << 31) || (stride != 100 && stride != 244)) return fail(SPL_ERR_ARG, "InitBroadcast: bad cloud metadata");
            const size_t bytes = count * stride;
            CU_TRY(cudaMalloc(&dAos, std::max<size_t>(bytes, 16)));
            if (isRoot && bytes) CU_TRY(cudaMemcpyAsync(dAos, cloudOnRoot->GetRawDataPtr(), bytes, cudaMemcpyHostToDevice, stream));
            CU_TRY(cudaEventCreate(&e0));
            CU_TRY(cudaEventCreate(&e1));
            CU_TRY(cudaEventRecord(e0, stream));
            if (bytes && !comm.Broadcast(dAos, bytes, root, stream)) return fail(SPL_ERR_CUDA, comm.LastError());
            CU_TRY(cudaEventRecord(e1, stream));
            CU_TRY(cudaStreamSynchronize(stream));
            if (broadcastMs) CU_TRY(cudaEventElapsedTime(broadcastMs, e0, e1));
            return initFromDevice(dAos, count, stride, meta[2] != 0, srgbIn);
        };
        rc = run();
        cudaStreamSynchronize(stream);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        cudaFree(dAos);
        cudaFree(dMeta);
        return rc;
    }

    static bool sameArgs(const float* a, const float* cam, const float* proj, const float* vp, const float* nf)
    {
        return !std::memcmp(a, cam, 64) && !std::memcmp(a + 16, proj, 64) && !std::memcmp(a + 32, vp, 16) && !std::memcmp(a + 36, nf, 8);
    }

    // default path, second half: expand the depth-ordered gaussian ids into (tile id, id) instances and sort them by
    // tile (stable passes keep the depth order); the last pass writes the tile ranges.  `order` is dValsA right after the
    // depth passes, or dOrder when a shared-sort Render re-projects for another eye.
    int expandAndSortTiles(const uint32_t* order, uint32_t numTiles)
    {
        const int tileBits = sortBits - kDepthBits, tilePasses = (tileBits + kRadixBits - 1) / kRadixBits;
        uint32_t* keysA32 = reinterpret_cast<uint32_t*>(dKeysA);
        uint32_t* keysB32 = reinterpret_cast<uint32_t*>(dKeysB);
        CU_TRY(launch_emit_sorted(order, reinterpret_cast<const uint2*>(dSide), n, fc, keysB32, dValsB, dCtr, dScan, (uint32_t)capacity,
                                  dHist + 4 * kRadixSize, stream, usePdl(1)));
        launches += n ? 1 : 0;
        if (profile) CU_TRY(cudaEventRecord(ev[6], stream));
        dSortedKeys = dKeysB;
        dSortedVals = dValsB;
        sortedKeyBytes = 4;
        if (tilePasses > 0)
        {
            SortWorkspace wsTile{dHist, dStatus, dCtr->sortTickets, capacity, sortGeneration, instKpt(), 4, rangesFused ? dRanges : nullptr, numTiles};
            wsTile.allHistogramsReady = tilePasses <= 2;  // k_emit_sorted left the histograms of both tile digits
            wsTile.pdl = usePdl(2);
            wsTile.groupAgg = dGroups + 4 * sort_group_words(std::max<uint32_t>(n, 1), depthKpt());
            wsTile.groupWordsPerPass = sort_group_words(std::max<uint64_t>(capacity, 1), instKpt());
            CU_TRY(launch_sort_passes32(keysB32, dValsB, keysA32, dValsA, &dCtr->usedInstances, capacity, 0, tileBits, wsTile, numSMs, stream));
            launches += tilePasses;
            if (tilePasses & 1) { dSortedKeys = dKeysA; dSortedVals = dValsA; }
        }
        return 0;
    }

    // Public entry points = poll the frame reports, do the work, and hand back SPL_ERR_OVERFLOW as a NOTIFICATION when an
    // earlier asynchronous frame turned out to be incomplete (the call itself has been carried out, with the capacity raised).
    int sort(const float* cam, const float* proj, const float* vp, const float* nf)
    {
        std::string rep;
        const int incomplete = pollReports(&rep);
        if (int rc = sortImpl(cam, proj, vp, nf)) return rc;
        return incomplete ? fail(SPL_ERR_OVERFLOW, rep) : ok();
    }
    int render(const float* cam, const float* proj, const float* vp, const float* nf)
    {
        std::string rep;
        const int incomplete = pollReports(&rep);
        if (int rc = renderImpl(cam, proj, vp, nf)) return rc;
        return incomplete ? fail(SPL_ERR_OVERFLOW, rep) : ok();
    }

    // the device work of one Sort: control-block memset + projection + sort passes (+ tile ranges), enqueued on `stream`
    int enqueueSortWork(uint32_t numTiles)
    {
        launches = 0;
        frameSortMode = sortMode;
        rangesFused = false;
        if (profile) CU_TRY(cudaEventRecord(ev[0], stream));
        {
            const int tilePassesNow = (sortBits - kDepthBits + kRadixBits - 1) / kRadixBits;
            const size_t clearBytes = frameSortMode == SPL_SORT_BINNED ? controlBytes
                                      : frameSortMode == SPL_SORT_LSD_FLAT ? frameControlBytes(4, sortPasses) : frameControlBytes(4, tilePassesNow);
            CU_TRY(cudaMemsetAsync(dControl, 0, clearBytes, stream));
        }
        if (frameSortMode == SPL_SORT_BINNED)
        {
            // MSD by tile: count -> scan (= tile ranges) -> scatter into segments -> per-segment shared-memory sort
            uint32_t* tileCount = dTileCtl;
            uint32_t* tileCursor = dTileCtl + rangesTiles;
            uint32_t* largeList = dTileCtl + 2 * (size_t)rangesTiles;
            CU_TRY(cudaMemsetAsync(tileCount, 0, (size_t)numTiles * sizeof(uint32_t), stream));
            nvtxRangePushA("pre-sort");
            const cudaError_t pe = launch_project_bin(cloud, fc, dRec, dSide, tileCount, dCtr, stream);
            nvtxRangePop();
            CU_TRY(pe);
            NvtxScope sortZone("sort");
            if (profile) CU_TRY(cudaEventRecord(ev[1], stream));
            CU_TRY(launch_tile_scan(tileCount, numTiles, dRanges, tileCursor, largeList, dCtr, (uint32_t)capacity, stream));
            CU_TRY(launch_bin_instances(dSide, n, fc.tilesX, tileCursor, dKeysA, (uint32_t)capacity, stream));
            if (profile) CU_TRY(cudaEventRecord(ev[6], stream));
            CU_TRY(launch_segment_sort(dKeysA, dKeysB, dRanges, numTiles, largeList, dCtr, numSMs, stream));
            if (profile) CU_TRY(cudaEventRecord(ev[2], stream));
            if (profile) CU_TRY(cudaEventRecord(ev[3], stream));
            launches += (n ? 2 : 0) + 3;
        }
        else
        {
            // LSD: device-wide onesweep radix sort of the packed (tile << 32 | depthKey) keys, then boundary detection.
            // look-back words are generation-tagged instead of cleared; re-zero only when the 6-bit tag wraps
            if (((++sortGeneration) & kSortGenerationMask) == 0)
            {
                CU_TRY(cudaMemsetAsync(dStatus, 0, sort_status_bytes(capacity, kMaxSortPasses), stream));
                ++sortGeneration;
            }
            // with more than one tile the last pass sorts on a tile-id digit and writes the tile ranges itself
            rangesFused = numTiles > 1;
            if (frameSortMode == SPL_SORT_LSD_FLAT)
            {
                // every digit pass runs on the D tile instances (emitted in gaussian order by the projection kernel,
                // which also produces the first digit's histogram; pass p produces pass p+1's)
                nvtxRangePushA("pre-sort");
                const cudaError_t pe = launch_project(cloud, fc, dRec, dKeysA, dValsA, dCtr, dScan, (uint32_t)capacity, dHist, stream);
                nvtxRangePop();
                CU_TRY(pe);
                NvtxScope sortZone("sort");
                launches += n ? 1 : 0;
                if (profile) CU_TRY(cudaEventRecord(ev[1], stream));
                if (profile) CU_TRY(cudaEventRecord(ev[6], stream));
                SortWorkspace ws{dHist, dStatus, dCtr->sortTickets, capacity, sortGeneration, instKpt(), 0, rangesFused ? dRanges : nullptr, numTiles};
                ws.groupAgg = dGroups + 4 * sort_group_words(std::max<uint32_t>(n, 1), depthKpt());  // same layout as the split path
                ws.groupWordsPerPass = sort_group_words(std::max<uint64_t>(capacity, 1), instKpt());
                CU_TRY(launch_sort_passes(dKeysA, dValsA, dKeysB, dValsB, &dCtr->usedInstances, capacity, 0, sortBits, ws, numSMs, stream));
                launches += sortPasses;
                sortedKeyBytes = 8;
                dSortedKeys = (sortPasses & 1) ? dKeysB : dKeysA;
                dSortedVals = (sortPasses & 1) ? dValsB : dValsA;
            }
            else
            {
                // the four depth digits are common to all instances of a gaussian: those passes run on the N
                // (u32 depthKey, id) pairs BEFORE tile expansion; the expansion emits (u32 tile id, id) in depth order; only
                // the tile digits are sorted on the D instances, which do not carry the depth key (stable passes keep the
                // depth order).  Same stable LSD sort of the same (tile | depth) order, half the element-passes at 8 instead
                // of 12 bytes per element.  The u32 arrays live in the u64 key buffers.
                uint32_t* keysA32 = reinterpret_cast<uint32_t*>(dKeysA);
                uint32_t* keysB32 = reinterpret_cast<uint32_t*>(dKeysB);
                nvtxRangePushA("pre-sort");  // the reference's zone names (src/splatrenderer.cpp:172,208)
                const cudaError_t pe = launch_project_keys(cloud, fc, dRec, keysA32, dValsA, reinterpret_cast<uint2*>(dSide), dCtr, dHist, stream);
                nvtxRangePop();
                CU_TRY(pe);
                launches += n ? 1 : 0;
                if (profile) CU_TRY(cudaEventRecord(ev[1], stream));
                NvtxScope sortZone("sort");
                SortWorkspace wsDepth{dHist, dStatus, dCtr->sortTickets, capacity, sortGeneration, depthKpt(), 0};
                wsDepth.allHistogramsReady = true;  // k_project_keys left all four depth-digit histograms
                wsDepth.pdl = usePdl(0);
                wsDepth.groupAgg = dGroups;
                wsDepth.groupWordsPerPass = sort_group_words(std::max<uint32_t>(n, 1), depthKpt());
                CU_TRY(launch_sort_passes32(keysA32, dValsA, keysB32, dValsB, &dCtr->sortCount, capacity, 0, kDepthBits, wsDepth, numSMs, stream));
                launches += 4;
                // 4 passes: the sorted ids are back in dValsA
                if (fc.stereoPass == 1)
                {
                    // the tile passes reuse dValsA: keep the depth order for the other eye's Render
                    CU_TRY(cudaMemcpyAsync(dOrder, dValsA, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
                }
                if (int rc = expandAndSortTiles(dValsA, numTiles)) return rc;
            }
            if (profile) CU_TRY(cudaEventRecord(ev[2], stream));
            if (!rangesFused)
            {
                CU_TRY(launch_tile_ranges(dSortedKeys, sortedKeyBytes, &dCtr->usedInstances, capacity, dRanges, numTiles, numSMs, stream));
                launches += 1;
            }
            if (profile) CU_TRY(cudaEventRecord(ev[3], stream));
        }
        return 0;
    }

    int sortImpl(const float* cam, const float* proj, const float* vp, const float* nf)
    {
        if (int rc = begin()) return rc;
        if (!initialized) return fail(SPL_ERR_ARG, "Sort before Init");
        if (growTo > capacity)
        {
            if (int rc = reserve(growTo)) return rc;  // an earlier frame overflowed (pollReports)
        }
        growTo = 0;
        if (vp[0] != 0.0f || vp[1] != 0.0f) return fail(SPL_ERR_ARG, "viewport origin must be (0,0) (the reference assumes it, SURVEY q6)");
        const int W = (int)vp[2], H = (int)vp[3];
        if (W <= 0 || H <= 0 || (float)W != vp[2] || (float)H != vp[3]) return fail(SPL_ERR_ARG, "viewport width/height must be positive integers");
        const int tilesX = (W + kTileW - 1) / kTileW, tilesY = (H + kTileH - 1) / kTileH;
        if (tilesX > 1024 || tilesY > 1024) return fail(SPL_ERR_ARG, "viewport larger than 32768 x 16384 pixels");

        // frame constants, computed as the reference's CPU side does (float, glm operation order)
        const mat4 camera = mat4::from(cam), projM = mat4::from(proj);
        const mat4 view = inverse(camera);          // src/splatrenderer.cpp:161,327
        const mat4 mvp = mul(projM, view);          // src/splatrenderer.cpp:175
        std::memcpy(fc.view, view.m, 64);
        std::memcpy(fc.mvp, mvp.m, 64);
        std::memcpy(fc.proj, projM.m, 64);
        fc.eye[0] = camera.at(3, 0); fc.eye[1] = camera.at(3, 1); fc.eye[2] = camera.at(3, 2);  // :328
        fc.W = vp[2]; fc.H = vp[3]; fc.zNear = nf[0]; fc.zFar = nf[1];
        fc.width = W; fc.height = H; fc.tilesX = tilesX; fc.tilesY = tilesY;
        fc.srgb = srgb ? 1 : 0;
        fc.zFarSort = nf[1];
        fc.stereoPass = (sharedSort && sortMode == SPL_SORT_LSD) ? 1 : 0;
        std::memcpy(lastArgs, cam, 64); std::memcpy(lastArgs + 16, proj, 64); std::memcpy(lastArgs + 32, vp, 16); std::memcpy(lastArgs + 36, nf, 8);
        std::memcpy(drawArgs, lastArgs, sizeof drawArgs);

        const uint32_t numTiles = (uint32_t)tilesX * tilesY;
        if (numTiles > rangesTiles)
        {
            CU_TRY(cudaStreamSynchronize(stream));
            cudaFree(dTileCtl);
            dTileCtl = nullptr;
            if (int rc = allocControl(numTiles)) return rc;
            CU_TRY(cudaMalloc(&dTileCtl, (size_t)numTiles * 3 * sizeof(uint32_t)));
            rangesTiles = numTiles;
        }
        int tileBits = 0;
        while ((1u << tileBits) < numTiles) tileBits++;
        sortBits = kDepthBits + tileBits;
        sortPasses = (sortBits + kRadixBits - 1) / kRadixBits;

        if (profile && !evCreated)
        {
            for (auto& e : ev) CU_TRY(cudaEventCreate(&e));
            evCreated = true;
        }
        // SPL_OPT_GRAPH: the frame's memset + kernels are captured and replayed as one CUDA graph (parameters change every
        // frame, so the capture is repeated and the executable graph UPDATED, not re-instantiated); the first frame after a
        // configuration change runs directly (lazy one-time CUDA calls -- occupancy queries, function attributes -- must
        // not land inside a capture)
        if (fc.stereoPass == 1 && !dOrder) CU_TRY(cudaMalloc(&dOrder, std::max<size_t>(n, 1) * sizeof(uint32_t)));  // not inside a capture
        bool useGraph = graphOption && !profile && warmFrames >= 1;
        if (useGraph)
        {
            // a caller that is itself capturing this stream gets the plain launches inside ITS graph
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) useGraph = false;
            cudaGetLastError();
        }
        if (useGraph)
        {
            CU_TRY(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
            const int rc = enqueueSortWork(numTiles);
            cudaGraph_t g = nullptr;
            const cudaError_t ce = cudaStreamEndCapture(stream, &g);  // always: the stream must leave capture mode
            if (rc) { if (g) cudaGraphDestroy(g); return rc; }
            if (ce != cudaSuccess) return fail(SPL_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
            bool haveExec = false;
            const uint64_t shape = graphShape();
            if (sortExec && sortExecShape != shape)
            {
                cudaGraphExecDestroy(sortExec);
                sortExec = nullptr;
            }
            sortExecShape = shape;
            if (sortExec)
            {
                cudaGraphExecUpdateResultInfo info{};
                haveExec = cudaGraphExecUpdate(sortExec, g, &info) == cudaSuccess;
                if (!haveExec)
                {
                    cudaGetLastError();  // a shape change graphShape() does not know of: instantiate again
                    cudaGraphExecDestroy(sortExec);
                    sortExec = nullptr;
                }
            }
            cudaError_t ie = cudaSuccess;
            if (!haveExec) ie = cudaGraphInstantiate(&sortExec, g, 0);
            cudaGraphDestroy(g);
            if (ie != cudaSuccess) { sortExec = nullptr; return fail(SPL_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie)); }
            CU_TRY(cudaGraphLaunch(sortExec, stream));
        }
        else
        {
            if (int rc = enqueueSortWork(numTiles)) return rc;
        }
        warmFrames++;
        haveSort = true;
        return ok();
    }

    const unsigned long long* sortedKeys() const { return dSortedKeys; }
    const uint32_t* sortedVals() const { return dSortedVals; }

    // Render for another eye in the order of the last Sort (src/app.cpp:603-607: Sort for eye 0, Render for both).
    // The reference's draw re-runs its vertex and geometry stages with the new matrices over the list the Sort left;
    // here: projection again (records + tile rectangles, no keys), then expansion and tile passes over the kept order.
    int reproject(const float* cam, const float* proj, const float* vp, const float* nf)
    {
        const mat4 camera = mat4::from(cam), projM = mat4::from(proj);
        const mat4 view = inverse(camera);
        std::memcpy(fc.view, view.m, 64);  // fc.mvp and fc.zFarSort stay the Sort's
        std::memcpy(fc.proj, projM.m, 64);
        fc.eye[0] = camera.at(3, 0); fc.eye[1] = camera.at(3, 1); fc.eye[2] = camera.at(3, 2);
        fc.zNear = nf[0]; fc.zFar = nf[1];
        fc.stereoPass = 2;
        std::memcpy(drawArgs, cam, 64); std::memcpy(drawArgs + 16, proj, 64); std::memcpy(drawArgs + 32, vp, 16); std::memcpy(drawArgs + 36, nf, 8);

        const uint32_t numTiles = (uint32_t)fc.tilesX * fc.tilesY;
        if (((++sortGeneration) & kSortGenerationMask) == 0)
        {
            CU_TRY(cudaMemsetAsync(dStatus, 0, sort_status_bytes(capacity, kMaxSortPasses), stream));
            ++sortGeneration;
        }
        launches = 0;
        if (profile) CU_TRY(cudaEventRecord(ev[0], stream));
        CU_TRY(cudaMemsetAsync(dControl, 0, frameControlBytes(4, (sortBits - kDepthBits + kRadixBits - 1) / kRadixBits), stream));  // counters, histograms, scan state, tile ranges, group words
        CU_TRY(launch_project_keys(cloud, fc, dRec, nullptr, nullptr, reinterpret_cast<uint2*>(dSide), dCtr, dHist, stream));
        launches += n ? 1 : 0;
        if (profile) CU_TRY(cudaEventRecord(ev[1], stream));
        if (int rc = expandAndSortTiles(dOrder, numTiles)) return rc;
        if (profile) CU_TRY(cudaEventRecord(ev[2], stream));
        if (!rangesFused)
        {
            CU_TRY(launch_tile_ranges(dSortedKeys, sortedKeyBytes, &dCtr->usedInstances, capacity, dRanges, numTiles, numSMs, stream));
            launches += 1;
        }
        if (profile) CU_TRY(cudaEventRecord(ev[3], stream));
        return 0;
    }

    int renderImpl(const float* cam, const float* proj, const float* vp, const float* nf)
    {
        if (int rc = begin()) return rc;
        if (!initialized) return fail(SPL_ERR_ARG, "Render before Init");
        // The reference's Render re-projects with its own arguments in the order of the last Sort (the VR path
        // sorts for eye 0 only, src/app.cpp:603-607).  Here projection lives in Sort: by default a Render whose
        // arguments differ from those the tile lists were built for sorts again (the image is exactly that of the
        // arguments given); with SPL_OPT_SHARED_SORT it does what the reference does.
        if (!haveSort)
        {
            if (int rc = sortImpl(cam, proj, vp, nf)) return rc;
        }
        else if (!sameArgs(drawArgs, cam, proj, vp, nf))
        {
            // shared sort (the reference's behaviour): keep the last Sort's list and order, re-project for these
            // arguments.  Needs the default sort path and the Sort's viewport (same tile grid); otherwise re-sort.
            const bool canShare = sharedSort && fc.stereoPass != 0 && frameSortMode == SPL_SORT_LSD && !std::memcmp(lastArgs + 32, vp, 16);
            if (int rc = canShare ? reproject(cam, proj, vp, nf) : sortImpl(cam, proj, vp, nf)) return rc;
        }
        const size_t need = (size_t)fc.width * fc.height * (fbFormat == SPL_FB_FP16 ? 8 : 16);
        if (need > fbBytes)
        {
            CU_TRY(cudaStreamSynchronize(stream));
            if (copyStream) CU_TRY(cudaStreamSynchronize(copyStream));
            for (int k = 0; k < 2; k++)
            {
                cudaFree(dFbs[k]);
                dFbs[k] = nullptr;
                CU_TRY(cudaMalloc(&dFbs[k], need));
                copyPending[k] = false;
            }
            fbBytes = need;
        }
        if (!copyStream)
        {
            CU_TRY(cudaStreamCreateWithFlags(&copyStream, cudaStreamNonBlocking));
            for (int k = 0; k < 2; k++)
            {
                CU_TRY(cudaEventCreateWithFlags(&evRendered[k], cudaEventDisableTiming));
                CU_TRY(cudaEventCreateWithFlags(&evCopied[k], cudaEventDisableTiming));
            }
        }
        // two render targets: frame k+1 is composited while frame k is still being copied to the host
        const int target = fbNext;
        fbNext ^= 1;
        if (copyPending[target])
        {
            CU_TRY(cudaStreamWaitEvent(stream, evCopied[target], 0));  // the copy that last read this target must be done
            copyPending[target] = false;
        }
        dFb = dFbs[target];
        fbCur = target;
        fbW = fc.width; fbH = fc.height;
        if (profile) CU_TRY(cudaEventRecord(ev[4], stream));
        const bool binned = frameSortMode == SPL_SORT_BINNED;
        const uint64_t seq = frameSeq + 1;
        const FrameReport report{dCtr, dReport + seq % kReportRing, (seq & ((1ull << (64 - kReportSeqShift)) - 1)) << kReportSeqShift};
        nvtxRangePushA("draw");  // the reference's zone name (src/splatrenderer.cpp:323)
        const cudaError_t ce = launch_composite(dRanges, binned ? reinterpret_cast<const uint32_t*>(dKeysA) : sortedVals(), binned ? 2 : 1, dRec, dFb, fbFormat, fc.width,
                                                fc.height, fc.tilesX, fc.tilesY, (float)earlyOutT, subblockCull, rangesFused ? kFusedRangeBeginXor : 0u, report, stream,
                                                usePdl(3) && frameSortMode == SPL_SORT_LSD);
        nvtxRangePop();
        CU_TRY(ce);
        frameSeq = seq;
        seqCapacity[seq % kReportRing] = capacity;
        launches += 1;
        if (profile) CU_TRY(cudaEventRecord(ev[5], stream));
        CU_TRY(cudaEventRecord(evRendered[target], stream));
        haveRender = true;
        return ok();
    }

    // enqueue (on the render stream) a wait for every framebuffer copy still in flight on the copy stream
    int fenceCopies()
    {
        if (int rc = begin()) return rc;
        for (int k = 0; k < 2; k++)
            if (copyPending[k]) CU_TRY(cudaStreamWaitEvent(stream, evCopied[k], 0));
        return ok();
    }

    // block the HOST until at most `keepInFlight` (0 or 1) framebuffer copies are still pending; with 1, the copy
    // of the frame before the current one has landed, so a consumer can process it while this frame's copy runs
    int waitCopies(int keepInFlight)
    {
        if (int rc = begin()) return rc;
        if (keepInFlight < 0 || keepInFlight > 1) return fail(SPL_ERR_ARG, "keep_in_flight must be 0 or 1");
        const int older = fbCur ^ 1;
        if (copyPending[older])
        {
            CU_TRY(cudaEventSynchronize(evCopied[older]));
            copyPending[older] = false;
        }
        if (keepInFlight == 0 && copyPending[fbCur])
        {
            CU_TRY(cudaEventSynchronize(evCopied[fbCur]));
            copyPending[fbCur] = false;
        }
        // the frames whose pixels have just landed have also posted their reports
        std::string rep;
        if (pollReports(&rep)) return fail(SPL_ERR_OVERFLOW, rep);
        return ok();
    }

    int readCounters(FrameCounters* out)
    {
        CU_TRY(cudaMemcpyAsync(out, dCtr, sizeof(FrameCounters), cudaMemcpyDeviceToHost, stream));
        CU_TRY(cudaStreamSynchronize(stream));
        return 0;
    }

    // rgba8 = the presented frame: what the reference's RenderDesktop blit (src/app.cpp:1076-1082,
    // shader/desktop_frag.glsl:32-37 with color = (1,1,1,1)) leaves in the 8-bit window: clamp to [0,1], x255, round
    int readFramebuffer(void* host, size_t bytes, bool wait, bool rgba8 = false)
    {
        if (int rc = begin()) return rc;
        if (!haveRender) return fail(SPL_ERR_ARG, "no frame rendered yet");
        const size_t need = (size_t)fbW * fbH * (rgba8 ? 4 : fbFormat == SPL_FB_FP16 ? 8 : 16);
        if (bytes < need) return fail(SPL_ERR_ARG, "host buffer too small for the framebuffer");
        if (wait)
        {
            // a frame that ran out of key capacity is incomplete: grow and redo it once
            FrameCounters c{};
            if (int rc = readCounters(&c)) return rc;
            if (c.overflow)
            {
                if (c.totalInstances > kMaxInstances) return fail(SPL_ERR_OVERFLOW, "frame needs more than 2^30 tile instances");
                if (int rc = reserve(c.totalInstances + c.totalInstances / 4)) return rc;
                float a[38], d[38];
                std::memcpy(a, lastArgs, sizeof a);
                std::memcpy(d, drawArgs, sizeof d);
                haveSort = false;
                if (int rc = sortImpl(a, a + 16, a + 32, a + 36)) return rc;
                if (int rc = renderImpl(d, d + 16, d + 32, d + 36)) return rc;  // re-projects when the frame was another eye's
                forgetReports();  // the incomplete frame has been replaced: nothing to notify
            }
        }
        // device -> host on the copy stream, ordered after the composite of this target by an event; the next
        // Render writes the other target, so it overlaps this copy
        const void* src = dFb;
        if (rgba8)
        {
            if (presentBytes != need)
            {
                CU_TRY(cudaStreamSynchronize(stream));
                CU_TRY(cudaStreamSynchronize(copyStream));
                for (int k = 0; k < 2; k++)
                {
                    cudaFree(dPresent[k]);
                    dPresent[k] = nullptr;
                    CU_TRY(cudaMalloc(&dPresent[k], need));
                }
                presentBytes = need;
            }
            // one 8-bit image per render target: the next Render into this target already waits for evCopied[fbCur]
            CU_TRY(launch_present_rgba8(dFb, fbFormat, static_cast<uint32_t*>(dPresent[fbCur]), (size_t)fbW * fbH, stream));
            launches += 1;
            CU_TRY(cudaEventRecord(evRendered[fbCur], stream));
            src = dPresent[fbCur];
        }
        CU_TRY(cudaStreamWaitEvent(copyStream, evRendered[fbCur], 0));
        CU_TRY(cudaMemcpyAsync(host, src, need, cudaMemcpyDeviceToHost, copyStream));
        CU_TRY(cudaEventRecord(evCopied[fbCur], copyStream));
        copyPending[fbCur] = true;
        if (wait) CU_TRY(cudaStreamSynchronize(copyStream));
        std::string rep;
        if (pollReports(&rep)) return fail(SPL_ERR_OVERFLOW, rep);  // notification (the copy has been enqueued / done)
        return ok();
    }

    int stats(spl_stats* out)
    {
        if (int rc = begin()) return rc;
        if (!initialized) return fail(SPL_ERR_ARG, "stats before Init");
        std::memset(out, 0, sizeof *out);
        FrameCounters c{};
        if (int rc = readCounters(&c)) return rc;
        out->num_gaussians = n;
        out->num_visible = c.visible;
        out->num_instances = c.totalInstances;
        out->instance_capacity = capacity;
        out->overflow = c.overflow;
        out->tiles_x = fc.tilesX; out->tiles_y = fc.tilesY; out->tile_w = kTileW; out->tile_h = kTileH;
        out->sort_passes = frameSortMode == SPL_SORT_BINNED ? 1 : sortPasses;  // LSD: 4 depth-digit passes + tile-digit passes
        out->sort_bits = sortBits;
        out->sort_mode = frameSortMode;
        out->sh_mode = cloud.shMode;
        out->kernel_launches = launches;
        if (profile && evCreated && haveSort)
        {
            float ms = 0;
            // stage_ms: [0] project(+memset) [1] sort histogram [2] sort scatter passes [3] tile ranges [4] composite [7] total
            if (cudaEventElapsedTime(&ms, ev[0], ev[1]) == cudaSuccess) out->stage_ms[0] = ms;
            if (cudaEventElapsedTime(&ms, ev[1], ev[6]) == cudaSuccess) out->stage_ms[1] = ms;
            if (cudaEventElapsedTime(&ms, ev[6], ev[2]) == cudaSuccess) out->stage_ms[2] = ms;
            if (cudaEventElapsedTime(&ms, ev[2], ev[3]) == cudaSuccess) out->stage_ms[3] = ms;
            if (haveRender)
            {
                if (cudaEventElapsedTime(&ms, ev[4], ev[5]) == cudaSuccess) out->stage_ms[4] = ms;
                if (cudaEventElapsedTime(&ms, ev[0], ev[5]) == cudaSuccess) out->stage_ms[7] = ms;
            }
        }
        return ok();
    }

    int tap(int what, void* host, size_t bytes)
    {
        if (int rc = begin()) return rc;
        if (!haveSort) return fail(SPL_ERR_ARG, "tap before Sort");
        FrameCounters c{};
        if (int rc = readCounters(&c)) return rc;
        const void* src = nullptr;
        size_t have = 0;
        if (frameSortMode == SPL_SORT_BINNED && (what == SPL_TAP_KEYS || what == SPL_TAP_VALS))
        {
            // binned sort keeps (depthKey << 32 | id) per tile segment: rebuild the LSD view (tile << 32 | depthKey), id
            const size_t d = c.usedInstances, tiles = (size_t)fc.tilesX * fc.tilesY;
            if (bytes < d * (what == SPL_TAP_KEYS ? 8 : 4)) return fail(SPL_ERR_ARG, "tap buffer too small");
            std::vector<unsigned long long> pairs(d);
            std::vector<uint2> rg(tiles);
            if (d) CU_TRY(cudaMemcpyAsync(pairs.data(), dKeysA, d * 8, cudaMemcpyDeviceToHost, stream));
            CU_TRY(cudaMemcpyAsync(rg.data(), dRanges, tiles * sizeof(uint2), cudaMemcpyDeviceToHost, stream));
            CU_TRY(cudaStreamSynchronize(stream));
            for (size_t t = 0; t < tiles; t++)
                for (uint32_t k = rg[t].x; k < rg[t].y; k++)
                {
                    if (what == SPL_TAP_KEYS) static_cast<unsigned long long*>(host)[k] = ((unsigned long long)t << 32) | (pairs[k] >> 32);
                    else static_cast<uint32_t*>(host)[k] = (uint32_t)pairs[k];
                }
            return ok();
        }
        if (what == SPL_TAP_KEYS && sortedKeyBytes == 4)
        {
            // default path: the instances carry bare tile ids; the tap shows the (tile << 32 | depthKey) stream they stand for
            const size_t d = c.usedInstances;
            if (bytes < d * 8) return fail(SPL_ERR_ARG, "tap buffer too small");
            std::vector<uint32_t> tiles(d), ids(d);
            std::vector<ProjRecord> recs(n);
            if (d) CU_TRY(cudaMemcpyAsync(tiles.data(), sortedKeys(), d * 4, cudaMemcpyDeviceToHost, stream));
            if (d) CU_TRY(cudaMemcpyAsync(ids.data(), sortedVals(), d * 4, cudaMemcpyDeviceToHost, stream));
            if (n) CU_TRY(cudaMemcpyAsync(recs.data(), dRec, (size_t)n * sizeof(ProjRecord), cudaMemcpyDeviceToHost, stream));
            CU_TRY(cudaStreamSynchronize(stream));
            for (size_t k = 0; k < d; k++)
                static_cast<unsigned long long*>(host)[k] = ((unsigned long long)tiles[k] << 32) | (ids[k] < n ? recs[ids[k]].key : 0u);
            return ok();
        }
        switch (what)
        {
        case SPL_TAP_RECORDS: src = dRec; have = (size_t)n * sizeof(ProjRecord); break;
        case SPL_TAP_KEYS: src = sortedKeys(); have = (size_t)c.usedInstances * 8; break;
        case SPL_TAP_VALS: src = sortedVals(); have = (size_t)c.usedInstances * 4; break;
        case SPL_TAP_TILE_RANGES: src = dRanges; have = (size_t)fc.tilesX * fc.tilesY * sizeof(uint2); break;
        default: return fail(SPL_ERR_ARG, "unknown tap");
        }
        if (bytes < have) return fail(SPL_ERR_ARG, "tap buffer too small");
        if (have) CU_TRY(cudaMemcpyAsync(host, src, have, cudaMemcpyDeviceToHost, stream));
        CU_TRY(cudaStreamSynchronize(stream));
        if (what == SPL_TAP_TILE_RANGES && rangesFused)
        {
            // decode {~begin, end}; a tile the sort never touched reads {0, 0}
            uint2* r = static_cast<uint2*>(host);
            for (size_t t = 0; t < have / sizeof(uint2); t++)
            {
                if (r[t].y == 0) r[t].x = 0;
                else r[t].x ^= kFusedRangeBeginXor;
            }
        }
        return ok();
    }
};

}  // namespace splat

using splat::RendererImpl;

SplatRenderer::SplatRenderer() : impl(new RendererImpl) {}
SplatRenderer::~SplatRenderer() = default;

#define FORWARD(call)            \
    do                           \
    {                            \
        status = (call);         \
        error = impl->error;     \
        return status == 0;      \
    } while (0)

bool SplatRenderer::Configure(int device, FrameBuffer format, void* cudaStream)
{
    if (impl->initialized) { status = SPL_ERR_ARG; error = "Configure must precede Init"; return false; }
    impl->device = device;
    impl->fbFormat = (int)format;
    if (impl->ownStream && impl->stream) { cudaSetDevice(impl->device); cudaStreamDestroy(impl->stream); impl->ownStream = false; }
    impl->stream = static_cast<cudaStream_t>(cudaStream);
    status = 0;
    return true;
}

bool SplatRenderer::Init(std::shared_ptr<GaussianCloud> gaussianCloud, bool isFramebufferSRGBEnabledIn, bool useRgcSortOverrideIn)
{
    (void)useRgcSortOverrideIn;  // both reference sorters give the same stable order; one CUDA sort replaces both
    if (!gaussianCloud) { status = SPL_ERR_ARG; error = "null cloud"; return false; }
    FORWARD(impl->initFromHost(*gaussianCloud, isFramebufferSRGBEnabledIn));
}

bool SplatRenderer::InitFromDevice(const void* deviceAos, size_t numGaussians, size_t stride, bool hasFullSH, bool isFramebufferSRGBEnabledIn)
{
    FORWARD(impl->initFromDevice(deviceAos, numGaussians, stride, hasFullSH, isFramebufferSRGBEnabledIn));
}

bool SplatRenderer::InitBroadcast(SplatComm& comm, std::shared_ptr<GaussianCloud> cloudOnRoot, int root, bool isFramebufferSRGBEnabledIn, float* broadcastMs)
{
    FORWARD(impl->initBroadcast(comm, cloudOnRoot.get(), root, isFramebufferSRGBEnabledIn, broadcastMs));
}

bool SplatRenderer::InitFromPly(const std::string& plyFilename, bool importFullSH, bool isFramebufferSRGBEnabledIn)
{
    auto failWith = [&](int code, const std::string& msg) { status = code; error = msg; return false; };
    std::ifstream file(plyFilename, std::ios::binary);
    if (!file.is_open()) return failWith(SPL_ERR_IO, "failed to open " + plyFilename);
    Ply ply;
    if (!ply.Parse(file)) return failWith(SPL_ERR_IO, "error parsing ply file \"" + plyFilename + "\": " + ply.GetError());
    splat::PlyDecodeLayout layout;
    bool fullSH = false;
    splat::make_ply_decode_layout(ply, importFullSH, &layout, &fullSH);
    if (int rc = impl->begin()) { status = rc; error = impl->error; return false; }
    void* dAos = nullptr;
    std::string err;
    if (!splat::decode_ply_on_device(ply.GetData(), ply.GetVertexCount(), ply.GetVertexSize(), layout, fullSH, impl->device, impl->stream, nullptr, &dAos,
                                     &err))
        return failWith(SPL_ERR_CUDA, "GPU ply decode of \"" + plyFilename + "\" failed: " + err);
    const int rc = impl->initFromDevice(dAos, ply.GetVertexCount(), fullSH ? 244 : 100, fullSH, isFramebufferSRGBEnabledIn);
    splat::free_decoded_aos(dAos);
    status = rc;
    error = impl->error;
    return rc == 0;
}

void SplatRenderer::Sort(const splat::mat4& cameraMat, const splat::mat4& projMat, const splat::vec4& viewport, const splat::vec2& nearFar)
{
    status = impl->sort(cameraMat.m, projMat.m, &viewport.x, &nearFar.x);
    error = impl->error;
    if (status) std::fprintf(stderr, "SplatRenderer::Sort: %s\n", error.c_str());  // the reference logs and carries on
}

void SplatRenderer::Render(const splat::mat4& cameraMat, const splat::mat4& projMat, const splat::vec4& viewport, const splat::vec2& nearFar)
{
    status = impl->render(cameraMat.m, projMat.m, &viewport.x, &nearFar.x);
    error = impl->error;
    if (status) std::fprintf(stderr, "SplatRenderer::Render: %s\n", error.c_str());
}

bool SplatRenderer::GetFramebuffer(void** devicePtr, int* width, int* height, size_t* rowPitchBytes, FrameBuffer* format) const
{
    if (!impl->haveRender) return false;
    if (devicePtr) *devicePtr = impl->dFb;
    if (width) *width = impl->fbW;
    if (height) *height = impl->fbH;
    if (rowPitchBytes) *rowPitchBytes = (size_t)impl->fbW * (impl->fbFormat == SPL_FB_FP16 ? 8 : 16);
    if (format) *format = (FrameBuffer)impl->fbFormat;
    return true;
}

bool SplatRenderer::ReadFramebufferAsync(void* host, size_t bytes) { FORWARD(impl->readFramebuffer(host, bytes, false)); }
bool SplatRenderer::ReadFramebuffer(void* host, size_t bytes) { FORWARD(impl->readFramebuffer(host, bytes, true)); }

bool SplatRenderer::ReadPresented8Async(void* host, size_t bytes) { FORWARD(impl->readFramebuffer(host, bytes, false, true)); }

bool SplatRenderer::ReadPresented8(void* host, size_t bytes) { FORWARD(impl->readFramebuffer(host, bytes, true, true)); }

bool SplatRenderer::Synchronize()
{
    if (int rc = impl->begin()) { status = rc; error = impl->error; return false; }
    cudaError_t e = cudaStreamSynchronize(impl->stream);
    if (e == cudaSuccess && impl->copyStream) e = cudaStreamSynchronize(impl->copyStream);
    if (e != cudaSuccess) { status = impl->fail(SPL_ERR_CUDA, std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e)); error = impl->error; return false; }
    std::string rep;
    if (impl->pollReports(&rep)) { status = impl->fail(SPL_ERR_OVERFLOW, rep); error = impl->error; return false; }  // notification
    status = 0;
    return true;
}

bool SplatRenderer::FenceCopies() { FORWARD(impl->fenceCopies()); }

bool SplatRenderer::WaitCopies(int keepInFlight) { FORWARD(impl->waitCopies(keepInFlight)); }

bool SplatRenderer::GetStats(spl_stats* out) { FORWARD(impl->stats(out)); }

bool SplatRenderer::Reserve(uint64_t instanceCapacity)
{
    if (int rc = impl->begin()) { status = rc; error = impl->error; return false; }
    FORWARD(impl->reserve(instanceCapacity));
}

bool SplatRenderer::SetOption(int option, double value)
{
    switch (option)
    {
    case SPL_OPT_EARLY_OUT_T: impl->earlyOutT = value; break;
    case SPL_OPT_PROFILE: impl->profile = value != 0.0; break;
    case SPL_OPT_INSTANCES_PER_GAUSSIAN: impl->instancesPerGaussian = value; break;
    case SPL_OPT_SORT_TILE:
        if (value != 0 && value != 2048 && value != 4096) { status = SPL_ERR_ARG; error = "sort tile must be 0 (automatic), 2048 or 4096 keys"; return false; }
        impl->sortKptOption = value == 0 ? 0 : (value == 2048 ? 8 : 16);
        impl->warmFrames = 0;  // other kernel variants: their one-time CUDA calls must not land inside a capture
        break;
    case SPL_OPT_SORT_MODE:
        if (value != SPL_SORT_BINNED && value != SPL_SORT_LSD && value != SPL_SORT_LSD_FLAT) { status = SPL_ERR_ARG; error = "unknown sort mode"; return false; }
        impl->sortMode = (int)value;
        impl->haveSort = false;
        impl->warmFrames = 0;
        break;
    case SPL_OPT_SHARED_SORT:
        if (value != 0 && value != 1) { status = SPL_ERR_ARG; error = "shared sort must be 0 or 1"; return false; }
        impl->sharedSort = value != 0;
        impl->haveSort = false;
        impl->warmFrames = 0;
        break;
    case SPL_OPT_GRAPH:
        if (value != 0 && value != 1) { status = SPL_ERR_ARG; error = "graph must be 0 or 1"; return false; }
        impl->graphOption = value != 0;
        impl->warmFrames = 0;
        break;
    case SPL_OPT_PDL:
        if (value < 0 || value > 15 || value != (int)value) { status = SPL_ERR_ARG; error = "pdl must be a mask 0..15"; return false; }
        impl->pdlMask = (int)value;
        impl->warmFrames = 0;
        break;
    case SPL_OPT_SUBBLOCK_CULL:
        if (value != 0 && value != 1 && value != 2) { status = SPL_ERR_ARG; error = "sub-block cull mode must be 0, 1 or 2"; return false; }
        impl->subblockCull = (int)value;
        break;
    default: status = SPL_ERR_ARG; error = "unknown option"; return false;
    }
    status = 0;
    return true;
}

bool SplatRenderer::Tap(int what, void* host, size_t bytes) { FORWARD(impl->tap(what, host, bytes)); }
