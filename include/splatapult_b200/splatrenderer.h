// splatrenderer.h -- B200-native drop-in for the reference's SplatRenderer (src/splatrenderer.h:23-67).
//
// Same call protocol as the reference: Init once with the cloud, then per frame Sort(...) followed by
// Render(...) with identical arguments (src/app.cpp:1067-1068).  The OpenGL pipeline behind those calls
// (presort_compute + multi_radixsort + splat_vert/geom/frag + ROP blending) is replaced by CUDA kernels
// for sm_100a; the output is an explicit float framebuffer instead of "whatever GL target is bound".
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>

#include "gaussiancloud.h"
#include "splatmath.h"

struct spl_stats;
class SplatComm;  // splatcomm.h

namespace splat { struct RendererImpl; }

class SPLAT_CPP_API SplatRenderer
{
public:
    enum class FrameBuffer { Float = 0 /* --fp32, GL_RGBA32F */, HalfFloat = 1 /* --fp16, GL_RGBA16F */ };

    SplatRenderer();
    ~SplatRenderer();
    SplatRenderer(const SplatRenderer&) = delete;
    SplatRenderer& operator=(const SplatRenderer&) = delete;

    // optional, before Init: device ordinal, target format, external cudaStream_t (nullptr = own stream)
    bool Configure(int device, FrameBuffer format, void* cudaStream = nullptr);

    // reference signature (src/splatrenderer.h:29-30).  useRgcSortOverrideIn selected the reference's
    // fallback GLSL sorter; both sorters produce the same stable order, so it is accepted and ignored.
    bool Init(std::shared_ptr<GaussianCloud> gaussianCloud, bool isFramebufferSRGBEnabledIn, bool useRgcSortOverrideIn);
    // same, from a reference-layout AoS already resident on the device (NCCL broadcast destination)
    bool InitFromDevice(const void* deviceAos, size_t numGaussians, size_t stride, bool hasFullSH, bool isFramebufferSRGBEnabledIn);
    // GaussianCloud(opt).ImportPly(file) + Init(cloud, ...) without the host decode: header parsed on the host, the raw
    // PLY body decoded by a CUDA kernel straight into device memory (same records as ImportPly, csrc/plydecode.cu)
    bool InitFromPly(const std::string& plyFilename, bool importFullSH, bool isFramebufferSRGBEnabledIn);

    // multi-GPU start-up (one renderer per GPU, one rank per renderer): EVERY rank of `comm` calls this; the rank `root`
    // passes the imported cloud, the others nullptr.  The reference-layout AoS travels by ONE ncclBroadcast into device
    // memory and is packed there (= Init on every rank, with a single host copy of the cloud in the whole job).
    // broadcastMs (optional): device time of the ncclBroadcast.  The renderer must be Configure()d for comm.Device().
    bool InitBroadcast(SplatComm& comm, std::shared_ptr<GaussianCloud> cloudOnRoot, int root, bool isFramebufferSRGBEnabledIn,
                       float* broadcastMs = nullptr);

    // reference signatures (src/splatrenderer.h:32-37); viewport = (x, y, width, height)
    void Sort(const splat::mat4& cameraMat, const splat::mat4& projMat, const splat::vec4& viewport, const splat::vec2& nearFar);
    void Render(const splat::mat4& cameraMat, const splat::mat4& projMat, const splat::vec4& viewport, const splat::vec2& nearFar);

    // --- explicit replacements for the reference's implicit GL state ---
    bool GetFramebuffer(void** devicePtr, int* width, int* height, size_t* rowPitchBytes, FrameBuffer* format) const;
    // copies run on an internal copy stream and overlap the next frame (two render targets alternate)
    bool ReadFramebufferAsync(void* host, size_t bytes);
    bool ReadFramebuffer(void* host, size_t bytes);  // waits; transparently re-renders once on capacity overflow
    // the PRESENTED frame, 8-bit RGBA (4 bytes / pixel, row 0 = bottom): what the reference's RenderDesktop blit leaves in
    // the window (src/app.cpp:1076-1082); converted on the device, so a quarter of the float target's PCIe traffic
    bool ReadPresented8Async(void* host, size_t bytes);
    bool ReadPresented8(void* host, size_t bytes);
    bool FenceCopies();   // make the render stream wait (device-side) for every copy still in flight
    bool WaitCopies(int keepInFlight = 0);  // host waits until at most keepInFlight (0|1) async copies are pending
    bool Synchronize();   // host waits for the render stream and the copy stream
    bool GetStats(spl_stats* out);
    bool Reserve(uint64_t instanceCapacity);
    bool SetOption(int option, double value);
    bool Tap(int what, void* host, size_t bytes);

    int LastStatus() const { return status; }            // 0 or a negative SPL_ERR_* of the last call
    const std::string& LastError() const { return error; }

public:
    // kept for source compatibility (src/splatrenderer.h:39); the CUDA sort has no such tuning knob
    uint32_t numBlocksPerWorkgroup = 1024;

protected:
    std::unique_ptr<splat::RendererImpl> impl;
    int status = 0;
    std::string error;
};
