// splatcomm.h -- multi-GPU start-up for the splat hot path (BASELINE.json north_star: "views shard naturally
// one-per-GPU across the 8xB200 box with an NCCL broadcast of the gaussian buffer at startup only").
//
// The reference (hyperlogic/splatapult) is single-GPU: App::Init loads one cloud and inits one SplatRenderer
// (src/app.cpp:535-559).  With N GPUs that call site becomes, per rank r (a process or a thread per GPU):
//
//     SplatComm comm;                              // rank 0: SplatComm::GetUniqueId(id); ship id to the others
//     comm.Init(id, r, N, /*device*/ r);           // ncclCommInitRank
//     auto cloud = (r == 0) ? load the .ply : nullptr;
//     SplatRenderer renderer;  renderer.Configure(r, ...);
//     renderer.InitBroadcast(comm, cloud, /*root*/ 0, sRGB);   // ONE ncclBroadcast of the AoS, packed on the device
//     ... then Sort/Render its own camera every frame: no collective on the frame path.
//
// NCCL is loaded at run time (dlopen "libnccl.so.2"); libsplat_b200.so does not link against it.
#pragma once
#include <cstddef>
#include <string>

#include "splatmath.h"  // SPLAT_CPP_API

#ifndef SPL_COMM_ID_BYTES
#define SPL_COMM_ID_BYTES 128 /* sizeof(ncclUniqueId) */
#endif

class SPLAT_CPP_API SplatComm
{
public:
    SplatComm() = default;
    ~SplatComm();
    SplatComm(const SplatComm&) = delete;
    SplatComm& operator=(const SplatComm&) = delete;

    // on ONE rank; the 128 bytes reach the other ranks by any means (a file, a socket, MPI, torch.distributed's store ...)
    static bool GetUniqueId(unsigned char id[SPL_COMM_ID_BYTES], std::string* error = nullptr);
    static int NcclVersion();  // ncclGetVersion() of the copy that was loaded, 0 if none could be

    // collective over the nranks participants (ncclCommInitRank); `device` = the CUDA ordinal this rank renders on
    bool Init(const unsigned char id[SPL_COMM_ID_BYTES], int rank, int nranks, int device);
    void Destroy();
    // one ncclBroadcast of `bytes` of device memory from `root`, in place, enqueued on `cudaStream`
    bool Broadcast(void* deviceBuf, size_t bytes, int root, void* cudaStream);

    int Rank() const { return rank; }
    int NumRanks() const { return nranks; }
    int Device() const { return device; }
    bool Ok() const { return comm != nullptr; }
    const std::string& LastError() const { return error; }

private:
    void* comm = nullptr;  // ncclComm_t
    int rank = 0, nranks = 0, device = 0;
    std::string error;
};
