// multigpu_consumer.cpp -- the multi-GPU form of the reference's start-up call site (src/app.cpp:535-559) as a C++ host
// would write it: one thread per GPU, ONE rank imports the .ply, SplatRenderer::InitBroadcast replicates the cloud with a
// single ncclBroadcast, every rank renders its own camera from a cameras.json (CamerasConfig).  Only the C++ headers of
// include/splatapult_b200/ are used; no torch, no C ABI.
//
//   multigpu_consumer <in.ply> <cameras.json> <out prefix> <num gpus> [width height [proj.bin]]
// proj.bin = 16 float32 of the projection matrix (default: perspective(45 deg, w/h, 0.1, 1000) as at src/app.cpp:1042)
// writes <out prefix><rank>.raw = width, height (int32) + float4 pixels of rank's view; prints "multigpu ok <n> ranks ...".
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "splatapult_b200/camerasconfig.h"
#include "splatapult_b200/gaussiancloud.h"
#include "splatapult_b200/splatcomm.h"
#include "splatapult_b200/splatrenderer.h"

int main(int argc, char** argv)
{
    if (argc < 5) { std::printf("usage: multigpu_consumer <in.ply> <cameras.json> <out prefix> <num gpus> [width height [proj.bin]]\n"); return 64; }
    const std::string ply = argv[1], camerasJson = argv[2], prefix = argv[3];
    const int nranks = std::atoi(argv[4]);
    const int w = argc > 6 ? std::atoi(argv[5]) : 640, h = argc > 6 ? std::atoi(argv[6]) : 480;
    splat::mat4 proj = splat::perspective(0.78539816339f, (float)w / (float)h, 0.1f, 1000.0f);
    if (argc > 7)
    {
        FILE* fp = std::fopen(argv[7], "rb");
        if (!fp || std::fread(proj.m, sizeof(float), 16, fp) != 16) { std::printf("bad proj file\n"); return 1; }
        std::fclose(fp);
    }
    CamerasConfig cameras;
    if (!cameras.ImportJson(camerasJson)) { std::printf("cameras.json failed\n"); return 1; }
    unsigned char id[SPL_COMM_ID_BYTES];
    std::string err;
    if (!SplatComm::GetUniqueId(id, &err)) { std::printf("nccl id failed: %s\n", err.c_str()); return 2; }

    std::atomic<int> failures{0};
    std::vector<float> broadcastMs(nranks, 0.0f);
    std::vector<std::thread> threads;
    for (int rank = 0; rank < nranks; rank++)
        threads.emplace_back([&, rank] {
            auto fail = [&](const std::string& what) { std::printf("rank %d: %s\n", rank, what.c_str()); failures++; };
            SplatComm comm;
            if (!comm.Init(id, rank, nranks, /*device*/ rank)) return fail("comm init: " + comm.LastError());
            std::shared_ptr<GaussianCloud> cloud;
            if (rank == 0)  // the only host copy of the cloud in the job
            {
                cloud = std::make_shared<GaussianCloud>(GaussianCloud::Options{true, true});
                if (!cloud->ImportPly(ply)) { fail("import failed"); cloud.reset(); }
            }
            SplatRenderer r;
            r.Configure(rank, SplatRenderer::FrameBuffer::Float);
            if (!r.InitBroadcast(comm, cloud, /*root*/ 0, false, &broadcastMs[rank])) return fail("InitBroadcast: " + r.LastError());
            const Camera& cam = cameras.GetCameraVec()[rank % cameras.GetNumCameras()];
            const splat::vec4 viewport{0.0f, 0.0f, (float)w, (float)h};
            const splat::vec2 nearFar{0.1f, 1000.0f};
            r.Sort(cam.mat, proj, viewport, nearFar);
            r.Render(cam.mat, proj, viewport, nearFar);
            std::vector<float> fb((size_t)w * h * 4);
            if (!r.ReadFramebuffer(fb.data(), fb.size() * sizeof(float))) return fail("read: " + r.LastError());
            const std::string out = prefix + std::to_string(rank) + ".raw";
            FILE* f = std::fopen(out.c_str(), "wb");
            if (!f) return fail("cannot write " + out);
            const int dims[2] = {w, h};
            std::fwrite(dims, sizeof(int), 2, f);
            std::fwrite(fb.data(), sizeof(float), fb.size(), f);
            std::fclose(f);
        });
    for (auto& t : threads) t.join();
    if (failures) return 3;
    std::printf("multigpu ok %d ranks nccl %d broadcast_ms", nranks, SplatComm::NcclVersion());
    for (float ms : broadcastMs) std::printf(" %.3f", ms);
    std::printf("\n");
    return 0;
}
