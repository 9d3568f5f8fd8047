// dropin_consumer.cpp -- what a splatapult translation unit looks like after the header swap of INTEGRATION.md:
// only the C++ classes (no C ABI, no CUDA headers), linked against libsplat_b200.so.
//
//   dropin_consumer cloud <out.ply>            InitDebugCloud -> ExportPly -> ImportPly round trip, prints "cloud ok <n>"
//   dropin_consumer nogpu                      SplatRenderer::Init without a CUDA device must fail loudly, prints the error
//   dropin_consumer render <in.ply> <args.bin> <out.raw>
//                                              ImportPly, Init, Sort, Render (app.cpp:549-559,1067-1068), ReadFramebuffer;
//                                              args.bin = cameraMat[16] projMat[16] viewport[4] nearFar[2] (float32);
//                                              writes width, height (int32) + float4 pixels
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "splatapult_b200/camerasconfig.h"
#include "splatapult_b200/gaussiancloud.h"
#include "splatapult_b200/splatrenderer.h"

int main(int argc, char** argv)
{
    const std::string mode = argc > 1 ? argv[1] : "";
    GaussianCloud::Options opt{true, true};
    if (mode == "cloud" && argc > 2)
    {
        GaussianCloud a(opt);
        a.InitDebugCloud();
        if (!a.ExportPly(argv[2])) { std::printf("export failed\n"); return 1; }
        GaussianCloud b(opt);
        if (!b.ImportPly(argv[2])) { std::printf("import failed\n"); return 1; }
        if (b.GetNumGaussians() != a.GetNumGaussians() || !b.GetPosWithAlphaAttrib().Valid()) { std::printf("mismatch\n"); return 1; }
        std::printf("cloud ok %zu stride %zu\n", b.GetNumGaussians(), b.GetStride());
        return 0;
    }
    if (mode == "nogpu")
    {
        auto gc = std::make_shared<GaussianCloud>(opt);
        gc->InitDebugCloud();
        SplatRenderer r;
        const bool ok = r.Init(gc, false, false);
        std::printf("init %s: %s\n", ok ? "ok" : "failed", r.LastError().c_str());
        return ok ? 0 : 3;
    }
    if (mode == "render" && argc > 4)
    {
        float args[38];
        FILE* fa = std::fopen(argv[3], "rb");
        if (!fa || std::fread(args, sizeof(float), 38, fa) != 38) { std::printf("bad args file\n"); return 1; }
        std::fclose(fa);
        auto gc = std::make_shared<GaussianCloud>(opt);
        if (!gc->ImportPly(argv[2])) { std::printf("import failed\n"); return 1; }
        SplatRenderer r;
        if (!r.Init(gc, false, false)) { std::printf("init failed: %s\n", r.LastError().c_str()); return 2; }
        const splat::mat4 cam = splat::mat4::from(args), proj = splat::mat4::from(args + 16);
        const splat::vec4 viewport{args[32], args[33], args[34], args[35]};
        const splat::vec2 nearFar{args[36], args[37]};
        const int w = (int)viewport.z, h = (int)viewport.w;
        r.Sort(cam, proj, viewport, nearFar);
        r.Render(cam, proj, viewport, nearFar);
        std::vector<float> fb((size_t)w * h * 4);
        if (!r.ReadFramebuffer(fb.data(), fb.size() * sizeof(float))) { std::printf("read failed: %s\n", r.LastError().c_str()); return 2; }
        FILE* f = std::fopen(argv[4], "wb");
        if (!f) return 1;
        const int dims[2] = {w, h};
        std::fwrite(dims, sizeof(int), 2, f);
        std::fwrite(fb.data(), sizeof(float), fb.size(), f);
        std::fclose(f);
        std::printf("render ok\n");
        return 0;
    }
    std::printf("usage: dropin_consumer cloud <out.ply> | nogpu | render <in.ply> <args.bin> <out.raw>\n");
    return 64;
}
