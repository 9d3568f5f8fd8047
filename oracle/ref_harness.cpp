// ref_harness.cpp -- runs the REFERENCE'S OWN SOURCE on the CPU (oracle/_ref).  TEST INFRASTRUCTURE ONLY.
//
// What is the reference's code here, and what is glue:
//   * GaussianCloud::ImportPly / InitDebugCloud, Ply::Parse, BinaryAttribute: compiled UNMODIFIED from
//     /root/reference/src/{gaussiancloud,ply}.cpp, core/{binaryattribute,log}.cpp against oracle/ref_shim's stand-ins
//     for <glm/glm.hpp> and <Eigen/Dense> (those libraries are not in the image);
//   * shader/presort_compute.glsl, splat_vert.glsl, splat_geom.glsl, splat_frag.glsl: their source text is included
//     below (gen_shaders.inc, produced by oracle/build_ref.py with a handful of declaration rewrites) and compiled as
//     C++ against oracle/ref_shim/glsl_shim.h -- every formula, constant, sign and index mapping in the four
//     shaders is the reference's;
//   * glue written here, mirroring the reference's host calls: the uniforms SplatRenderer::Sort/Render set
//     (src/splatrenderer.cpp:160-177,324-335), the (x,y,z,1) position buffer (:106-111), a stable sort standing in
//     for the GLSL radix sort (same contract: stable, ascending), point-in-quad coverage standing in for the
//     fixed-function rasteriser, and the blend/clear state of App::Clear (src/app.cpp:144-164).
// Nothing in splatapult_b200/ links or loads this file; tests use it to pin the oracle (tests/test_ref_pin.py).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "camerasconfig.h"  // the reference's header (src/camerasconfig.h)
#include "gaussiancloud.h"  // the reference's header (src/gaussiancloud.h)
#include "glsl_shim.h"

#include "gen_shaders.inc"  // namespaces ref_presort, ref_vs_{base,full,base_srgb,full_srgb}, ref_gs, ref_fs

#define REF_API extern "C" __attribute__((visibility("default")))

struct ref_record
{
    float px, py, i00, i01, i10, i11, r, g, b, alpha;
    uint32_t key, visible;
};

REF_API int ref_import_ply(const char* path, int importFullSH, size_t* n, size_t* stride, int* fullSH, float* out, size_t capFloats)
{
    GaussianCloud::Options opt{importFullSH != 0, true};
    GaussianCloud gc(opt);
    if (!gc.ImportPly(path)) return -1;
    *n = gc.GetNumGaussians();
    *stride = gc.GetStride();
    *fullSH = gc.HasFullSH() ? 1 : 0;
    if (!out) return 0;
    if (capFloats * 4 < gc.GetTotalSize()) return -2;
    std::memcpy(out, gc.GetRawDataPtr(), gc.GetTotalSize());
    return 0;
}

// src/core/util.cpp is full of GL calls and is not compiled; CamerasConfig::EstimateFloorPlane needs this one function of
// it, restated from src/core/util.cpp:128-139 (normalize, or the fallback for a zero vector)
glm::vec3 SafeNormalize(const glm::vec3& v, const glm::vec3& ifZero)
{
    const float len = glm::length(v);
    return len > 0.0f ? glm::normalize(v) : ifZero;
}

// 1: camerasconfig.cpp was compiled against the real nlohmann/json found in the image, 0: against oracle/ref_shim/nlohmann
REF_API int ref_json_is_real()
{
#ifdef REF_REAL_NLOHMANN
    return 1;
#else
    return 0;
#endif
}

// CamerasConfig::ImportJson (src/camerasconfig.cpp:20-67) + EstimateFloorPlane (:69-96), executed by the reference's source.
// mats: 16 floats per camera, column-major (glm layout); fovs: 2 per camera; floor: normal[3] then position[3]
REF_API int ref_cameras_import(const char* path, float* mats, float* fovs, size_t capCameras, float* floor6)
{
    CamerasConfig cc;
    if (!cc.ImportJson(path)) return -1;
    const auto& v = cc.GetCameraVec();
    if (mats && fovs)
    {
        if (v.size() > capCameras) return -2;
        for (size_t i = 0; i < v.size(); i++)
        {
            for (int c = 0; c < 4; c++)
            {
                mats[i * 16 + c * 4 + 0] = v[i].mat[c].x; mats[i * 16 + c * 4 + 1] = v[i].mat[c].y;
                mats[i * 16 + c * 4 + 2] = v[i].mat[c].z; mats[i * 16 + c * 4 + 3] = v[i].mat[c].w;
            }
            fovs[i * 2 + 0] = v[i].fov.x;
            fovs[i * 2 + 1] = v[i].fov.y;
        }
    }
    if (floor6)
    {
        glm::vec3 nrm, pos;
        cc.EstimateFloorPlane(nrm, pos);
        floor6[0] = nrm.x; floor6[1] = nrm.y; floor6[2] = nrm.z; floor6[3] = pos.x; floor6[4] = pos.y; floor6[5] = pos.z;
    }
    return (int)v.size();
}

// what a caller sees after ImportJson, successful or not: the return value and how many cameras the object then holds (the
// reference emplaces per entry, so the entries before a bad one stay)
REF_API int ref_cameras_import_state(const char* path, int* ok)
{
    CamerasConfig cc;
    *ok = cc.ImportJson(path) ? 1 : 0;
    return (int)cc.GetNumCameras();
}

// the sixteen BinaryAttributes of GaussianCloud::InitAttribs (src/gaussiancloud.cpp:628-657) after importing `path`:
// out[16][3] = (type, size, offset) in the order of the accessors of src/gaussiancloud.h:42-57
REF_API int ref_attrib_table(const char* path, int importFullSH, uint64_t* out)
{
    GaussianCloud::Options opt{importFullSH != 0, true};
    GaussianCloud gc(opt);
    if (!gc.ImportPly(path)) return -1;
    const BinaryAttribute* a[16] = {&gc.GetPosWithAlphaAttrib(), &gc.GetR_SH0Attrib(), &gc.GetR_SH1Attrib(), &gc.GetR_SH2Attrib(), &gc.GetR_SH3Attrib(),
                                    &gc.GetG_SH0Attrib(), &gc.GetG_SH1Attrib(), &gc.GetG_SH2Attrib(), &gc.GetG_SH3Attrib(),
                                    &gc.GetB_SH0Attrib(), &gc.GetB_SH1Attrib(), &gc.GetB_SH2Attrib(), &gc.GetB_SH3Attrib(),
                                    &gc.GetCov3_Col0Attrib(), &gc.GetCov3_Col1Attrib(), &gc.GetCov3_Col2Attrib()};
    for (int i = 0; i < 16; i++)
    {
        out[i * 3 + 0] = (uint64_t)a[i]->type;
        out[i * 3 + 1] = (uint64_t)a[i]->size;
        out[i * 3 + 2] = (uint64_t)a[i]->offset;
    }
    return (int)gc.GetStride();
}

// GaussianCloud::PruneSplats (src/gaussiancloud.cpp:581-626) on an imported file: the kept records, in the reference's order
REF_API int ref_prune_splats(const char* path, int importFullSH, const float* origin, uint32_t keep, size_t* n, size_t* stride, float* out, size_t capFloats)
{
    GaussianCloud::Options opt{importFullSH != 0, true};
    GaussianCloud gc(opt);
    if (!gc.ImportPly(path)) return -1;
    gc.PruneSplats(glm::vec3(origin[0], origin[1], origin[2]), keep);
    *n = gc.GetNumGaussians();
    *stride = gc.GetStride();
    if (!out) return 0;
    if (capFloats * 4 < gc.GetTotalSize()) return -2;
    std::memcpy(out, gc.GetRawDataPtr(), gc.GetTotalSize());
    return 0;
}

// GaussianCloud::ExportPly (src/gaussiancloud.cpp:367-503) of an imported file, written by the reference's own Ply::Dump
REF_API int ref_export_ply(const char* inPath, int importFullSH, int exportFullSH, const char* outPath)
{
    GaussianCloud::Options opt{importFullSH != 0, exportFullSH != 0};
    GaussianCloud gc(opt);
    if (!gc.ImportPly(inPath)) return -1;
    return gc.ExportPly(outPath) ? 0 : -2;
}

REF_API int ref_init_debug_cloud(float* out, size_t capFloats)
{
    GaussianCloud::Options opt{true, true};
    GaussianCloud gc(opt);
    gc.InitDebugCloud();
    if (capFloats * 4 < gc.GetTotalSize()) return -2;
    std::memcpy(out, gc.GetRawDataPtr(), gc.GetTotalSize());
    return (int)gc.GetNumGaussians();
}

namespace
{
glsl::mat4 to_glsl(const glm::mat4& m)
{
    glsl::mat4 r;
    for (int c = 0; c < 4; c++) r[c] = glsl::vec4(m.at(c, 0), m.at(c, 1), m.at(c, 2), m.at(c, 3));
    return r;
}
glm::mat4 load(const float* p)
{
    glm::mat4 m;
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) m.at(c, r) = p[c * 4 + r];
    return m;
}

struct Quad
{
    glsl::vec4 pos[4];
    glsl::vec4 color, cov2inv;
    glsl::vec2 p;
    bool emitted;
};

// set the vertex inputs from one AoS record (attribute table: src/gaussiancloud.cpp:633-657, src/splatrenderer.cpp:368-387)
template <typename SetFull> void set_common(const float* g, glsl::vec4& position, glsl::vec4& r0, glsl::vec4& g0, glsl::vec4& b0, glsl::vec3& c0,
                                            glsl::vec3& c1, glsl::vec3& c2, SetFull setFull)
{
    position = glsl::vec4(g[0], g[1], g[2], g[3]);
    r0 = glsl::vec4(g[4], g[5], g[6], g[7]);
    g0 = glsl::vec4(g[8], g[9], g[10], g[11]);
    b0 = glsl::vec4(g[12], g[13], g[14], g[15]);
    c0 = glsl::vec3(g[16], g[17], g[18]);
    c1 = glsl::vec3(g[19], g[20], g[21]);
    c2 = glsl::vec3(g[22], g[23], g[24]);
    setFull();
}

#define RUN_VS(NS, FULL)                                                                                                     \
    {                                                                                                                        \
        NS::viewMat = viewG; NS::projMat = projG; NS::projParams = projParams; NS::viewport = viewportV; NS::eye = eyeV;      \
        set_common(g, NS::position, NS::r_sh0, NS::g_sh0, NS::b_sh0, NS::cov3_col0, NS::cov3_col1, NS::cov3_col2, [&]() { FULL });   \
        NS::shader_main();                                                                                                   \
        vsColor = NS::geom_color; vsCov2 = NS::geom_cov2; vsP = NS::geom_p; vsPos = NS::gl_Position;                          \
    }
#define SET_FULL(NS)                                                                                                           \
    NS::r_sh1 = glsl::vec4(g[25], g[26], g[27], g[28]); NS::r_sh2 = glsl::vec4(g[29], g[30], g[31], g[32]);                   \
    NS::r_sh3 = glsl::vec4(g[33], g[34], g[35], g[36]); NS::g_sh1 = glsl::vec4(g[37], g[38], g[39], g[40]);                   \
    NS::g_sh2 = glsl::vec4(g[41], g[42], g[43], g[44]); NS::g_sh3 = glsl::vec4(g[45], g[46], g[47], g[48]);                   \
    NS::b_sh1 = glsl::vec4(g[49], g[50], g[51], g[52]); NS::b_sh2 = glsl::vec4(g[53], g[54], g[55], g[56]);                   \
    NS::b_sh3 = glsl::vec4(g[57], g[58], g[59], g[60]);
}  // namespace

// One frame through the reference's shader source.  fb: H x W x 4 floats, row 0 = bottom (GL window coordinates).
REF_API size_t ref_render(const float* aos, size_t n, size_t strideBytes, int fullSH, int srgb, const float* cameraMat, const float* projMatIn,
                          const float* viewport, const float* nearFar, float* fb, ref_record* rec, uint32_t* orderOut)
{
    const size_t sf = strideBytes / 4;
    const int W = (int)viewport[2], H = (int)viewport[3];

    // ---- SplatRenderer::Sort host side, src/splatrenderer.cpp:160-177
    const glm::mat4 cameraM = load(cameraMat), projM = load(projMatIn);
    const glm::mat4 modelViewMat = glm::inverse(cameraM);
    const glm::mat4 mvp = projM * modelViewMat;
    ref_presort::modelViewProj = to_glsl(mvp);
    ref_presort::nearFar = glsl::vec2(nearFar[0], nearFar[1]);
    ref_presort::keyMax = 0xFFFFFFFFu;
    ref_presort::output_count.value = 0;
    ref_presort::positions.v.resize(n);
    for (size_t i = 0; i < n; i++) ref_presort::positions.v[i] = glsl::vec4(aos[i * sf], aos[i * sf + 1], aos[i * sf + 2], 1.0f);  // :106-111
    ref_presort::quantizedZs.v.assign(n, 0u);
    ref_presort::indices.v.assign(n, 0u);
    for (size_t i = 0; i < n; i++)
    {
        ref_presort::gl_GlobalInvocationID.x = (glsl::uint)i;  // dispatch :188-190 (atomic append visited in index order)
        ref_presort::shader_main();
    }
    const size_t sortCount = ref_presort::output_count.value;
    std::vector<std::pair<uint32_t, uint32_t>> kv(sortCount);
    for (size_t k = 0; k < sortCount; k++) kv[k] = {ref_presort::quantizedZs.v[k], ref_presort::indices.v[k]};
    // uint() of a value >= 2^32 is undefined (presort_compute.glsl:53): such gaussians (depth >= far) are dropped, like the oracle does
    std::vector<uint8_t> farCulled(n, 0);
    for (size_t i = 0; i < n; i++)
    {
        const float wclip = mvp.at(0, 3) * aos[i * sf] + mvp.at(1, 3) * aos[i * sf + 1] + mvp.at(2, 3) * aos[i * sf + 2] + mvp.at(3, 3);
        if (!(wclip < nearFar[1])) farCulled[i] = 1;
    }
    std::stable_sort(kv.begin(), kv.end(), [](const auto& a, const auto& b) { return a.first < b.first; });  // :206-264 contract

    // ---- SplatRenderer::Render host side, src/splatrenderer.cpp:324-335
    const glm::mat4 viewMatM = glm::inverse(cameraM);
    const glsl::mat4 viewG = to_glsl(viewMatM), projG = to_glsl(projM);
    const glsl::vec4 projParams(0.0f, nearFar[0], nearFar[1], 0.0f);
    const glsl::vec4 viewportV(viewport[0], viewport[1], viewport[2], viewport[3]);
    const glsl::vec3 eyeV(cameraM.at(3, 0), cameraM.at(3, 1), cameraM.at(3, 2));
    ref_gs::viewport = viewportV;

    for (size_t i = 0; i < n; i++) { std::memset(&rec[i], 0, sizeof(ref_record)); rec[i].alpha = aos[i * sf + 3]; }
    for (size_t k = 0; k < sortCount; k++) rec[kv[k].second].key = kv[k].first;

    // clear (0,0,0,1): src/app.cpp:158-160
    for (size_t p = 0; p < (size_t)W * H; p++) { fb[4 * p] = 0.f; fb[4 * p + 1] = 0.f; fb[4 * p + 2] = 0.f; fb[4 * p + 3] = 1.f; }

    size_t nOrder = 0;
    for (size_t k = 0; k < sortCount; k++)  // glDrawElements(GL_POINTS, sortCount) in sorted order, :338
    {
        const uint32_t id = kv[k].second;
        if (farCulled[id]) continue;
        const float* g = aos + (size_t)id * sf;
        glsl::vec4 vsColor, vsCov2, vsPos;
        glsl::vec2 vsP;
        if (fullSH)
        {
            if (srgb) RUN_VS(ref_vs_full_srgb, SET_FULL(ref_vs_full_srgb))
            else RUN_VS(ref_vs_full, SET_FULL(ref_vs_full))
        }
        else
        {
            if (srgb) RUN_VS(ref_vs_base_srgb, )
            else RUN_VS(ref_vs_base, )
        }
        // geometry stage
        ref_gs::geom_color[0] = vsColor; ref_gs::geom_cov2[0] = vsCov2; ref_gs::geom_p[0] = vsP;
        ref_gs::gl_in[0].gl_Position = vsPos;
        ref_gs::numEmitted = 0;
        ref_gs::shader_main();
        if (ref_gs::numEmitted != 4) continue;  // discarded by the guard band
        const glsl::vec4 fcol = ref_gs::emitted[0].color, fcov = ref_gs::emitted[0].cov2inv;
        const glsl::vec2 fp = ref_gs::emitted[0].p;
        // degenerate covariance (det <= 0 or non-finite): the reference would rasterise inf/NaN; skipped like the oracle
        const float m00 = vsCov2.x, m01 = vsCov2.y, m10 = vsCov2.z, m11 = vsCov2.w;
        const float det = m00 * m11 - m01 * m10;
        if (!(det > 0.0f) || !(det <= 3.4028234e38f)) continue;
        // whole quad has one clip z,w: clipped if outside the near/far planes
        const float ndcz = vsPos.z / vsPos.w;
        if (ndcz > 1.0f || ndcz < -1.0f) continue;

        ref_record& r = rec[id];
        r.px = fp.x; r.py = fp.y;
        r.i00 = fcov.x; r.i01 = fcov.y; r.i10 = fcov.z; r.i11 = fcov.w;
        r.r = fcol.x; r.g = fcol.y; r.b = fcol.z; r.alpha = fcol.w;
        r.visible = 1;
        if (orderOut) orderOut[nOrder] = id;
        nOrder++;

        // ---- rasterise the triangle strip (0,1,2),(1,2,3): pixel centres inside either triangle
        float wx[4], wy[4];
        for (int v = 0; v < 4; v++)
        {
            const glsl::vec4 c = ref_gs::emitted[v].pos;
            wx[v] = (c.x / c.w * 0.5f + 0.5f) * (float)W;
            wy[v] = (c.y / c.w * 0.5f + 0.5f) * (float)H;
        }
        const float minx = std::min(std::min(wx[0], wx[1]), std::min(wx[2], wx[3])), maxx = std::max(std::max(wx[0], wx[1]), std::max(wx[2], wx[3]));
        const float miny = std::min(std::min(wy[0], wy[1]), std::min(wy[2], wy[3])), maxy = std::max(std::max(wy[0], wy[1]), std::max(wy[2], wy[3]));
        if (!(maxx >= 0.f) || !(maxy >= 0.f) || !(minx <= (float)W) || !(miny <= (float)H)) continue;
        const int x0 = std::max(0, (int)std::floor(minx)), x1 = std::min(W - 1, (int)std::ceil(maxx));
        const int y0 = std::max(0, (int)std::floor(miny)), y1 = std::min(H - 1, (int)std::ceil(maxy));
        auto edge = [](float ax, float ay, float bx, float by, float px, float py) { return (double)(bx - ax) * (py - ay) - (double)(by - ay) * (px - ax); };
        auto inTri = [&](int a, int b, int c, float px, float py) {
            const double e0 = edge(wx[a], wy[a], wx[b], wy[b], px, py), e1 = edge(wx[b], wy[b], wx[c], wy[c], px, py), e2 = edge(wx[c], wy[c], wx[a], wy[a], px, py);
            return (e0 >= 0 && e1 >= 0 && e2 >= 0) || (e0 <= 0 && e1 <= 0 && e2 <= 0);
        };
        ref_fs::frag_color = fcol; ref_fs::frag_cov2inv = fcov; ref_fs::frag_p = fp;
        for (int y = y0; y <= y1; y++)
            for (int x = x0; x <= x1; x++)
            {
                const float px = (float)x + 0.5f, py = (float)y + 0.5f;
                if (!inTri(0, 1, 2, px, py) && !inTri(1, 2, 3, px, py)) continue;
                ref_fs::gl_FragCoord = glsl::vec4(px, py, 0.0f, 1.0f);
                ref_fs::gl_discarded = false;
                ref_fs::shader_main();
                if (ref_fs::gl_discarded) continue;
                const glsl::vec4 src = ref_fs::out_color;
                float* d = fb + 4 * ((size_t)y * W + x);
                const float om = 1.0f - src.a;  // glBlendFunc(GL_ONE, GL_ONE_MINUS_SRC_ALPHA), src/app.cpp:153-156
                d[0] = src.r + om * d[0];
                d[1] = src.g + om * d[1];
                d[2] = src.b + om * d[2];
                d[3] = src.a + om * d[3];
            }
    }
    return nOrder;
}
