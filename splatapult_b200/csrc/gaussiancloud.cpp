// gaussiancloud.cpp -- PLY <-> reference AoS conversion, debug cloud, nearest-K prune.
// Behavioural contract: src/gaussiancloud.cpp:32-657 of the reference.  Written without glm / Eigen.
//   decode  alpha = sigmoid(opacity) (:119-122,260), scale = exp(scale_i) (:335-340),
//           Sigma = R S S^T R^T with q = (rot_0=w, rot_1=x, rot_2=y, rot_3=z) normalised (:86-94),
//           SH: f_dc_c -> {r,g,b}_sh0[0], f_rest_{15c+k-1} -> {r,g,b}_sh{k>>2}[k&3] (:262-314)
//   export  inverse of the above; the reference indexes r_sh0[i+1] for i<15 straight across struct
//           fields (:463-474), which does not match the vec4-per-degree interleave -- fixed here.
// Compile with -ffp-contract=off (bit-reproducible against the oracle).
#include "splatapult_b200/gaussiancloud.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <new>
#include <numeric>
#include <vector>

#include "plydecode.h"
#include "splatapult_b200/ply.h"

namespace
{
// float offsets inside one record (src/gaussiancloud.cpp:32-56)
constexpr size_t kPos = 0, kRsh0 = 4, kGsh0 = 8, kBsh0 = 12, kCov = 16;
constexpr size_t kBaseFloats = 25, kFullFloats = 61;
// start (in floats) of {r,g,b}_sh{0..3}
constexpr size_t kShBlock[3][4] = {{4, 25, 29, 33}, {8, 37, 41, 45}, {12, 49, 53, 57}};

inline float* ShSlot(float* rec, int channel, int k) { return rec + kShBlock[channel][k >> 2] + (k & 3); }
inline const float* ShSlot(const float* rec, int channel, int k) { return rec + kShBlock[channel][k >> 2] + (k & 3); }

// R S S^T R^T, column-major 9 floats
void CovarianceFromRotScale(const float rot[4], const float scale[3], float out[9])
{
    using namespace splat;
    const mat3 R = rotationFromQuat(rot[0], rot[1], rot[2], rot[3]);
    mat3 S;
    std::memset(S.m, 0, sizeof S.m);
    S.at(0, 0) = scale[0]; S.at(1, 1) = scale[1]; S.at(2, 2) = scale[2];
    const mat3 V = mul(mul(mul(R, S), transpose(S)), transpose(R));
    std::memcpy(out, V.m, sizeof V.m);
}

// cyclic Jacobi eigen-decomposition of a symmetric 3x3 (double precision); eigenvalues ascending like
// Eigen::SelfAdjointEigenSolver, eigenvectors in the columns of vec.
void SymmetricEigen3(const float covColMajor[9], double val[3], double vec[3][3])
{
    double a[3][3];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) a[r][c] = 0.5 * ((double)covColMajor[c * 3 + r] + (double)covColMajor[r * 3 + c]);
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 64; sweep++)
    {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        if (off < 1e-300) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++)
            {
                if (std::fabs(a[p][q]) < 1e-300) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; k++)
                {
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; k++)
                {
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; k++)
                {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
    int idx[3] = {0, 1, 2};
    std::sort(idx, idx + 3, [&](int i, int j) { return a[i][i] < a[j][j]; });
    for (int j = 0; j < 3; j++)
    {
        val[j] = a[idx[j]][idx[j]];
        for (int k = 0; k < 3; k++) vec[k][j] = v[k][idx[j]];
    }
}

// rotation matrix (columns = eigenvectors, det forced to +1) -> unit quaternion (w,x,y,z)
void QuatFromRotation(double m[3][3], float q[4])
{
    const double det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                       m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
    if (det < 0)
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) m[r][c] = -m[r][c];
    double w, x, y, z;
    const double tr = m[0][0] + m[1][1] + m[2][2];
    if (tr > 0)
    {
        const double s = std::sqrt(tr + 1.0) * 2.0;
        w = 0.25 * s; x = (m[2][1] - m[1][2]) / s; y = (m[0][2] - m[2][0]) / s; z = (m[1][0] - m[0][1]) / s;
    }
    else if (m[0][0] > m[1][1] && m[0][0] > m[2][2])
    {
        const double s = std::sqrt(1.0 + m[0][0] - m[1][1] - m[2][2]) * 2.0;
        w = (m[2][1] - m[1][2]) / s; x = 0.25 * s; y = (m[0][1] + m[1][0]) / s; z = (m[0][2] + m[2][0]) / s;
    }
    else if (m[1][1] > m[2][2])
    {
        const double s = std::sqrt(1.0 + m[1][1] - m[0][0] - m[2][2]) * 2.0;
        w = (m[0][2] - m[2][0]) / s; x = (m[0][1] + m[1][0]) / s; y = 0.25 * s; z = (m[1][2] + m[2][1]) / s;
    }
    else
    {
        const double s = std::sqrt(1.0 + m[2][2] - m[0][0] - m[1][1]) * 2.0;
        w = (m[1][0] - m[0][1]) / s; x = (m[0][2] + m[2][0]) / s; y = (m[1][2] + m[2][1]) / s; z = 0.25 * s;
    }
    const double n = std::sqrt(w * w + x * x + y * y + z * z);
    q[0] = (float)(w / n); q[1] = (float)(x / n); q[2] = (float)(y / n); q[3] = (float)(z / n);
}
}  // namespace

GaussianCloud::GaussianCloud(const Options& options) : numGaussians(0), gaussianSize(0), opt(options), hasFullSH(false) {}

void GaussianCloud::Allocate(size_t count, size_t stride)
{
    // checked: a wrapped product would allocate one float and let the decode loop write `count` records past it
    if (stride && count > static_cast<size_t>(-1) / stride) throw std::bad_alloc();
    numGaussians = count;
    gaussianSize = stride;
    const size_t floats = count * stride / sizeof(float);
    float* p = new float[floats ? floats : 1]();
    data = std::shared_ptr<void>(p, [](void* q) { delete[] static_cast<float*>(q); });
}

void GaussianCloud::InitAttribs()
{
    using T = BinaryAttribute::Type;
    auto at = [](size_t floatOffset) { return BinaryAttribute(T::Float, floatOffset * sizeof(float)); };
    posWithAlphaAttrib = at(kPos);
    r_sh0Attrib = at(kShBlock[0][0]); g_sh0Attrib = at(kShBlock[1][0]); b_sh0Attrib = at(kShBlock[2][0]);
    cov3_col0Attrib = at(kCov); cov3_col1Attrib = at(kCov + 3); cov3_col2Attrib = at(kCov + 6);
    // the degree 2/3 blocks exist only for clouds imported with full SH (InitDebugCloud leaves them Unknown
    // although its stride is 244: GetStride() is authoritative)
    r_sh1Attrib = r_sh2Attrib = r_sh3Attrib = BinaryAttribute();
    g_sh1Attrib = g_sh2Attrib = g_sh3Attrib = BinaryAttribute();
    b_sh1Attrib = b_sh2Attrib = b_sh3Attrib = BinaryAttribute();
    if (hasFullSH)
    {
        r_sh1Attrib = at(kShBlock[0][1]); r_sh2Attrib = at(kShBlock[0][2]); r_sh3Attrib = at(kShBlock[0][3]);
        g_sh1Attrib = at(kShBlock[1][1]); g_sh2Attrib = at(kShBlock[1][2]); g_sh3Attrib = at(kShBlock[1][3]);
        b_sh1Attrib = at(kShBlock[2][1]); b_sh2Attrib = at(kShBlock[2][2]); b_sh3Attrib = at(kShBlock[2][3]);
    }
}

namespace
{
// name -> (type, offset) lookups of the INRIA schema; a missing required property only reads as zero
struct PlyProps
{
    BinaryAttribute px, py, pz, opacity, dc[3], rest[45], scale[3], rot[4];
    bool fullSH = false;
};
PlyProps LookupProps(const Ply& ply, bool wantFullSH)
{
    PlyProps p;
    // `!Get("x") || !Get("y") || !Get("z")` in the reference (gaussiancloud.cpp:174-176) short-circuits: after a missing x
    // neither y nor z is looked up, after a missing y not z -- they then read as zero like any missing property
    if (ply.GetProperty("x", p.px) && ply.GetProperty("y", p.py)) ply.GetProperty("z", p.pz);
    ply.GetProperty("opacity", p.opacity);
    for (int i = 0; i < 3; i++) ply.GetProperty("f_dc_" + std::to_string(i), p.dc[i]);
    for (int i = 0; i < 3; i++) ply.GetProperty("scale_" + std::to_string(i), p.scale[i]);
    for (int i = 0; i < 4; i++) ply.GetProperty("rot_" + std::to_string(i), p.rot[i]);
    p.fullSH = wantFullSH;
    for (int i = 0; p.fullSH && i < 45; i++)
        if (!ply.GetProperty("f_rest_" + std::to_string(i), p.rest[i])) p.fullSH = false;  // all 45 or none
    return p;
}
}  // namespace

bool GaussianCloud::ImportPly(const std::string& plyFilename)
{
    std::ifstream file(plyFilename, std::ios::binary);
    if (!file.is_open()) { error = "failed to open " + plyFilename; return false; }
    Ply ply;
    if (!ply.Parse(file)) { error = "error parsing ply file \"" + plyFilename + "\": " + ply.GetError(); return false; }

    const PlyProps P = LookupProps(ply, opt.importFullSH);
    hasFullSH = P.fullSH;
    auto readf = [](const BinaryAttribute& a, const void* v) { return a.type == BinaryAttribute::Type::Float ? a.Read<float>(v) : 0.0f; };

    InitAttribs();
    try
    {
        Allocate(ply.GetVertexCount(), (hasFullSH ? kFullFloats : kBaseFloats) * sizeof(float));
    }
    catch (const std::exception&)
    {
        error = "cannot allocate " + std::to_string(ply.GetVertexCount()) + " gaussians";
        numGaussians = 0;
        return false;
    }
    const size_t recFloats = gaussianSize / sizeof(float);
    float* rec = static_cast<float*>(data.get());
    const uint8_t* v = ply.GetData();
    const size_t vsize = ply.GetVertexSize();
    for (size_t n = 0; n < numGaussians; n++, rec += recFloats, v += vsize)
    {
        rec[kPos + 0] = readf(P.px, v);
        rec[kPos + 1] = readf(P.py, v);
        rec[kPos + 2] = readf(P.pz, v);
        rec[kPos + 3] = 1.0f / (1.0f + expf(-readf(P.opacity, v)));
        for (int c = 0; c < 3; c++)
        {
            *ShSlot(rec, c, 0) = readf(P.dc[c], v);
            if (hasFullSH)
                for (int k = 1; k < 16; k++) *ShSlot(rec, c, k) = readf(P.rest[15 * c + k - 1], v);
            // base layout: the three degree-1 slots stay 0 (Allocate zero-fills)
        }
        const float s[3] = {expf(readf(P.scale[0], v)), expf(readf(P.scale[1], v)), expf(readf(P.scale[2], v))};
        const float q[4] = {readf(P.rot[0], v), readf(P.rot[1], v), readf(P.rot[2], v), readf(P.rot[3], v)};
        CovarianceFromRotScale(q, s, rec + kCov);
    }
    return true;
}

// Same result as ImportPly, with the per-vertex decode done by a CUDA kernel (csrc/plydecode.cu): header parsed
// here, raw body copied to the device, the decoded AoS copied back.  SURVEY 8f-f4.  No CPU fallback: without a
// CUDA device this returns false.
bool GaussianCloud::ImportPlyGpu(const std::string& plyFilename, int device)
{
    std::ifstream file(plyFilename, std::ios::binary);
    if (!file.is_open()) { error = "failed to open " + plyFilename; return false; }
    Ply ply;
    if (!ply.Parse(file)) { error = "error parsing ply file \"" + plyFilename + "\": " + ply.GetError(); return false; }
    bool fullSH = false;
    splat::PlyDecodeLayout layout;
    splat::make_ply_decode_layout(ply, opt.importFullSH, &layout, &fullSH);
    hasFullSH = fullSH;
    InitAttribs();
    try
    {
        Allocate(ply.GetVertexCount(), (hasFullSH ? kFullFloats : kBaseFloats) * sizeof(float));
    }
    catch (const std::exception&)
    {
        error = "cannot allocate " + std::to_string(ply.GetVertexCount()) + " gaussians";
        numGaussians = 0;
        return false;
    }
    std::string err;
    if (!splat::decode_ply_on_device(ply.GetData(), numGaussians, ply.GetVertexSize(), layout, hasFullSH, device, nullptr,
                                     static_cast<float*>(data.get()), nullptr, &err))
    {
        error = "GPU ply decode of \"" + plyFilename + "\" failed: " + err;
        return false;
    }
    return true;
}

void splat::make_ply_decode_layout(const Ply& ply, bool wantFullSH, splat::PlyDecodeLayout* out, bool* fullSHOut)
{
    const PlyProps P = LookupProps(ply, wantFullSH);
    auto off = [](const BinaryAttribute& a) { return a.type == BinaryAttribute::Type::Float ? (int)a.offset : -1; };
    splat::PlyDecodeLayout& L = *out;
    L.pos[0] = off(P.px); L.pos[1] = off(P.py); L.pos[2] = off(P.pz);
    L.opacity = off(P.opacity);
    for (int i = 0; i < 3; i++) { L.dc[i] = off(P.dc[i]); L.scale[i] = off(P.scale[i]); }
    for (int i = 0; i < 4; i++) L.rot[i] = off(P.rot[i]);
    for (int i = 0; i < 45; i++) L.rest[i] = P.fullSH ? off(P.rest[i]) : -1;
    bool aligned = ply.GetVertexSize() % 4 == 0;
    const int* all = &L.pos[0];
    for (size_t i = 0; aligned && i < (sizeof(splat::PlyDecodeLayout) / sizeof(int)) - 1; i++)
        if (all[i] >= 0 && all[i] % 4 != 0) aligned = false;
    L.aligned = aligned ? 1 : 0;
    *fullSHOut = P.fullSH;
}

bool GaussianCloud::ExportPly(const std::string& plyFilename) const
{
    std::ofstream file(plyFilename, std::ios::binary);
    if (!file.is_open()) { error = "failed to open " + plyFilename; return false; }

    using T = BinaryAttribute::Type;
    Ply ply;
    std::vector<std::string> names = {"x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"};
    if (opt.exportFullSH)
        for (int i = 0; i < 45; i++) names.push_back("f_rest_" + std::to_string(i));
    for (const char* n : {"opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"}) names.push_back(n);
    for (const std::string& n : names) ply.AddProperty(n, T::Float);
    ply.AllocData(numGaussians);

    const bool srcHasRest = gaussianSize >= kFullFloats * sizeof(float);
    const size_t recFloats = gaussianSize / sizeof(float);
    const float* rec = static_cast<const float*>(data.get());
    float* out = reinterpret_cast<float*>(ply.GetData());
    const size_t outFloats = ply.GetVertexSize() / sizeof(float);
    for (size_t n = 0; n < numGaussians; n++, rec += recFloats, out += outFloats)
    {
        float* o = out;
        *o++ = rec[kPos]; *o++ = rec[kPos + 1]; *o++ = rec[kPos + 2];
        *o++ = 0.0f; *o++ = 0.0f; *o++ = 0.0f;  // normals
        for (int c = 0; c < 3; c++) *o++ = *ShSlot(rec, c, 0);
        if (opt.exportFullSH)
            for (int c = 0; c < 3; c++)
                for (int k = 1; k < 16; k++) *o++ = (k < 4 || srcHasRest) ? *ShSlot(rec, c, k) : 0.0f;
        const float alpha = rec[kPos + 3];
        *o++ = -logf((1.0f / alpha) - 1.0f);  // logit

        double val[3], vec[3][3];
        SymmetricEigen3(rec + kCov, val, vec);
        float q[4];
        QuatFromRotation(vec, q);
        for (int i = 0; i < 3; i++) *o++ = logf(sqrtf((float)std::max(val[i], 0.0)));
        for (int i = 0; i < 4; i++) *o++ = q[i];
    }
    ply.Dump(file);
    return true;
}

void GaussianCloud::InitDebugCloud()
{
    // red / green / blue axes (5 splats each) + white origin; stride 244 with hasFullSH == false
    constexpr int kPerAxis = 5;
    const float delta = 1.0f / (float)kPerAxis;
    const float covDiag = 0.005f;
    const float shC0 = 0.28209479177387814f;
    const float shOne = 1.0f / (2.0f * shC0), shZero = -1.0f / (2.0f * shC0);
    hasFullSH = false;
    InitAttribs();
    Allocate(3 * kPerAxis + 1, kFullFloats * sizeof(float));
    float* rec = static_cast<float*>(data.get());
    auto fill = [&](float* g, float x, float y, float z, float r, float gr, float b) {
        g[kPos] = x; g[kPos + 1] = y; g[kPos + 2] = z; g[kPos + 3] = 1.0f;
        g[kRsh0] = r; g[kGsh0] = gr; g[kBsh0] = b;
        g[kCov] = covDiag; g[kCov + 4] = covDiag; g[kCov + 8] = covDiag;
    };
    for (int axis = 0; axis < 3; axis++)
        for (int i = 0; i < kPerAxis; i++)
        {
            float p[3] = {0.0f, 0.0f, 0.0f};
            p[axis] = i * delta + delta + (axis == 2 ? 0.0001f : 0.0f);
            fill(rec + (axis * kPerAxis + i) * kFullFloats, p[0], p[1], p[2], axis == 0 ? shOne : shZero, axis == 1 ? shOne : shZero,
                 axis == 2 ? shOne : shZero);
        }
    fill(rec + 3 * kPerAxis * kFullFloats, 0.0f, 0.0f, 0.0f, shOne, shOne, shOne);
}

void GaussianCloud::PruneSplats(const splat::vec3& origin, uint32_t keep)
{
    if (!data || (size_t)keep >= numGaussians) return;
    const size_t recFloats = gaussianSize / sizeof(float);
    const float* src = static_cast<const float*>(data.get());
    std::vector<float> dist(numGaussians);
    for (size_t i = 0; i < numGaussians; i++)
    {
        const float* p = src + i * recFloats;
        const float dx = origin.x - p[0], dy = origin.y - p[1], dz = origin.z - p[2];
        dist[i] = sqrtf(dx * dx + dy * dy + dz * dz);
    }
    std::vector<uint32_t> order(numGaussians);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return dist[a] < dist[b]; });
    std::shared_ptr<void> old = data;
    const size_t stride = gaussianSize;
    Allocate(keep, stride);
    uint8_t* dst = static_cast<uint8_t*>(data.get());
    for (uint32_t i = 0; i < keep; i++) std::memcpy(dst + (size_t)i * stride, static_cast<const uint8_t*>(old.get()) + (size_t)order[i] * stride, stride);
}

bool GaussianCloud::SetRawData(const void* aos, size_t count, size_t stride, bool fullSH)
{
    if (stride != kBaseFloats * sizeof(float) && stride != kFullFloats * sizeof(float)) { error = "stride must be 100 or 244"; return false; }
    if (fullSH && stride != kFullFloats * sizeof(float)) { error = "full SH needs stride 244"; return false; }
    hasFullSH = fullSH;
    InitAttribs();
    Allocate(count, stride);
    if (count) std::memcpy(data.get(), aos, count * stride);
    return true;
}

void GaussianCloud::ForEachPosWithAlpha(const ForEachPosWithAlphaCallback& cb) const
{
    posWithAlphaAttrib.ForEach<float>(GetRawDataPtr(), GetStride(), GetNumGaussians(), cb);
}
