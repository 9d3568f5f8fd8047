// plydecode.cu -- GPU decode of an INRIA gaussian-splat PLY body into the reference's AoS.
//
// Stands in for the per-vertex loop of GaussianCloud::ImportPly (reference src/gaussiancloud.cpp:254-361; here
// csrc/gaussiancloud.cpp:157-201): alpha = sigmoid(opacity), scale = exp(scale_i), Sigma = R S S^T R^T with the
// normalised quaternion, SH repack.  The header is parsed on the host (csrc/ply.cpp), the raw body is copied to the
// device and one thread decodes one vertex straight into the 100 / 244-byte records the packer (pack.cu) reads.
//
// Bit parity with the host decode is kept, not approximated:
//   * float arithmetic in the host's operation order; this TU is built -fmad=false (the host TU is -ffp-contract=off),
//     divisions and square roots are IEEE;
//   * expf is the host's libm algorithm restated (glibc 2.27+ / ARM optimized-routines expf: k = round(x 32/ln2) by the
//     1.5*2^52 shift, r = x 32/ln2 - k, 2^(k/32) from a 32-entry table, degree-3 polynomial in double, one rounding to
//     float).  Doubles are IEEE on both sides, so the results are the host's: checked exhaustively on the CPU against
//     glibc 2.39 expf for every float with |x| <= 40 (2.2e9 values): one single input differs (x = 0x1.04845ep+5, by
//     1 ulp).  Outside |x| < 80 (overflow / underflow / NaN handling of libm) CUDA's expf is used.
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "plydecode.h"

namespace splat
{

namespace
{

__constant__ unsigned long long kExp2Tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
};

__device__ __forceinline__ float expf_libm(float x)
{
    if (!(fabsf(x) < 80.0f)) return expf(x);
    const double kInvLn2N = 0x1.71547652b82fep+5;  // 32 / ln 2
    const double kShift = 0x1.8p+52;
    const double c0 = 0x1.c6af84b912394p-20, c1 = 0x1.ebfce50fac4f3p-13, c2 = 0x1.62e42ff0c52d6p-6;
    double z = __dmul_rn(kInvLn2N, (double)x);
    double kd = __dadd_rn(z, kShift);
    const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
    kd = __dadd_rn(kd, -kShift);
    const double r = __dadd_rn(z, -kd);
    const unsigned long long t = kExp2Tab[ki & 31ull] + (ki << 47);
    const double s = __longlong_as_double((long long)t);
    z = __dadd_rn(__dmul_rn(c0, r), c1);
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(c2, r), 1.0);
    y = __dadd_rn(__dmul_rn(z, r2), y);
    y = __dmul_rn(y, s);
    return (float)y;
}

struct Mat3
{
    float m[9];  // column-major, at(c, r) = m[c * 3 + r]   (include/splatapult_b200/splatmath.h)
    __device__ float& at(int c, int r) { return m[c * 3 + r]; }
    __device__ float at(int c, int r) const { return m[c * 3 + r]; }
};
__device__ __forceinline__ Mat3 mul3(const Mat3& a, const Mat3& b)
{
    Mat3 r;
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int row = 0; row < 3; row++) r.at(c, row) = (a.at(0, row) * b.at(c, 0) + a.at(1, row) * b.at(c, 1)) + a.at(2, row) * b.at(c, 2);
    return r;
}
__device__ __forceinline__ Mat3 transpose3(const Mat3& a)
{
    Mat3 r;
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int row = 0; row < 3; row++) r.at(c, row) = a.at(row, c);
    return r;
}
// splatmath.h rotationFromQuat
__device__ __forceinline__ Mat3 rotation_from_quat(float w, float x, float y, float z)
{
    const float len = sqrtf((w * w + x * x) + (y * y + z * z));
    if (len <= 0.0f) { w = 1.0f; x = y = z = 0.0f; }
    else { const float s = 1.0f / len; w = w * s; x = x * s; y = y * s; z = z * s; }
    const float xx = x * x, yy = y * y, zz = z * z, xz = x * z, xy = x * y, yz = y * z;
    const float wx = w * x, wy = w * y, wz = w * z;
    Mat3 r;
    r.at(0, 0) = 1.0f - 2.0f * (yy + zz); r.at(0, 1) = 2.0f * (xy + wz);        r.at(0, 2) = 2.0f * (xz - wy);
    r.at(1, 0) = 2.0f * (xy - wz);        r.at(1, 1) = 1.0f - 2.0f * (xx + zz); r.at(1, 2) = 2.0f * (yz + wx);
    r.at(2, 0) = 2.0f * (xz + wy);        r.at(2, 1) = 2.0f * (yz - wx);        r.at(2, 2) = 1.0f - 2.0f * (xx + yy);
    return r;
}

// float offsets inside one record (csrc/gaussiancloud.cpp; reference src/gaussiancloud.cpp:32-56)
constexpr int kPos = 0, kCov = 16;
__device__ __forceinline__ int sh_slot(int channel, int k)
{
    const int block[3][4] = {{4, 25, 29, 33}, {8, 37, 41, 45}, {12, 49, 53, 57}};
    return block[channel][k >> 2] + (k & 3);
}

__global__ void __launch_bounds__(128)
k_decode_ply(const unsigned char* __restrict__ body, size_t n, size_t vsize, const PlyDecodeLayout L, int fullSH, float* __restrict__ aos,
             int recFloats)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned char* v = body + i * vsize;
    // a property that is missing or not a float reads as 0 (host: readf)
    auto rd = [&](int off) -> float {
        if (off < 0) return 0.0f;
        if (L.aligned) return *reinterpret_cast<const float*>(v + off);
        float f;
        memcpy(&f, v + off, 4);  // odd layouts (char / short properties before a float): byte loads
        return f;
    };
    float* rec = aos + i * (size_t)recFloats;
    for (int k = 0; k < recFloats; k++) rec[k] = 0.0f;  // the host Allocate() zero-fills
    rec[kPos + 0] = rd(L.pos[0]);
    rec[kPos + 1] = rd(L.pos[1]);
    rec[kPos + 2] = rd(L.pos[2]);
    rec[kPos + 3] = 1.0f / (1.0f + expf_libm(-rd(L.opacity)));
    for (int c = 0; c < 3; c++)
    {
        rec[sh_slot(c, 0)] = rd(L.dc[c]);
        if (fullSH)
            for (int k = 1; k < 16; k++) rec[sh_slot(c, k)] = rd(L.rest[15 * c + k - 1]);
    }
    const float s0 = expf_libm(rd(L.scale[0])), s1 = expf_libm(rd(L.scale[1])), s2 = expf_libm(rd(L.scale[2]));
    const Mat3 R = rotation_from_quat(rd(L.rot[0]), rd(L.rot[1]), rd(L.rot[2]), rd(L.rot[3]));
    Mat3 S;
#pragma unroll
    for (int k = 0; k < 9; k++) S.m[k] = 0.0f;
    S.at(0, 0) = s0; S.at(1, 1) = s1; S.at(2, 2) = s2;
    const Mat3 V = mul3(mul3(mul3(R, S), transpose3(S)), transpose3(R));
#pragma unroll
    for (int k = 0; k < 9; k++) rec[kCov + k] = V.m[k];
}

std::string cuda_err(const char* what, cudaError_t e) { return std::string(what) + ": " + cudaGetErrorString(e); }

}  // namespace

bool decode_ply_on_device(const unsigned char* hostBody, size_t numVertices, size_t vertexSize, const PlyDecodeLayout& layout, bool fullSH, int device,
                          void* cudaStream, float* hostAosOut, void** deviceAosOut, std::string* err)
{
    auto fail = [&](const std::string& m) { if (err) *err = m; return false; };
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(cuda_err("cudaSetDevice", e) + " (GPU PLY decode needs a CUDA device; ImportPly is the host path)");
    cudaStream_t stream = static_cast<cudaStream_t>(cudaStream);
    const int recFloats = fullSH ? 61 : 25;
    // the renderer takes at most 2^31 - 1 gaussians; beyond that the byte counts below (and the 32-bit grid) would wrap
    if (numVertices >= (size_t(1) << 31) || (vertexSize && numVertices > static_cast<size_t>(-1) / vertexSize))
        return fail("vertex count " + std::to_string(numVertices) + " is out of range (at most 2^31 - 1 vertices)");
    const size_t bodyBytes = numVertices * vertexSize, aosBytes = numVertices * (size_t)recFloats * sizeof(float);
    unsigned char* dBody = nullptr;
    float* dAos = nullptr;
    if ((e = cudaMalloc(&dBody, bodyBytes ? bodyBytes : 16)) != cudaSuccess) return fail(cuda_err("cudaMalloc(ply body)", e));
    if ((e = cudaMalloc(&dAos, aosBytes ? aosBytes : 16)) != cudaSuccess)
    {
        cudaFree(dBody);
        return fail(cuda_err("cudaMalloc(aos)", e));
    }
    bool ok = true;
    std::string msg;
    if (numVertices)
    {
        if ((e = cudaMemcpyAsync(dBody, hostBody, bodyBytes, cudaMemcpyHostToDevice, stream)) != cudaSuccess) { ok = false; msg = cuda_err("H2D ply body", e); }
        if (ok)
        {
            const unsigned blocks = (unsigned)((numVertices + 127) / 128);
            k_decode_ply<<<blocks, 128, 0, stream>>>(dBody, numVertices, vertexSize, layout, fullSH ? 1 : 0, dAos, recFloats);
            if ((e = cudaGetLastError()) != cudaSuccess) { ok = false; msg = cuda_err("k_decode_ply", e); }
        }
        if (ok && hostAosOut && (e = cudaMemcpyAsync(hostAosOut, dAos, aosBytes, cudaMemcpyDeviceToHost, stream)) != cudaSuccess)
        {
            ok = false;
            msg = cuda_err("D2H aos", e);
        }
    }
    if (ok && (e = cudaStreamSynchronize(stream)) != cudaSuccess) { ok = false; msg = cuda_err("decode sync", e); }
    cudaFree(dBody);
    if (!ok || !deviceAosOut) cudaFree(dAos);
    else *deviceAosOut = dAos;
    if (!ok) return fail(msg);
    return true;
}

void free_decoded_aos(void* deviceAos) { cudaFree(deviceAos); }

}  // namespace splat
