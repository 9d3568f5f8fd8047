// pack.cu -- Init-time repack of the reference's interleaved cloud into coalescing-friendly planes.
//
// The reference uploads the AoS verbatim as a vertex buffer and fetches 16 attributes per vertex through
// an index indirection (src/splatrenderer.cpp:345-391).  Here the AoS (100 B or 244 B records,
// src/gaussiancloud.cpp:32-56) is read once and split into float4 / float planes so that the projection
// kernel's one-thread-per-gaussian loads are fully coalesced.  Values are copied bit for bit.
#include "device_common.cuh"
#include "kernels.h"

namespace splat
{

// any non-zero degree-1 coefficient (r_sh0.yzw, g_sh0.yzw, b_sh0.yzw)?  ImportPly writes exact zeros there
// when the PLY has no f_rest (src/gaussiancloud.cpp:316-332) -> DC-only fast layout.
__global__ void k_detect_degree1(const float* __restrict__ aos, uint32_t n, uint32_t strideFloats, unsigned int* flag)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool nz = false;
    if (i < n)
    {
        const float* g = aos + (size_t)i * strideFloats;
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int k = 1; k < 4; k++) nz |= (__float_as_uint(g[4 + 4 * c + k]) & 0x7fffffffu) != 0u;
    }
    if (__any_sync(0xffffffffu, nz) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}

struct PackTargets
{
    float4* posAlpha;
    float4* covA;
    float4* covB;
    float* covC;
    float* dc[3];
    float4* sh[12];
};

__global__ void k_pack_cloud(const float* __restrict__ aos, uint32_t n, uint32_t strideFloats, int shMode, PackTargets t)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* g = aos + (size_t)i * strideFloats;
    t.posAlpha[i] = make_float4(g[0], g[1], g[2], g[3]);
    t.covA[i] = make_float4(g[16], g[17], g[18], g[19]);
    t.covB[i] = make_float4(g[20], g[21], g[22], g[23]);
    t.covC[i] = g[24];
    if (shMode == 0)
    {
        t.dc[0][i] = g[4];
        t.dc[1][i] = g[8];
        t.dc[2][i] = g[12];
    }
    else if (shMode == 1)
    {
#pragma unroll
        for (int c = 0; c < 3; c++) t.sh[c][i] = make_float4(g[4 + 4 * c], g[5 + 4 * c], g[6 + 4 * c], g[7 + 4 * c]);
    }
    else
    {
        // plane order r_sh0..3, g_sh0..3, b_sh0..3 ; AoS float offsets: sh0 at 4/8/12, sh1..3 at 25+12c+4(j-1)
#pragma unroll
        for (int c = 0; c < 3; c++)
        {
            t.sh[4 * c][i] = make_float4(g[4 + 4 * c], g[5 + 4 * c], g[6 + 4 * c], g[7 + 4 * c]);
#pragma unroll
            for (int j = 1; j < 4; j++)
            {
                const float* s = g + 25 + 12 * c + 4 * (j - 1);
                t.sh[4 * c + j][i] = make_float4(s[0], s[1], s[2], s[3]);
            }
        }
    }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

size_t packed_cloud_bytes(uint32_t n, int shMode)
{
    const size_t a = 256;
    size_t bytes = 3 * align_up((size_t)n * 16, a) + align_up((size_t)n * 4, a);
    if (shMode == 0) bytes += 3 * align_up((size_t)n * 4, a);
    else bytes += (size_t)(shMode == 1 ? 3 : 12) * align_up((size_t)n * 16, a);
    return bytes + a;
}

cudaError_t detect_sh_mode(const void* deviceAos, uint32_t n, uint32_t strideBytes, bool hasFullSH, unsigned int* dFlagScratch,
                           cudaStream_t stream, int* shModeOut)
{
    if (hasFullSH && strideBytes == 244) { *shModeOut = 3; return cudaSuccess; }
    cudaError_t e = cudaMemsetAsync(dFlagScratch, 0, sizeof(unsigned int), stream);
    if (e != cudaSuccess) return e;
    if (n) k_detect_degree1<<<(n + 255) / 256, 256, 0, stream>>>((const float*)deviceAos, n, strideBytes / 4, dFlagScratch);
    unsigned int flag = 0;
    e = cudaMemcpyAsync(&flag, dFlagScratch, sizeof flag, cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return e;
    *shModeOut = flag ? 1 : 0;
    return cudaSuccess;
}

cudaError_t pack_cloud(const void* deviceAos, uint32_t n, uint32_t strideBytes, int shMode, void* planes, PackedCloud* out, cudaStream_t stream)
{
    const size_t a = 256;
    uint8_t* p = (uint8_t*)planes;
    auto take = [&](size_t bytes) { void* r = p; p += align_up(bytes, a); return r; };
    PackTargets t{};
    t.posAlpha = (float4*)take((size_t)n * 16);
    t.covA = (float4*)take((size_t)n * 16);
    t.covB = (float4*)take((size_t)n * 16);
    t.covC = (float*)take((size_t)n * 4);
    if (shMode == 0)
        for (int c = 0; c < 3; c++) t.dc[c] = (float*)take((size_t)n * 4);
    else
        for (int k = 0; k < (shMode == 1 ? 3 : 12); k++) t.sh[k] = (float4*)take((size_t)n * 16);
    if (n) k_pack_cloud<<<(n + 255) / 256, 256, 0, stream>>>((const float*)deviceAos, n, strideBytes / 4, shMode, t);
    out->posAlpha = t.posAlpha; out->covA = t.covA; out->covB = t.covB; out->covC = t.covC;
    for (int c = 0; c < 3; c++) out->dc[c] = t.dc[c];
    for (int k = 0; k < 12; k++) out->sh[k] = t.sh[k];
    out->n = n;
    out->shMode = shMode;
    return cudaGetLastError();
}

}  // namespace splat
