// plydecode.h -- host-callable entry of the GPU PLY decode (plydecode.cu); no CUDA types, so the host-only
// translation units (gaussiancloud.cpp) can use it.
#pragma once
#include <cstddef>
#include <string>

class Ply;

namespace splat
{

// byte offset of each float property inside one vertex row; -1 = missing or not a float (decodes as 0, like the
// host reader).  Names as in the INRIA schema: x y z, opacity, f_dc_0..2, f_rest_0..44, scale_0..2, rot_0..3.
struct PlyDecodeLayout
{
    int pos[3];
    int opacity;
    int dc[3];
    int rest[45];
    int scale[3];
    int rot[4];
    int aligned;  // every offset and the vertex size are multiples of 4: rows can be read with 32-bit loads
};

// Copies the raw vertex rows to `device`, decodes them there into the reference's AoS (stride 100 bytes, or 244 when
// fullSH) and returns it in host memory (hostAosOut, numVertices * stride bytes) and / or leaves it on the device
// (*deviceAosOut, release with free_decoded_aos).  Blocks until done.  false + *err on failure; never falls back to the CPU.
bool decode_ply_on_device(const unsigned char* hostBody, size_t numVertices, size_t vertexSize, const PlyDecodeLayout& layout, bool fullSH, int device,
                          void* cudaStream, float* hostAosOut, void** deviceAosOut, std::string* err);
void free_decoded_aos(void* deviceAos);

// property lookup shared by GaussianCloud::ImportPlyGpu and SplatRenderer::InitFromPly (defined in gaussiancloud.cpp):
// *fullSH = wantFullSH and all 45 f_rest_* properties are present
void make_ply_decode_layout(const Ply& ply, bool wantFullSH, PlyDecodeLayout* out, bool* fullSH);

}  // namespace splat
