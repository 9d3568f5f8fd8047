// gaussiancloud.h -- host-side gaussian cloud in the reference's interleaved layout.
// Drop-in for the reference's GaussianCloud (src/gaussiancloud.h:17-91): same constructor, methods,
// accessors and byte-identical record layouts (100 B base / 244 B full SH, src/gaussiancloud.cpp:32-56),
// without glm / Eigen.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <string>

#include "binaryattribute.h"
#include "splatmath.h"

class SPLAT_CPP_API GaussianCloud
{
public:
    struct Options
    {
        bool importFullSH;
        bool exportFullSH;
    };

    explicit GaussianCloud(const Options& options);

    bool ImportPly(const std::string& plyFilename);
    // same result (bit for bit, see csrc/plydecode.cu), with the per-vertex decode run as a CUDA kernel on `device`;
    // fails without a CUDA device -- ImportPly is the host path
    bool ImportPlyGpu(const std::string& plyFilename, int device = 0);
    bool ExportPly(const std::string& plyFilename) const;
    void InitDebugCloud();
    // only keep the `numGaussians` splats nearest to `origin`
    void PruneSplats(const splat::vec3& origin, uint32_t numGaussians);
    // adopt a copy of `count` records of `stride` bytes (100 or 244) already in the reference layout
    bool SetRawData(const void* aos, size_t count, size_t stride, bool fullSH);

    size_t GetNumGaussians() const { return numGaussians; }
    size_t GetStride() const { return gaussianSize; }
    size_t GetTotalSize() const { return GetNumGaussians() * gaussianSize; }
    void* GetRawDataPtr() { return data.get(); }
    const void* GetRawDataPtr() const { return data.get(); }

    const BinaryAttribute& GetPosWithAlphaAttrib() const { return posWithAlphaAttrib; }
    const BinaryAttribute& GetR_SH0Attrib() const { return r_sh0Attrib; }
    const BinaryAttribute& GetR_SH1Attrib() const { return r_sh1Attrib; }
    const BinaryAttribute& GetR_SH2Attrib() const { return r_sh2Attrib; }
    const BinaryAttribute& GetR_SH3Attrib() const { return r_sh3Attrib; }
    const BinaryAttribute& GetG_SH0Attrib() const { return g_sh0Attrib; }
    const BinaryAttribute& GetG_SH1Attrib() const { return g_sh1Attrib; }
    const BinaryAttribute& GetG_SH2Attrib() const { return g_sh2Attrib; }
    const BinaryAttribute& GetG_SH3Attrib() const { return g_sh3Attrib; }
    const BinaryAttribute& GetB_SH0Attrib() const { return b_sh0Attrib; }
    const BinaryAttribute& GetB_SH1Attrib() const { return b_sh1Attrib; }
    const BinaryAttribute& GetB_SH2Attrib() const { return b_sh2Attrib; }
    const BinaryAttribute& GetB_SH3Attrib() const { return b_sh3Attrib; }
    const BinaryAttribute& GetCov3_Col0Attrib() const { return cov3_col0Attrib; }
    const BinaryAttribute& GetCov3_Col1Attrib() const { return cov3_col1Attrib; }
    const BinaryAttribute& GetCov3_Col2Attrib() const { return cov3_col2Attrib; }

    using ForEachPosWithAlphaCallback = std::function<void(const float*)>;
    void ForEachPosWithAlpha(const ForEachPosWithAlphaCallback& cb) const;

    bool HasFullSH() const { return hasFullSH; }
    const std::string& GetError() const { return error; }

protected:
    void InitAttribs();
    void Allocate(size_t count, size_t stride);

    std::shared_ptr<void> data;

    BinaryAttribute posWithAlphaAttrib;
    BinaryAttribute r_sh0Attrib, r_sh1Attrib, r_sh2Attrib, r_sh3Attrib;
    BinaryAttribute g_sh0Attrib, g_sh1Attrib, g_sh2Attrib, g_sh3Attrib;
    BinaryAttribute b_sh0Attrib, b_sh1Attrib, b_sh2Attrib, b_sh3Attrib;
    BinaryAttribute cov3_col0Attrib, cov3_col1Attrib, cov3_col2Attrib;

    size_t numGaussians;
    size_t gaussianSize;

    Options opt;
    bool hasFullSH;
    mutable std::string error;
};
