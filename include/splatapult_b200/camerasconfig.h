// camerasconfig.h -- INRIA cameras.json reader.  Drop-in for the reference's CamerasConfig
// (src/camerasconfig.h:12-32, src/camerasconfig.cpp:20-95) without nlohmann-json / glm.
#pragma once
#include <cstddef>
#include <string>
#include <vector>

#include "splatmath.h"

struct Camera
{
    splat::mat4 mat;  // inverse view matrix (camera-to-world), the cameraMat SplatRenderer::Sort/Render take
    splat::vec2 fov;
};

class SPLAT_CPP_API CamerasConfig
{
public:
    CamerasConfig() = default;

    // false on a missing file, malformed JSON or a missing / non-numeric field -- the decisions nlohmann/json makes for the
    // reference (tests/test_ref_pin.py); entries read before a bad one stay in the list, as the reference's emplace_back leaves them
    bool ImportJson(const std::string& jsonFilename);

    const std::vector<Camera>& GetCameraVec() const { return cameraVec; }
    size_t GetNumCameras() const { return cameraVec.size(); }

    void EstimateFloorPlane(splat::vec3& normalOut, splat::vec3& posOut) const;
    const std::string& GetError() const { return error; }

protected:
    std::vector<Camera> cameraVec;
    std::string error;
};
