// camerasconfig.cpp -- cameras.json (INRIA gaussian-splatting output) -> camera-to-world matrices.
// Contract: src/camerasconfig.cpp:20-95 of the reference.  Per entry {id, position[3], rotation[3][3], width, height,
// fx, fy}: mat = columns (R col 0, -R col 1, -R col 2, position) -- the JSON rotation is camera-to-world with y down /
// z forward, the renderer wants GL's y up / -z forward (:50-53); fov = 2 atan(size / (2 fx)) for both axes (the
// reference uses fx for the vertical fov too; fov only feeds gizmos).  Any malformed or missing field
// makes ImportJson return false, like the nlohmann exception path (:58-63).
#include "splatapult_b200/camerasconfig.h"

#include <cmath>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>

namespace
{
// a deliberately small JSON value model: enough for cameras.json, strict enough to reject garbage
struct Json
{
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    double num = 0;
    bool b = false;
    std::string str;
    std::vector<Json> arr;
    std::map<std::string, Json> obj;

    const Json& at(const std::string& k) const
    {
        if (kind != Object) throw std::runtime_error("not an object");
        auto it = obj.find(k);
        if (it == obj.end()) throw std::runtime_error("missing key \"" + k + "\"");
        return it->second;
    }
    const Json& at(size_t i) const
    {
        if (kind != Array || i >= arr.size()) throw std::runtime_error("array index out of range");
        return arr[i];
    }
    double number() const
    {
        if (kind != Number) throw std::runtime_error("not a number");
        return num;
    }
};

class Parser
{
public:
    explicit Parser(const std::string& t) : s(t) {}
    Json parseDocument()
    {
        Json v = value();
        ws();
        if (p != s.size()) fail("trailing characters");
        return v;
    }

private:
    const std::string& s;
    size_t p = 0;
    int depth = 0;  // nesting of the value being parsed: bounded, the parser recurses (a cameras.json nests 3 deep)
    static constexpr int kMaxDepth = 64;
    struct DepthGuard
    {
        int& d;
        explicit DepthGuard(int& dd) : d(dd) { ++d; }
        ~DepthGuard() { --d; }
    };
    [[noreturn]] void fail(const char* m) const { throw std::runtime_error(std::string("json: ") + m + " at offset " + std::to_string(p)); }
    void ws() { while (p < s.size() && (s[p] == ' ' || s[p] == '\n' || s[p] == '\t' || s[p] == '\r')) p++; }
    bool eat(char c) { ws(); if (p < s.size() && s[p] == c) { p++; return true; } return false; }
    void expect(char c) { if (!eat(c)) fail("unexpected character"); }
    Json value()
    {
        DepthGuard guard(depth);
        if (depth > kMaxDepth) fail("nesting deeper than 64 levels");
        ws();
        if (p >= s.size()) fail("unexpected end");
        const char c = s[p];
        Json v;
        if (c == '{')
        {
            p++;
            v.kind = Json::Object;
            if (eat('}')) return v;
            do
            {
                ws();
                Json k = value();
                if (k.kind != Json::String) fail("object key must be a string");
                expect(':');
                v.obj[k.str] = value();
            } while (eat(','));
            expect('}');
        }
        else if (c == '[')
        {
            p++;
            v.kind = Json::Array;
            if (eat(']')) return v;
            do v.arr.push_back(value()); while (eat(','));
            expect(']');
        }
        else if (c == '"')
        {
            p++;
            v.kind = Json::String;
            while (p < s.size() && s[p] != '"')
            {
                if (s[p] == '\\' && p + 1 < s.size()) { p++; v.str.push_back(s[p] == 'n' ? '\n' : s[p] == 't' ? '\t' : s[p]); p++; }
                else v.str.push_back(s[p++]);
            }
            if (p >= s.size()) fail("unterminated string");
            p++;
        }
        else if (s.compare(p, 4, "true") == 0) { v.kind = Json::Bool; v.b = true; p += 4; }
        else if (s.compare(p, 5, "false") == 0) { v.kind = Json::Bool; p += 5; }
        else if (s.compare(p, 4, "null") == 0) { p += 4; }
        else
        {
            char* end = nullptr;
            v.num = std::strtod(s.c_str() + p, &end);
            if (end == s.c_str() + p) fail("bad number");
            v.kind = Json::Number;
            p = (size_t)(end - s.c_str());
        }
        return v;
    }
};
}  // namespace

bool CamerasConfig::ImportJson(const std::string& jsonFilename)
{
    std::ifstream f(jsonFilename);
    if (f.fail()) { error = "failed to open " + jsonFilename; return false; }
    std::stringstream buf;
    buf << f.rdbuf();
    const std::string text = buf.str();
    try
    {
        const Json doc = Parser(text).parseDocument();
        if (doc.kind != Json::Array) throw std::runtime_error("top level must be an array");
        std::vector<Camera> cams;
        for (const Json& o : doc.arr)
        {
            (void)o.at("id").number();
            const Json& jp = o.at("position");
            const float pos[3] = {(float)jp.at(0).number(), (float)jp.at(1).number(), (float)jp.at(2).number()};
            const Json& jr = o.at("rotation");
            float R[3][3];
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) R[r][c] = (float)jr.at(r).at(c).number();
            const int width = (int)o.at("width").number(), height = (int)o.at("height").number();
            const float fx = (float)o.at("fx").number();
            (void)o.at("fy").number();
            Camera cam;
            cam.fov = splat::vec2{2.0f * atanf(width / (2.0f * fx)), 2.0f * atanf(height / (2.0f * fx))};
            cam.mat = splat::mat4::identity();
            for (int r = 0; r < 3; r++)
            {
                cam.mat.at(0, r) = R[r][0];
                cam.mat.at(1, r) = -R[r][1];
                cam.mat.at(2, r) = -R[r][2];
                cam.mat.at(3, r) = pos[r];
            }
            cams.push_back(cam);
        }
        cameraVec.insert(cameraVec.end(), cams.begin(), cams.end());  // the reference appends (emplace_back per entry)
    }
    catch (const std::exception& e)
    {
        error = std::string("CamerasConfig::ImportJson: ") + e.what();
        return false;
    }
    return true;
}

void CamerasConfig::EstimateFloorPlane(splat::vec3& normalOut, splat::vec3& posOut) const
{
    if (cameraVec.empty())
    {
        normalOut = splat::vec3{0.0f, 1.0f, 0.0f};
        posOut = splat::vec3{0.0f, 0.0f, 0.0f};
        return;  // (the reference falls through and divides by zero here; SURVEY q-list style fix)
    }
    const float weight = 1.0f / (float)cameraVec.size();
    float up[3] = {0.0f, 0.0f, 0.0f};
    for (const Camera& c : cameraVec)
        for (int r = 0; r < 3; r++) up[r] += weight * c.mat.at(1, r);
    // SafeNormalize (src/core/util.cpp:128-139) -> glm::normalize = v * inversesqrt(dot(v, v)): multiply by the reciprocal,
    // not divide by the length (1 ulp apart; tests/test_ref_pin.py holds this to the reference's source bit for bit)
    const float len2 = up[0] * up[0] + up[1] * up[1] + up[2] * up[2];
    if (std::sqrt(len2) > 0.0f)
    {
        const float inv = 1.0f / std::sqrt(len2);
        up[0] *= inv; up[1] *= inv; up[2] *= inv;
    }
    else { up[0] = 0.0f; up[1] = 1.0f; up[2] = 0.0f; }
    float dist = 0.0f;
    for (const Camera& c : cameraVec) dist += weight * (c.mat.at(3, 0) * up[0] + c.mat.at(3, 1) * up[1] + c.mat.at(3, 2) * up[2]);
    normalOut = splat::vec3{up[0], up[1], up[2]};
    posOut = splat::vec3{up[0] * dist, up[1] * dist, up[2] * dist};
}
