// camerasconfig.cpp -- cameras.json (INRIA gaussian-splatting output) -> camera-to-world matrices.
// Contract: src/camerasconfig.cpp:20-95 of the reference.  Per entry {id, position[3], rotation[3][3], width, height,
// fx, fy}: mat = columns (R col 0, -R col 1, -R col 2, position) -- the JSON rotation is camera-to-world with y down /
// z forward, the renderer wants GL's y up / -z forward (:50-53); fov = 2 atan(size / (2 fx)) for both axes (the
// reference uses fx for the vertical fov too; fov only feeds gizmos).  Any malformed or missing field
// makes ImportJson return false, like the nlohmann exception path (:58-63); the cameras read before it stay in the list.
#include "splatapult_b200/camerasconfig.h"

#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>

namespace
{
// A small JSON reader that makes the decisions nlohmann/json 3.11 makes (the library the reference parses with; held to it
// by tests/test_ref_pin.py, where the reference's camerasconfig.cpp runs against the real library): RFC 8259 grammar, no
// extensions (no comments, NaN, leading zeros or '+'), UTF-8 and escape validation inside strings, a leading byte-order mark
// skipped, integers kept as 64-bit integers (so "-0" reads as +0.0f and 2^32+1 as the int 1, as static_cast does there),
// an integer that overflows 64 bits or any fraction / exponent read with strtod, a non-finite result refused.
struct Json
{
    enum Kind { Null, Bool, Int, Uint, Float, String, Array, Object } kind = Null;
    bool b = false;
    long long i = 0;
    unsigned long long u = 0;
    double d = 0;
    std::string str;
    std::vector<Json> arr;
    std::map<std::string, Json> obj;  // key order = iteration order of a top-level object, as with nlohmann's std::map

    // o["key"] / a[i] of the reference are NON-const: a missing member reads as null and the get<T>() that follows throws;
    // a value of the wrong kind throws at once.  Either way ImportJson returns false, so both are "throw" here.
    const Json& at(const std::string& k) const
    {
        if (kind != Object) throw std::runtime_error(kind == Null ? "missing key \"" + k + "\"" : "not an object");
        auto it = obj.find(k);
        if (it == obj.end()) throw std::runtime_error("missing key \"" + k + "\"");
        return it->second;
    }
    const Json& at(size_t idx) const
    {
        if (kind != Array || idx >= arr.size()) throw std::runtime_error("array index out of range");
        return arr[idx];
    }
    // get<T>(): numbers and booleans convert with static_cast from the type they were parsed as; everything else throws
    template <typename T> T get() const
    {
        switch (kind)
        {
        case Int: return static_cast<T>(i);
        case Uint: return static_cast<T>(u);
        case Float: return static_cast<T>(d);
        case Bool: return static_cast<T>(b);
        default: throw std::runtime_error("not a number");
        }
    }
};

class Parser
{
public:
    explicit Parser(const std::string& t) : s(t) {}
    Json parseDocument()
    {
        if (s.size() >= 3 && (unsigned char)s[0] == 0xEF && (unsigned char)s[1] == 0xBB && (unsigned char)s[2] == 0xBF) p = 3;  // byte-order mark
        Json v = value();
        ws();
        if (p != s.size()) fail("trailing characters");
        return v;
    }

private:
    const std::string& s;
    size_t p = 0;
    int depth = 0;  // nesting of the value being parsed: bounded, the parser recurses (a cameras.json nests 3 deep)
    static constexpr int kMaxDepth = 256;
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
    bool digit() const { return p < s.size() && s[p] >= '0' && s[p] <= '9'; }
    unsigned char byte() { if (p >= s.size()) fail("unterminated string"); return (unsigned char)s[p++]; }
    void continuation(std::string& out, unsigned lo, unsigned hi)
    {
        const unsigned char c = byte();
        if (c < lo || c > hi) fail("invalid UTF-8 in string");
        out.push_back((char)c);
    }
    unsigned hex4()
    {
        unsigned cp = 0;
        for (int k = 0; k < 4; k++)
        {
            const unsigned char h = byte();
            cp <<= 4;
            if (h >= '0' && h <= '9') cp |= h - '0';
            else if (h >= 'a' && h <= 'f') cp |= h - 'a' + 10;
            else if (h >= 'A' && h <= 'F') cp |= h - 'A' + 10;
            else fail("\\u must be followed by four hex digits");
        }
        return cp;
    }
    static void utf8(std::string& out, unsigned cp)
    {
        if (cp < 0x80) out.push_back((char)cp);
        else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
        else if (cp < 0x10000) { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
        else { out.push_back((char)(0xF0 | (cp >> 18))); out.push_back((char)(0x80 | ((cp >> 12) & 0x3F))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
    }
    std::string string()
    {
        std::string out;
        p++;  // opening quote
        for (;;)
        {
            const unsigned char c = byte();
            if (c == '"') return out;
            if (c < 0x20) fail("control character in string must be escaped");
            if (c == '\\')
            {
                const unsigned char e = byte();
                switch (e)
                {
                case '"': out.push_back('"'); break;
                case '\\': out.push_back('\\'); break;
                case '/': out.push_back('/'); break;
                case 'b': out.push_back('\b'); break;
                case 'f': out.push_back('\f'); break;
                case 'n': out.push_back('\n'); break;
                case 'r': out.push_back('\r'); break;
                case 't': out.push_back('\t'); break;
                case 'u':
                {
                    unsigned cp = hex4();
                    if (cp >= 0xD800 && cp <= 0xDBFF)
                    {
                        if (byte() != '\\' || byte() != 'u') fail("high surrogate must be followed by a low surrogate");
                        const unsigned lo = hex4();
                        if (lo < 0xDC00 || lo > 0xDFFF) fail("high surrogate must be followed by a low surrogate");
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    }
                    else if (cp >= 0xDC00 && cp <= 0xDFFF) fail("low surrogate without a high surrogate");
                    utf8(out, cp);
                    break;
                }
                default: fail("forbidden character after backslash");
                }
                continue;
            }
            out.push_back((char)c);
            if (c < 0x80) continue;
            // well-formed UTF-8 only (Unicode table 3-7), as nlohmann's lexer checks
            if (c >= 0xC2 && c <= 0xDF) continuation(out, 0x80, 0xBF);
            else if (c == 0xE0) { continuation(out, 0xA0, 0xBF); continuation(out, 0x80, 0xBF); }
            else if ((c >= 0xE1 && c <= 0xEC) || c == 0xEE || c == 0xEF) { continuation(out, 0x80, 0xBF); continuation(out, 0x80, 0xBF); }
            else if (c == 0xED) { continuation(out, 0x80, 0x9F); continuation(out, 0x80, 0xBF); }
            else if (c == 0xF0) { continuation(out, 0x90, 0xBF); continuation(out, 0x80, 0xBF); continuation(out, 0x80, 0xBF); }
            else if (c >= 0xF1 && c <= 0xF3) { continuation(out, 0x80, 0xBF); continuation(out, 0x80, 0xBF); continuation(out, 0x80, 0xBF); }
            else if (c == 0xF4) { continuation(out, 0x80, 0x8F); continuation(out, 0x80, 0xBF); continuation(out, 0x80, 0xBF); }
            else fail("invalid UTF-8 in string");
        }
    }
    Json number()
    {
        const size_t start = p;
        bool isFloat = false;
        if (s[p] == '-') p++;
        if (!digit()) fail("bad number");
        if (s[p] == '0') p++;  // no leading zeros: "01" stops here and fails as trailing garbage of the enclosing value
        else while (digit()) p++;
        if (p < s.size() && s[p] == '.')
        {
            isFloat = true;
            p++;
            if (!digit()) fail("digit expected after '.'");
            while (digit()) p++;
        }
        if (p < s.size() && (s[p] == 'e' || s[p] == 'E'))
        {
            isFloat = true;
            p++;
            if (p < s.size() && (s[p] == '+' || s[p] == '-')) p++;
            if (!digit()) fail("digit expected in exponent");
            while (digit()) p++;
        }
        const std::string tok = s.substr(start, p - start);
        Json v;
        if (!isFloat)
        {
            errno = 0;
            if (tok[0] == '-')
            {
                const long long x = std::strtoll(tok.c_str(), nullptr, 10);
                if (errno == 0) { v.kind = Json::Int; v.i = x; return v; }
            }
            else
            {
                const unsigned long long x = std::strtoull(tok.c_str(), nullptr, 10);
                if (errno == 0) { v.kind = Json::Uint; v.u = x; return v; }
            }
        }
        v.kind = Json::Float;
        v.d = std::strtod(tok.c_str(), nullptr);
        if (!std::isfinite(v.d)) fail("number overflow");
        return v;
    }
    Json value()
    {
        DepthGuard guard(depth);
        if (depth > kMaxDepth) fail("nesting deeper than 256 levels");
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
                if (p >= s.size() || s[p] != '"') fail("object key must be a string");
                const std::string k = string();
                expect(':');
                v.obj[k] = value();  // a repeated key keeps the last value
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
        else if (c == '"') { v.kind = Json::String; v.str = string(); }
        else if (s.compare(p, 4, "true") == 0) { v.kind = Json::Bool; v.b = true; p += 4; }
        else if (s.compare(p, 5, "false") == 0) { v.kind = Json::Bool; p += 5; }
        else if (s.compare(p, 4, "null") == 0) { p += 4; }
        else if (c == '-' || (c >= '0' && c <= '9')) v = number();
        else fail("invalid literal");
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
        // `for (auto&& o : data)` (camerasconfig.cpp:31): the elements of an array, the VALUES of an object in key order,
        // nothing for null, a primitive itself once (which then fails on o["id"])
        std::vector<const Json*> entries;
        if (doc.kind == Json::Array) for (const Json& o : doc.arr) entries.push_back(&o);
        else if (doc.kind == Json::Object) for (const auto& kv : doc.obj) entries.push_back(&kv.second);
        else if (doc.kind != Json::Null) entries.push_back(&doc);
        for (const Json* e : entries)
        {
            const Json& o = *e;
            (void)o.at("id").get<int>();
            const Json& jp = o.at("position");
            const float pos[3] = {jp.at(0).get<float>(), jp.at(1).get<float>(), jp.at(2).get<float>()};
            const Json& jr = o.at("rotation");
            float R[3][3];
            // the reference reads jRot[0][0], [1][0], [2][0], [0][1], ... (column by column): same values, and the first
            // missing element fails either way
            for (int c = 0; c < 3; c++)
                for (int r = 0; r < 3; r++) R[r][c] = jr.at(r).at(c).get<float>();
            const int width = o.at("width").get<int>(), height = o.at("height").get<int>();
            const float fx = o.at("fx").get<float>();
            (void)o.at("fy").get<float>();
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
            cameraVec.push_back(cam);  // per entry, as the reference does: the entries before a bad one stay (:55)
        }
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
