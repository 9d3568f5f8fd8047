// splatmath.h -- glm-layout-compatible PODs and the handful of glm operations the splat hot path uses.
//
// The reference passes glm::mat4 / glm::vec4 / glm::vec2 through SplatRenderer::Sort/Render
// (src/splatrenderer.h:32-37).  glm is not a dependency here; these PODs have the same memory layout
// (column-major float[16], tightly packed vectors) so a glm object can be reinterpret_cast / memcpy'd.
// Operation order follows glm 1.0.1's published algorithms so results are reproducible bit for bit:
//   inverse(mat4)  glm/detail/func_matrix.inl compute_inverse<4,4>   (used at src/splatrenderer.cpp:161,327)
//   mat4 * mat4    glm/detail/type_mat4x4.inl                         (src/splatrenderer.cpp:175)
//   perspective    glm/ext/matrix_clip_space.inl perspectiveRH_NO     (src/app.cpp:1042)
//   quat -> mat3   glm/gtc/quaternion.inl normalize + mat3_cast       (src/gaussiancloud.cpp:86-94)
// Host translation units that include this header must be compiled with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstring>

// the C++ drop-in classes (GaussianCloud, SplatRenderer, CamerasConfig, Ply) are exported from libsplat_b200.so next to
// the C ABI; everything else in the library has hidden visibility
#if defined(_WIN32)
#define SPLAT_CPP_API
#else
#define SPLAT_CPP_API __attribute__((visibility("default")))
#endif

namespace splat
{

struct vec2 { float x, y; };
struct vec3 { float x, y, z; };
struct vec4 { float x, y, z, w; };

struct mat3
{
    float m[9];  // m[col * 3 + row]
    float& at(int col, int row) { return m[col * 3 + row]; }
    float at(int col, int row) const { return m[col * 3 + row]; }
};

struct mat4
{
    float m[16];  // m[col * 4 + row]
    float& at(int col, int row) { return m[col * 4 + row]; }
    float at(int col, int row) const { return m[col * 4 + row]; }
    vec4& operator[](int col) { return *reinterpret_cast<vec4*>(m + 4 * col); }
    const vec4& operator[](int col) const { return *reinterpret_cast<const vec4*>(m + 4 * col); }
    static mat4 identity()
    {
        mat4 r;
        std::memset(r.m, 0, sizeof r.m);
        r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f;
        return r;
    }
    static mat4 from(const float* p)
    {
        mat4 r;
        std::memcpy(r.m, p, sizeof r.m);
        return r;
    }
};

inline mat4 translate(float x, float y, float z)
{
    mat4 r = mat4::identity();
    r.at(3, 0) = x; r.at(3, 1) = y; r.at(3, 2) = z;
    return r;
}

inline mat4 mul(const mat4& a, const mat4& b)
{
    mat4 r;
    for (int c = 0; c < 4; c++)
        for (int row = 0; row < 4; row++)
            r.at(c, row) = ((a.at(0, row) * b.at(c, 0) + a.at(1, row) * b.at(c, 1)) + a.at(2, row) * b.at(c, 2)) +
                           a.at(3, row) * b.at(c, 3);
    return r;
}

inline mat3 mul(const mat3& a, const mat3& b)
{
    mat3 r;
    for (int c = 0; c < 3; c++)
        for (int row = 0; row < 3; row++)
            r.at(c, row) = (a.at(0, row) * b.at(c, 0) + a.at(1, row) * b.at(c, 1)) + a.at(2, row) * b.at(c, 2);
    return r;
}

inline mat3 transpose(const mat3& a)
{
    mat3 r;
    for (int c = 0; c < 3; c++)
        for (int row = 0; row < 3; row++) r.at(c, row) = a.at(row, c);
    return r;
}

// cofactor expansion, 2x2 sub-determinants shared between columns (glm compute_inverse<4,4>)
inline mat4 inverse(const mat4& s)
{
    const float c00 = s.at(2,2) * s.at(3,3) - s.at(3,2) * s.at(2,3);
    const float c02 = s.at(1,2) * s.at(3,3) - s.at(3,2) * s.at(1,3);
    const float c03 = s.at(1,2) * s.at(2,3) - s.at(2,2) * s.at(1,3);
    const float c04 = s.at(2,1) * s.at(3,3) - s.at(3,1) * s.at(2,3);
    const float c06 = s.at(1,1) * s.at(3,3) - s.at(3,1) * s.at(1,3);
    const float c07 = s.at(1,1) * s.at(2,3) - s.at(2,1) * s.at(1,3);
    const float c08 = s.at(2,1) * s.at(3,2) - s.at(3,1) * s.at(2,2);
    const float c10 = s.at(1,1) * s.at(3,2) - s.at(3,1) * s.at(1,2);
    const float c11 = s.at(1,1) * s.at(2,2) - s.at(2,1) * s.at(1,2);
    const float c12 = s.at(2,0) * s.at(3,3) - s.at(3,0) * s.at(2,3);
    const float c14 = s.at(1,0) * s.at(3,3) - s.at(3,0) * s.at(1,3);
    const float c15 = s.at(1,0) * s.at(2,3) - s.at(2,0) * s.at(1,3);
    const float c16 = s.at(2,0) * s.at(3,2) - s.at(3,0) * s.at(2,2);
    const float c18 = s.at(1,0) * s.at(3,2) - s.at(3,0) * s.at(1,2);
    const float c19 = s.at(1,0) * s.at(2,2) - s.at(2,0) * s.at(1,2);
    const float c20 = s.at(2,0) * s.at(3,1) - s.at(3,0) * s.at(2,1);
    const float c22 = s.at(1,0) * s.at(3,1) - s.at(3,0) * s.at(1,1);
    const float c23 = s.at(1,0) * s.at(2,1) - s.at(2,0) * s.at(1,1);

    const float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
    const float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
    const float v0[4] = {s.at(1,0), s.at(0,0), s.at(0,0), s.at(0,0)};
    const float v1[4] = {s.at(1,1), s.at(0,1), s.at(0,1), s.at(0,1)};
    const float v2[4] = {s.at(1,2), s.at(0,2), s.at(0,2), s.at(0,2)};
    const float v3[4] = {s.at(1,3), s.at(0,3), s.at(0,3), s.at(0,3)};
    mat4 adj;
    for (int i = 0; i < 4; i++)
    {
        const float sa = (i & 1) ? -1.0f : 1.0f, sb = -sa;
        adj.at(0, i) = (v1[i] * f0[i] - v2[i] * f1[i] + v3[i] * f2[i]) * sa;
        adj.at(1, i) = (v0[i] * f0[i] - v2[i] * f3[i] + v3[i] * f4[i]) * sb;
        adj.at(2, i) = (v0[i] * f1[i] - v1[i] * f3[i] + v3[i] * f5[i]) * sa;
        adj.at(3, i) = (v0[i] * f2[i] - v1[i] * f4[i] + v2[i] * f5[i]) * sb;
    }
    const float d0 = s.at(0,0) * adj.at(0,0), d1 = s.at(0,1) * adj.at(1,0);
    const float d2 = s.at(0,2) * adj.at(2,0), d3 = s.at(0,3) * adj.at(3,0);
    const float rdet = 1.0f / ((d0 + d1) + (d2 + d3));
    mat4 r;
    for (int i = 0; i < 16; i++) r.m[i] = adj.m[i] * rdet;
    return r;
}

// right-handed, clip z in [-1,1] (OpenGL)
inline mat4 perspective(float fovy, float aspect, float zNear, float zFar)
{
    const float t = std::tan(fovy / 2.0f);
    mat4 r;
    std::memset(r.m, 0, sizeof r.m);
    r.at(0, 0) = 1.0f / (aspect * t);
    r.at(1, 1) = 1.0f / (t);
    r.at(2, 2) = -(zFar + zNear) / (zFar - zNear);
    r.at(2, 3) = -1.0f;
    r.at(3, 2) = -(2.0f * zFar * zNear) / (zFar - zNear);
    return r;
}

// unit quaternion (w,x,y,z) -> rotation matrix; zero-length input maps to identity like glm::normalize
inline mat3 rotationFromQuat(float w, float x, float y, float z)
{
    const float len = std::sqrt((w * w + x * x) + (y * y + z * z));
    if (len <= 0.0f) { w = 1.0f; x = y = z = 0.0f; }
    else { const float s = 1.0f / len; w = w * s; x = x * s; y = y * s; z = z * s; }
    const float xx = x * x, yy = y * y, zz = z * z, xz = x * z, xy = x * y, yz = y * z;
    const float wx = w * x, wy = w * y, wz = w * z;
    mat3 r;
    r.at(0, 0) = 1.0f - 2.0f * (yy + zz); r.at(0, 1) = 2.0f * (xy + wz);        r.at(0, 2) = 2.0f * (xz - wy);
    r.at(1, 0) = 2.0f * (xy - wz);        r.at(1, 1) = 1.0f - 2.0f * (xx + zz); r.at(1, 2) = 2.0f * (yz + wx);
    r.at(2, 0) = 2.0f * (xz + wy);        r.at(2, 1) = 2.0f * (yz - wx);        r.at(2, 2) = 1.0f - 2.0f * (xx + yy);
    return r;
}

}  // namespace splat
