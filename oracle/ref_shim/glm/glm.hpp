// Minimal stand-in for the parts of glm 1.0.1 that the reference's HOST sources on the hot path use
// (src/gaussiancloud.cpp, src/camerasconfig.cpp, src/core/util.h).  TEST INFRASTRUCTURE for oracle/_ref only.
// Operation order follows glm's published implementations (type_mat3x3.inl, gtc/quaternion.inl, func_geometric.inl).
#pragma once
#include <cmath>
#include <cstddef>

namespace glm
{
struct vec2 { float x, y; vec2() : x(0), y(0) {} vec2(float a, float b) : x(a), y(b) {} };
struct vec3
{
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
    vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
    vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    explicit vec3(const struct vec4& v);
};
struct vec4
{
    float x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(const vec3& v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
};
inline vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator*(float s, const vec3& v) { return vec3(s * v.x, s * v.y, s * v.z); }
inline vec3 operator*(const vec3& v, float s) { return vec3(v.x * s, v.y * s, v.z * s); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float length(const vec3& v) { return std::sqrt(dot(v, v)); }
inline float distance(const vec3& a, const vec3& b) { return length(b - a); }
inline vec3 normalize(const vec3& v) { return v * (1.0f / std::sqrt(dot(v, v))); }  // func_geometric.inl: v * inversesqrt(dot(v, v))

struct quat
{
    float x, y, z, w;  // storage order of glm's default build; constructor takes (w, x, y, z)
    quat() : x(0), y(0), z(0), w(1) {}
    quat(float w_, float x_, float y_, float z_) : x(x_), y(y_), z(z_), w(w_) {}
    explicit quat(const struct mat3& m);
};
inline float dot(const quat& a, const quat& b)
{
    const float tw = a.w * b.w, tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z;  // compute_dot<qua>: vec4(w,x,y,z) products
    return (tw + tx) + (ty + tz);
}
inline quat normalize(const quat& q)
{
    const float len = std::sqrt(dot(q, q));
    if (len <= 0.0f) return quat(1.0f, 0.0f, 0.0f, 0.0f);
    const float oneOverLen = 1.0f / len;
    return quat(q.w * oneOverLen, q.x * oneOverLen, q.y * oneOverLen, q.z * oneOverLen);
}

struct mat2 { vec2 c[2]; };
struct mat3
{
    vec3 c[3];
    mat3() {}
    mat3(const vec3& a, const vec3& b, const vec3& d) { c[0] = a; c[1] = b; c[2] = d; }
    mat3(float x0, float y0, float z0, float x1, float y1, float z1, float x2, float y2, float z2)
    {
        c[0] = vec3(x0, y0, z0); c[1] = vec3(x1, y1, z1); c[2] = vec3(x2, y2, z2);
    }
    explicit mat3(const quat& q)  // mat3_cast
    {
        const float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z, qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z;
        const float qwx = q.w * q.x, qwy = q.w * q.y, qwz = q.w * q.z;
        c[0] = vec3(1.0f - 2.0f * (qyy + qzz), 2.0f * (qxy + qwz), 2.0f * (qxz - qwy));
        c[1] = vec3(2.0f * (qxy - qwz), 1.0f - 2.0f * (qxx + qzz), 2.0f * (qyz + qwx));
        c[2] = vec3(2.0f * (qxz + qwy), 2.0f * (qyz - qwx), 1.0f - 2.0f * (qxx + qyy));
    }
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
struct mat4
{
    vec4 c[4];
    mat4() {}
    mat4(const vec4& a, const vec4& b, const vec4& d, const vec4& e) { c[0] = a; c[1] = b; c[2] = d; c[3] = e; }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
    float& at(int col, int row) { return (&c[col].x)[row]; }
    float at(int col, int row) const { return (&c[col].x)[row]; }
};
// type_mat4x4.inl operator*
inline mat4 operator*(const mat4& a, const mat4& b)
{
    mat4 r;
    for (int col = 0; col < 4; col++)
        for (int row = 0; row < 4; row++)
            r.at(col, row) = a.at(0, row) * b.at(col, 0) + a.at(1, row) * b.at(col, 1) + a.at(2, row) * b.at(col, 2) + a.at(3, row) * b.at(col, 3);
    return r;
}
// func_matrix.inl compute_inverse<4,4>
inline mat4 inverse(const mat4& m)
{
#define G(c, r) m.at(c, r)
    const float Coef00 = G(2,2) * G(3,3) - G(3,2) * G(2,3), Coef02 = G(1,2) * G(3,3) - G(3,2) * G(1,3), Coef03 = G(1,2) * G(2,3) - G(2,2) * G(1,3);
    const float Coef04 = G(2,1) * G(3,3) - G(3,1) * G(2,3), Coef06 = G(1,1) * G(3,3) - G(3,1) * G(1,3), Coef07 = G(1,1) * G(2,3) - G(2,1) * G(1,3);
    const float Coef08 = G(2,1) * G(3,2) - G(3,1) * G(2,2), Coef10 = G(1,1) * G(3,2) - G(3,1) * G(1,2), Coef11 = G(1,1) * G(2,2) - G(2,1) * G(1,2);
    const float Coef12 = G(2,0) * G(3,3) - G(3,0) * G(2,3), Coef14 = G(1,0) * G(3,3) - G(3,0) * G(1,3), Coef15 = G(1,0) * G(2,3) - G(2,0) * G(1,3);
    const float Coef16 = G(2,0) * G(3,2) - G(3,0) * G(2,2), Coef18 = G(1,0) * G(3,2) - G(3,0) * G(1,2), Coef19 = G(1,0) * G(2,2) - G(2,0) * G(1,2);
    const float Coef20 = G(2,0) * G(3,1) - G(3,0) * G(2,1), Coef22 = G(1,0) * G(3,1) - G(3,0) * G(1,1), Coef23 = G(1,0) * G(2,1) - G(2,0) * G(1,1);
    const float Fac0[4] = {Coef00, Coef00, Coef02, Coef03}, Fac1[4] = {Coef04, Coef04, Coef06, Coef07}, Fac2[4] = {Coef08, Coef08, Coef10, Coef11};
    const float Fac3[4] = {Coef12, Coef12, Coef14, Coef15}, Fac4[4] = {Coef16, Coef16, Coef18, Coef19}, Fac5[4] = {Coef20, Coef20, Coef22, Coef23};
    const float Vec0[4] = {G(1,0), G(0,0), G(0,0), G(0,0)}, Vec1[4] = {G(1,1), G(0,1), G(0,1), G(0,1)};
    const float Vec2[4] = {G(1,2), G(0,2), G(0,2), G(0,2)}, Vec3[4] = {G(1,3), G(0,3), G(0,3), G(0,3)};
    const float SignA[4] = {+1, -1, +1, -1}, SignB[4] = {-1, +1, -1, +1};
    mat4 Inverse;
    for (int i = 0; i < 4; i++)
    {
        Inverse.at(0, i) = (Vec1[i] * Fac0[i] - Vec2[i] * Fac1[i] + Vec3[i] * Fac2[i]) * SignA[i];
        Inverse.at(1, i) = (Vec0[i] * Fac0[i] - Vec2[i] * Fac3[i] + Vec3[i] * Fac4[i]) * SignB[i];
        Inverse.at(2, i) = (Vec0[i] * Fac1[i] - Vec1[i] * Fac3[i] + Vec3[i] * Fac5[i]) * SignA[i];
        Inverse.at(3, i) = (Vec0[i] * Fac2[i] - Vec1[i] * Fac4[i] + Vec2[i] * Fac5[i]) * SignB[i];
    }
    const float Dot1 = (G(0,0) * Inverse.at(0,0) + G(0,1) * Inverse.at(1,0)) + (G(0,2) * Inverse.at(2,0) + G(0,3) * Inverse.at(3,0));
    const float OneOverDeterminant = 1.0f / Dot1;
    mat4 r;
    for (int col = 0; col < 4; col++) for (int row = 0; row < 4; row++) r.at(col, row) = Inverse.at(col, row) * OneOverDeterminant;
    return r;
#undef G
}

inline mat3 operator*(const mat3& m1, const mat3& m2)
{
    mat3 r;
    for (int col = 0; col < 3; col++)
        for (int row = 0; row < 3; row++) r[col][row] = m1[0][row] * m2[col][0] + m1[1][row] * m2[col][1] + m1[2][row] * m2[col][2];
    return r;
}
inline mat3 transpose(const mat3& m)
{
    mat3 r;
    for (int col = 0; col < 3; col++)
        for (int row = 0; row < 3; row++) r[col][row] = m[row][col];
    return r;
}
inline float determinant(const mat3& m)
{
    return +m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2]) +
           m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]);
}
inline quat::quat(const mat3& m)  // quat_cast
{
    const float fourXSquaredMinus1 = m[0][0] - m[1][1] - m[2][2];
    const float fourYSquaredMinus1 = m[1][1] - m[0][0] - m[2][2];
    const float fourZSquaredMinus1 = m[2][2] - m[0][0] - m[1][1];
    const float fourWSquaredMinus1 = m[0][0] + m[1][1] + m[2][2];
    int biggestIndex = 0;
    float fourBiggestSquaredMinus1 = fourWSquaredMinus1;
    if (fourXSquaredMinus1 > fourBiggestSquaredMinus1) { fourBiggestSquaredMinus1 = fourXSquaredMinus1; biggestIndex = 1; }
    if (fourYSquaredMinus1 > fourBiggestSquaredMinus1) { fourBiggestSquaredMinus1 = fourYSquaredMinus1; biggestIndex = 2; }
    if (fourZSquaredMinus1 > fourBiggestSquaredMinus1) { fourBiggestSquaredMinus1 = fourZSquaredMinus1; biggestIndex = 3; }
    const float biggestVal = std::sqrt(fourBiggestSquaredMinus1 + 1.0f) * 0.5f;
    const float mult = 0.25f / biggestVal;
    switch (biggestIndex)
    {
    case 0: w = biggestVal; x = (m[1][2] - m[2][1]) * mult; y = (m[2][0] - m[0][2]) * mult; z = (m[0][1] - m[1][0]) * mult; break;
    case 1: w = (m[1][2] - m[2][1]) * mult; x = biggestVal; y = (m[0][1] + m[1][0]) * mult; z = (m[2][0] + m[0][2]) * mult; break;
    case 2: w = (m[2][0] - m[0][2]) * mult; x = (m[0][1] + m[1][0]) * mult; y = biggestVal; z = (m[1][2] + m[2][1]) * mult; break;
    default: w = (m[0][1] - m[1][0]) * mult; x = (m[2][0] + m[0][2]) * mult; y = (m[1][2] + m[2][1]) * mult; z = biggestVal; break;
    }
}
}  // namespace glm
