// glsl_shim.h -- just enough GLSL (types, constructors, swizzles, built-ins) to compile the reference's shader
// SOURCE TEXT as C++ and run it on the CPU.  TEST INFRASTRUCTURE (oracle/_ref), never part of the product.
//
// Used by oracle/build_ref.py: shader/presort_compute.glsl, splat_vert.glsl, splat_geom.glsl, splat_frag.glsl of
// /root/reference are included (after a few mechanical rewrites of declarations GLSL has and C++ lacks) inside
// one namespace per stage/variant.  Where GLSL leaves evaluation order to the driver (mat*vec, mat*mat, dot,
// normalize) this shim sums left to right, column by column -- the same order the oracle restates -- so records
// produced by the real shader text can be compared with the oracle bit for bit (build with -ffp-contract=off).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace glsl
{
typedef unsigned int uint;

struct vec2;
struct vec3;
struct vec4;

template <int A, int B> struct Swz2  // aliases the owner's float storage
{
    float d[4];
    operator vec2() const;
    Swz2& operator=(const vec2& v);
};
template <int A, int B, int C> struct Swz3
{
    float d[4];
    operator vec3() const;
    Swz3& operator=(const vec3& v);
};

struct vec2
{
    float x, y;
    vec2() : x(0), y(0) {}
    vec2(float a, float b) : x(a), y(b) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
struct vec3
{
    union
    {
        struct { float x, y, z; };
        struct { float r, g, b; };
    };
    vec3() : x(0), y(0), z(0) {}
    vec3(float a, float b_, float c) : x(a), y(b_), z(c) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
struct vec4
{
    union
    {
        struct { float x, y, z, w; };
        struct { float r, g, b, a; };
        Swz3<0, 1, 2> xyz;
        Swz3<0, 1, 2> rgb;
        Swz2<0, 1> xy;
        Swz2<2, 3> zw;
    };
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(float a_, float b_, float c, float d) : x(a_), y(b_), z(c), w(d) {}
    vec4(const vec3& v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    vec4(const vec2& p, const vec2& q) : x(p.x), y(p.y), z(q.x), w(q.y) {}
    vec4(const vec4& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
    vec4& operator=(const vec4& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
struct uvec3 { uint x, y, z; };

template <int A, int B> Swz2<A, B>::operator vec2() const { return vec2(d[A], d[B]); }
template <int A, int B> Swz2<A, B>& Swz2<A, B>::operator=(const vec2& v) { d[A] = v.x; d[B] = v.y; return *this; }
template <int A, int B, int C> Swz3<A, B, C>::operator vec3() const { return vec3(d[A], d[B], d[C]); }
template <int A, int B, int C> Swz3<A, B, C>& Swz3<A, B, C>::operator=(const vec3& v) { d[A] = v.x; d[B] = v.y; d[C] = v.z; return *this; }

// ---- vector arithmetic (component-wise, like GLSL)
inline vec2 operator+(const vec2& a, const vec2& b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(const vec2& a, const vec2& b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator-(const vec2& a) { return vec2(-a.x, -a.y); }
inline vec2 operator*(const vec2& a, float s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator*(float s, const vec2& a) { return vec2(s * a.x, s * a.y); }
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator-(const vec4& a, const vec4& b) { return vec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline vec4 operator*(const vec4& a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4 operator*(float s, const vec4& a) { return vec4(s * a.x, s * a.y, s * a.z, s * a.w); }

inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline float dot(const vec3& a, const vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline vec3 normalize(const vec3& v)
{
    const float len = sqrtf(dot(v, v));
    return vec3(v.x / len, v.y / len, v.z / len);
}

// ---- matrices, column-major: m[col][row]
struct mat3;
struct mat2
{
    vec2 c[2];
    mat2() {}
    mat2(const vec2& a, const vec2& b) { c[0] = a; c[1] = b; }
    explicit mat2(const mat3& m);  // upper-left 2x2
    vec2& operator[](int i) { return c[i]; }
    const vec2& operator[](int i) const { return c[i]; }
};
struct mat4;
struct mat3
{
    vec3 c[3];
    mat3() {}
    mat3(const vec3& a, const vec3& b, const vec3& d) { c[0] = a; c[1] = b; c[2] = d; }
    explicit mat3(const mat4& m);
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
struct mat4
{
    vec4 c[4];
    mat4() {}
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
inline mat3::mat3(const mat4& m)
{
    for (int i = 0; i < 3; i++) c[i] = vec3(m[i].x, m[i].y, m[i].z);
}
inline mat2::mat2(const mat3& m)
{
    c[0] = vec2(m[0].x, m[0].y);
    c[1] = vec2(m[1].x, m[1].y);
}

inline vec2 operator*(const mat2& m, const vec2& v) { return vec2(m[0].x * v.x + m[1].x * v.y, m[0].y * v.x + m[1].y * v.y); }
inline vec4 operator*(const mat4& m, const vec4& v)
{
    vec4 r;
    for (int row = 0; row < 4; row++) r[row] = ((m[0][row] * v.x + m[1][row] * v.y) + m[2][row] * v.z) + m[3][row] * v.w;
    return r;
}
inline mat3 operator*(const mat3& a, const mat3& b)
{
    mat3 r;
    for (int col = 0; col < 3; col++)
        for (int row = 0; row < 3; row++) r[col][row] = (a[0][row] * b[col][0] + a[1][row] * b[col][1]) + a[2][row] * b[col][2];
    return r;
}
inline mat3 transpose(const mat3& a)
{
    mat3 r;
    for (int col = 0; col < 3; col++)
        for (int row = 0; row < 3; row++) r[col][row] = a[row][col];
    return r;
}

// ---- built-in functions the four shaders call (float overloads only)
inline float sqrt(float x) { return sqrtf(x); }
inline float exp(float x) { return expf(x); }
inline float pow(float x, float y) { return powf(x, y); }
inline float cos(float x) { return cosf(x); }
inline float sin(float x) { return sinf(x); }
inline float atan(float y, float x) { return atan2f(y, x); }
inline float radians(float deg) { return deg * 0.017453292519943295f; }

// ---- resources
template <typename T> struct Buffer  // "buffer Block { T name[]; }" with .length()
{
    std::vector<T> v;
    T& operator[](uint i) { return v[i]; }
    int length() const { return (int)v.size(); }
};
struct atomic_uint { uint value = 0; };
inline uint atomicCounterIncrement(atomic_uint& c) { return c.value++; }
}  // namespace glsl
