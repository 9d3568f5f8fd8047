/*
 * splat_oracle.c -- CPU ORACLE for the splatapult hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A literal, deterministic float32 restatement of the reference's
 *   decode      src/gaussiancloud.cpp:86-94,119-122,138-365 + src/ply.cpp:72-87,140-254
 *   pre-sort    src/splatrenderer.cpp:160-177 + shader/presort_compute.glsl:31-57        (K1)
 *   sort        src/splatrenderer.cpp:206-264 + shader/multi_radixsort*.glsl              (K2/K3)
 *   vertex      src/splatrenderer.cpp:324-335 + shader/splat_vert.glsl:51-127,153-222     (K5)
 *   geometry    shader/splat_geom.glsl:22-54                                              (K6)
 *   fragment    shader/splat_frag.glsl:18-42                                              (K7)
 *   blend/clear src/app.cpp:144-164                                                       (K8)
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (splatapult_b200/) never links, imports or calls anything in this directory.
 *
 * PINNING STATUS (see DESIGN.md "Oracle"): the reference has no tests, golden vectors or known-answer
 * fixtures for this path and its OpenGL pipeline cannot execute in this container.  The oracle is pinned by
 * (a) oracle/_ref (oracle/build_ref.py, tests/test_ref_pin.py): the reference's OWN source -- gaussiancloud.cpp,
 *     ply.cpp, binaryattribute.cpp compiled unmodified against glm/Eigen stand-ins, and the text of
 *     presort_compute / splat_vert / splat_geom / splat_frag .glsl compiled as C++ against a GLSL shim -- agrees
 *     with this file BIT FOR BIT on decode, per-gaussian records, draw order and blended framebuffer;
 * (b) the reference's only data fixture, data/test.ply, cross-checked with the constants hard-coded in
 *     GaussianCloud::InitDebugCloud (gaussiancloud.cpp:505-578);
 * (c) first-principles known-answer tests derived from the shader math (tests/test_oracle_kat.py).
 * What stays outside any pin: the GL driver's own exp/normalize/matrix-product rounding and the fixed-function
 * rasteriser/ROP, which no source in the reference defines (the shims sum left to right).
 *
 * Third-party arithmetic the reference calls but does not vendor (vcpkg submodule 1b2372fd, glm 1.0.1):
 * glm::inverse(mat4), glm::perspective, mat4*mat4, quat normalize / mat3_cast, mat3*mat3.  Their
 * published algorithms are restated below (glm/detail/func_matrix.inl, type_mat4x4.inl,
 * gtc/quaternion.inl, ext/matrix_clip_space.inl).
 *
 * All arithmetic is IEEE binary32, evaluated strictly in source order.  Build with -ffp-contract=off and
 * without -ffast-math (oracle/Makefile does); the CUDA projection kernel is built with -fmad=false so
 * its per-gaussian records are bit-comparable with orc_project().
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------ */
/* reference AoS layout: gaussiancloud.cpp:32-56                                                     */
/* ------------------------------------------------------------------------------------------------ */
enum {
    OFF_POS = 0,   /* posWithAlpha[4] */
    OFF_RSH0 = 4,  /* r_sh0[4] */
    OFF_GSH0 = 8,
    OFF_BSH0 = 12,
    OFF_COV0 = 16, /* cov3_col0[3] */
    OFF_COV1 = 19,
    OFF_COV2 = 22,
    BASE_FLOATS = 25, /* 100 B */
    OFF_RSH1 = 25, OFF_RSH2 = 29, OFF_RSH3 = 33,
    OFF_GSH1 = 37, OFF_GSH2 = 41, OFF_GSH3 = 45,
    OFF_BSH1 = 49, OFF_BSH2 = 53, OFF_BSH3 = 57,
    FULL_FLOATS = 61 /* 244 B */
};

/* one record per gaussian; 12 words = 48 B; mirrors the product's projected record bit for bit */
typedef struct {
    float px, py;             /* screen centre, GL window coords (y up)       splat_vert.glsl:200-203 */
    float i00, i01, i10, i11; /* inverse 2D covariance inv[col][row]          splat_geom.glsl:22-32  */
    float r, g, b;            /* 0.5 + SH, unclamped                           splat_vert.glsl:126    */
    float alpha;              /* posWithAlpha.w                                splat_vert.glsl:156    */
    uint32_t key;             /* 0xFFFFFFFF - uint(depth/far * 2^32)           presort_compute.glsl:53 */
    uint32_t visible;         /* K1 && depth<far && K6                                               */
} orc_record;

/* ------------------------------------------------------------------------------------------------ */
/* glm restatements                                                                                  */
/* ------------------------------------------------------------------------------------------------ */
/* column-major m[col*4+row] like glm::mat4 / GLSL */
#define M(m, c, r) ((m)[(c) * 4 + (r)])

/* glm::inverse(mat4): glm/detail/func_matrix.inl compute_inverse<4,4,...>  (splatrenderer.cpp:161,327) */
static void mat4_inverse(const float* m, float* out)
{
    float Coef00 = M(m,2,2) * M(m,3,3) - M(m,3,2) * M(m,2,3);
    float Coef02 = M(m,1,2) * M(m,3,3) - M(m,3,2) * M(m,1,3);
    float Coef03 = M(m,1,2) * M(m,2,3) - M(m,2,2) * M(m,1,3);
    float Coef04 = M(m,2,1) * M(m,3,3) - M(m,3,1) * M(m,2,3);
    float Coef06 = M(m,1,1) * M(m,3,3) - M(m,3,1) * M(m,1,3);
    float Coef07 = M(m,1,1) * M(m,2,3) - M(m,2,1) * M(m,1,3);
    float Coef08 = M(m,2,1) * M(m,3,2) - M(m,3,1) * M(m,2,2);
    float Coef10 = M(m,1,1) * M(m,3,2) - M(m,3,1) * M(m,1,2);
    float Coef11 = M(m,1,1) * M(m,2,2) - M(m,2,1) * M(m,1,2);
    float Coef12 = M(m,2,0) * M(m,3,3) - M(m,3,0) * M(m,2,3);
    float Coef14 = M(m,1,0) * M(m,3,3) - M(m,3,0) * M(m,1,3);
    float Coef15 = M(m,1,0) * M(m,2,3) - M(m,2,0) * M(m,1,3);
    float Coef16 = M(m,2,0) * M(m,3,2) - M(m,3,0) * M(m,2,2);
    float Coef18 = M(m,1,0) * M(m,3,2) - M(m,3,0) * M(m,1,2);
    float Coef19 = M(m,1,0) * M(m,2,2) - M(m,2,0) * M(m,1,2);
    float Coef20 = M(m,2,0) * M(m,3,1) - M(m,3,0) * M(m,2,1);
    float Coef22 = M(m,1,0) * M(m,3,1) - M(m,3,0) * M(m,1,1);
    float Coef23 = M(m,1,0) * M(m,2,1) - M(m,2,0) * M(m,1,1);

    float Fac0[4] = {Coef00, Coef00, Coef02, Coef03};
    float Fac1[4] = {Coef04, Coef04, Coef06, Coef07};
    float Fac2[4] = {Coef08, Coef08, Coef10, Coef11};
    float Fac3[4] = {Coef12, Coef12, Coef14, Coef15};
    float Fac4[4] = {Coef16, Coef16, Coef18, Coef19};
    float Fac5[4] = {Coef20, Coef20, Coef22, Coef23};
    float Vec0[4] = {M(m,1,0), M(m,0,0), M(m,0,0), M(m,0,0)};
    float Vec1[4] = {M(m,1,1), M(m,0,1), M(m,0,1), M(m,0,1)};
    float Vec2[4] = {M(m,1,2), M(m,0,2), M(m,0,2), M(m,0,2)};
    float Vec3[4] = {M(m,1,3), M(m,0,3), M(m,0,3), M(m,0,3)};
    static const float SignA[4] = {+1.0f, -1.0f, +1.0f, -1.0f};
    static const float SignB[4] = {-1.0f, +1.0f, -1.0f, +1.0f};
    float Inv[16];
    for (int i = 0; i < 4; i++) {
        float Inv0 = Vec1[i] * Fac0[i] - Vec2[i] * Fac1[i] + Vec3[i] * Fac2[i];
        float Inv1 = Vec0[i] * Fac0[i] - Vec2[i] * Fac3[i] + Vec3[i] * Fac4[i];
        float Inv2 = Vec0[i] * Fac1[i] - Vec1[i] * Fac3[i] + Vec3[i] * Fac5[i];
        float Inv3 = Vec0[i] * Fac2[i] - Vec1[i] * Fac4[i] + Vec2[i] * Fac5[i];
        M(Inv,0,i) = Inv0 * SignA[i];
        M(Inv,1,i) = Inv1 * SignB[i];
        M(Inv,2,i) = Inv2 * SignA[i];
        M(Inv,3,i) = Inv3 * SignB[i];
    }
    /* Row0 = (Inv[0][0], Inv[1][0], Inv[2][0], Inv[3][0]); Dot0 = m[0] * Row0 */
    float d0 = M(m,0,0) * M(Inv,0,0);
    float d1 = M(m,0,1) * M(Inv,1,0);
    float d2 = M(m,0,2) * M(Inv,2,0);
    float d3 = M(m,0,3) * M(Inv,3,0);
    float Dot1 = (d0 + d1) + (d2 + d3);
    float OneOverDeterminant = 1.0f / Dot1;
    for (int i = 0; i < 16; i++) out[i] = Inv[i] * OneOverDeterminant;
}

/* glm mat4*mat4: glm/detail/type_mat4x4.inl operator*  (splatrenderer.cpp:175 projMat * modelViewMat) */
static void mat4_mul(const float* a, const float* b, float* out)
{
    float r[16];
    for (int c = 0; c < 4; c++)
        for (int row = 0; row < 4; row++)
            r[c * 4 + row] = ((M(a,0,row) * M(b,c,0) + M(a,1,row) * M(b,c,1)) + M(a,2,row) * M(b,c,2)) +
                             M(a,3,row) * M(b,c,3);
    memcpy(out, r, sizeof r);
}

/* GLSL mat4 * vec4(x,y,z,1): column sum, left to right                                              */
static inline void mat4_xform_point(const float* m, float x, float y, float z, float* o)
{
    for (int row = 0; row < 4; row++)
        o[row] = ((M(m,0,row) * x + M(m,1,row) * y) + M(m,2,row) * z) + M(m,3,row);
}
/* GLSL mat4 * vec4 (general w)                                                                      */
static inline void mat4_xform_vec4(const float* m, const float* v, float* o)
{
    for (int row = 0; row < 4; row++)
        o[row] = ((M(m,0,row) * v[0] + M(m,1,row) * v[1]) + M(m,2,row) * v[2]) + M(m,3,row) * v[3];
}

/* 3x3 column-major m[c*3+r]; C = A*B with left-to-right sums (GLSL mat3*mat3 / glm type_mat3x3.inl) */
#define M3(m, c, r) ((m)[(c) * 3 + (r)])
static inline void mat3_mul(const float* a, const float* b, float* out)
{
    float r[9];
    for (int c = 0; c < 3; c++)
        for (int row = 0; row < 3; row++)
            r[c * 3 + row] = (M3(a,0,row) * M3(b,c,0) + M3(a,1,row) * M3(b,c,1)) + M3(a,2,row) * M3(b,c,2);
    memcpy(out, r, sizeof r);
}
static inline void mat3_transpose(const float* a, float* out)
{
    float r[9];
    for (int c = 0; c < 3; c++)
        for (int row = 0; row < 3; row++) r[c * 3 + row] = M3(a, row, c);
    memcpy(out, r, sizeof r);
}

/* glm::perspective (RH, z in [-1,1]): glm/ext/matrix_clip_space.inl perspectiveRH_NO (app.cpp:1042) */
ORC_API void orc_perspective(float fovy, float aspect, float zNear, float zFar, float* out)
{
    float tanHalfFovy = tanf(fovy / 2.0f);
    memset(out, 0, 16 * sizeof(float));
    M(out,0,0) = 1.0f / (aspect * tanHalfFovy);
    M(out,1,1) = 1.0f / (tanHalfFovy);
    M(out,2,2) = -(zFar + zNear) / (zFar - zNear);
    M(out,2,3) = -1.0f;
    M(out,3,2) = -(2.0f * zFar * zNear) / (zFar - zNear);
}

/* ------------------------------------------------------------------------------------------------ */
/* decode: PLY -> reference AoS                                                                      */
/* ------------------------------------------------------------------------------------------------ */
/* gaussiancloud.cpp:86-94 ComputeCovMatFromRotScale: R*S*S^T*R^T, q=(w,x,y,z)=rot_0..3 normalised     */
static void cov_from_rot_scale(const float rot[4], const float scale[3], float* V /*col-major 3x3*/)
{
    float w = rot[0], x = rot[1], y = rot[2], z = rot[3];
    /* glm::normalize(quat): gtc/quaternion.inl; dot = (w*w + x*x) + (y*y + z*z) */
    float len = sqrtf((w * w + x * x) + (y * y + z * z));
    if (len <= 0.0f) { w = 1.0f; x = 0.0f; y = 0.0f; z = 0.0f; }
    else { float oneOverLen = 1.0f / len; w = w * oneOverLen; x = x * oneOverLen; y = y * oneOverLen; z = z * oneOverLen; }
    /* glm::mat3_cast */
    float qxx = x * x, qyy = y * y, qzz = z * z, qxz = x * z, qxy = x * y, qyz = y * z;
    float qwx = w * x, qwy = w * y, qwz = w * z;
    float R[9];
    M3(R,0,0) = 1.0f - 2.0f * (qyy + qzz);
    M3(R,0,1) = 2.0f * (qxy + qwz);
    M3(R,0,2) = 2.0f * (qxz - qwy);
    M3(R,1,0) = 2.0f * (qxy - qwz);
    M3(R,1,1) = 1.0f - 2.0f * (qxx + qzz);
    M3(R,1,2) = 2.0f * (qyz + qwx);
    M3(R,2,0) = 2.0f * (qxz + qwy);
    M3(R,2,1) = 2.0f * (qyz - qwx);
    M3(R,2,2) = 1.0f - 2.0f * (qxx + qyy);
    float S[9] = {scale[0], 0.0f, 0.0f, 0.0f, scale[1], 0.0f, 0.0f, 0.0f, scale[2]};
    float St[9], Rt[9], T[9];
    mat3_transpose(S, St);
    mat3_transpose(R, Rt);
    mat3_mul(R, S, T);   /* R*S         */
    mat3_mul(T, St, T);  /* (R*S)*S^T   */
    mat3_mul(T, Rt, V);  /* ...*R^T     */
}

typedef struct {
    int type; /* 0 unknown, 1 char,2 uchar,3 short,4 ushort,5 int,6 uint,7 float,8 double   binaryattribute.h:15-27 */
    size_t size, offset;
} orc_attr;
static const size_t k_type_size[9] = {0, 1, 1, 2, 2, 4, 4, 4, 8}; /* binaryattribute.cpp:8-18 */

/* BinaryAttribute::Read<float>: Unknown -> 0 (binaryattribute.h:57-65).  The reference asserts
 * size==sizeof(float) in debug builds; a non-float property is treated as missing here. */
static inline float attr_read_f(const orc_attr* a, const uint8_t* rec)
{
    if (a->type != 7) return 0.0f;
    float v;
    memcpy(&v, rec + a->offset, 4);
    return v;
}

typedef struct {
    char name[64];
    orc_attr attr;
} orc_prop;

typedef struct {
    size_t vertexCount, vertexSize, nprops;
    orc_prop* props;
    uint8_t* body;
} orc_ply;

static int ply_next_line(FILE* f, char* buf, size_t cap)
{
    /* ply.cpp:27-39 GetNextPlyLine: std::getline, skipping lines that START with "comment" */
    while (fgets(buf, (int)cap, f)) {
        size_t n = strlen(buf);
        if (n && buf[n - 1] == '\n') buf[n - 1] = 0; /* a trailing \r is kept, as std::getline would */
        if (strncmp(buf, "comment", 7) != 0) return 1;
    }
    return 0;
}

static int ply_type_from_token(const char* t)
{
    /* ply.cpp:214-250 */
    if (!strcmp(t, "char") || !strcmp(t, "int8")) return 1;
    if (!strcmp(t, "uchar") || !strcmp(t, "uint8")) return 2;
    if (!strcmp(t, "short") || !strcmp(t, "int16")) return 3;
    if (!strcmp(t, "ushort") || !strcmp(t, "uint16")) return 4;
    if (!strcmp(t, "int") || !strcmp(t, "int32")) return 5;
    if (!strcmp(t, "uint") || !strcmp(t, "uint32")) return 6;
    if (!strcmp(t, "float") || !strcmp(t, "float32")) return 7;
    if (!strcmp(t, "double") || !strcmp(t, "float64")) return 8;
    return 0;
}

static void ply_free(orc_ply* p)
{
    free(p->props);
    free(p->body);
    memset(p, 0, sizeof *p);
}

/* Ply::Parse / ParseHeader: ply.cpp:72-87,140-254.  returns 0 ok, <0 error */
static int ply_parse(const char* path, orc_ply* out)
{
    memset(out, 0, sizeof *out);
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    char line[1024];
    int rc = -2;
    if (!ply_next_line(f, line, sizeof line) || strcmp(line, "ply") != 0) goto done;
    if (!ply_next_line(f, line, sizeof line) || strcmp(line, "format binary_little_endian 1.0") != 0) goto done;
    if (!ply_next_line(f, line, sizeof line)) goto done;
    /* ply.cpp:176-209: ONE istringstream and three tokens for all header lines -- an extraction that finds nothing leaves
     * the token as it was (sscanf does the same with unmatched %s), so a blank line repeats the previous property line and
     * "property float" re-uses the previous name */
    char t1[64] = "", t2[64] = "", t3[64] = "";
    {
        unsigned long long cnt;
        if (sscanf(line, "%63s %63s %llu", t1, t2, &cnt) != 3 || strcmp(t1, "element") || strcmp(t2, "vertex")) goto done;
        out->vertexCount = (size_t)cnt;
    }
    size_t cap = 0;
    for (;;) {
        if (!ply_next_line(f, line, sizeof line)) goto done;
        if (!strcmp(line, "end_header")) break;
        sscanf(line, "%63s %63s %63s", t1, t2, t3);
        if (strcmp(t1, "property")) goto done; /* ply.cpp:209-213: anything else is an error */
        int ty = ply_type_from_token(t2);
        if (!ty) goto done;
        if (out->nprops == cap) {
            cap = cap ? cap * 2 : 64;
            out->props = (orc_prop*)realloc(out->props, cap * sizeof(orc_prop));
        }
        orc_prop* p = &out->props[out->nprops++];
        memset(p, 0, sizeof *p);
        memcpy(p->name, t3, sizeof p->name - 1); /* both 64 bytes, t3 is NUL-terminated within 63 */
        p->attr.type = ty;
        p->attr.size = k_type_size[ty];
        p->attr.offset = out->vertexSize; /* ply.cpp:106-112 AddProperty */
        out->vertexSize += p->attr.size;
    }
    {
        size_t bytes = out->vertexSize * out->vertexCount;
        out->body = (uint8_t*)calloc(bytes ? bytes : 1, 1);
        size_t got = fread(out->body, 1, bytes, f); /* ply.cpp:83 one bulk read; short reads are not checked */
        (void)got;
    }
    rc = 0;
done:
    fclose(f);
    if (rc) ply_free(out);
    return rc;
}

static int ply_get(const orc_ply* p, const char* key, orc_attr* a)
{
    /* unordered_map::emplace keeps the FIRST property of a given name (ply.cpp:106-112) */
    for (size_t i = 0; i < p->nprops; i++)
        if (!strcmp(p->props[i].name, key)) { *a = p->props[i].attr; return 1; }
    memset(a, 0, sizeof *a);
    return 0;
}

/* GaussianCloud::ImportPly: gaussiancloud.cpp:138-365.
 * Two-call protocol: aos==NULL -> fills n/stride/full_sh only.  returns 0 ok. */
ORC_API int orc_import_ply(const char* path, int import_full_sh, size_t* n_out, size_t* stride_out,
                           int* full_sh_out, float* aos, size_t aos_floats)
{
    orc_ply ply;
    int rc = ply_parse(path, &ply);
    if (rc) return rc;
    orc_attr ax, ay, az, adc[3], arest[45], aop, asc[3], arot[4];
    /* gaussiancloud.cpp:174-176: !Get(x) || !Get(y) || !Get(z) short-circuits -- nothing after the first missing one is looked up */
    memset(&ax, 0, sizeof ax); memset(&ay, 0, sizeof ay); memset(&az, 0, sizeof az);
    if (ply_get(&ply, "x", &ax) && ply_get(&ply, "y", &ay)) ply_get(&ply, "z", &az);
    char nm[32];
    for (int i = 0; i < 3; i++) { snprintf(nm, sizeof nm, "f_dc_%d", i); ply_get(&ply, nm, &adc[i]); }
    int full = 0;
    if (import_full_sh) { /* gaussiancloud.cpp:188-205 */
        full = 1;
        for (int i = 0; i < 45; i++) {
            snprintf(nm, sizeof nm, "f_rest_%d", i);
            if (!ply_get(&ply, nm, &arest[i])) { full = 0; break; }
        }
    }
    ply_get(&ply, "opacity", &aop);
    for (int i = 0; i < 3; i++) { snprintf(nm, sizeof nm, "scale_%d", i); ply_get(&ply, nm, &asc[i]); }
    for (int i = 0; i < 4; i++) { snprintf(nm, sizeof nm, "rot_%d", i); ply_get(&ply, nm, &arot[i]); }

    size_t nf = full ? FULL_FLOATS : BASE_FLOATS;
    *n_out = ply.vertexCount;
    *stride_out = nf * 4;
    *full_sh_out = full;
    if (!aos) { ply_free(&ply); return 0; }
    if (aos_floats < nf * ply.vertexCount) { ply_free(&ply); return -3; }

    for (size_t i = 0; i < ply.vertexCount; i++) {
        const uint8_t* v = ply.body + i * ply.vertexSize;
        float* g = aos + i * nf;
        g[OFF_POS + 0] = attr_read_f(&ax, v);
        g[OFF_POS + 1] = attr_read_f(&ay, v);
        g[OFF_POS + 2] = attr_read_f(&az, v);
        g[OFF_POS + 3] = 1.0f / (1.0f + expf(-attr_read_f(&aop, v))); /* :119-122,260 */
        if (full) { /* :262-314  f_rest_{0..14}->R, {15..29}->G, {30..44}->B, vec4 blocks per degree */
            static const int sh_off[3][4] = {{OFF_RSH0, OFF_RSH1, OFF_RSH2, OFF_RSH3},
                                             {OFF_GSH0, OFF_GSH1, OFF_GSH2, OFF_GSH3},
                                             {OFF_BSH0, OFF_BSH1, OFF_BSH2, OFF_BSH3}};
            for (int c = 0; c < 3; c++) {
                g[sh_off[c][0]] = attr_read_f(&adc[c], v);
                for (int k = 1; k < 16; k++)
                    g[sh_off[c][k >> 2] + (k & 3)] = attr_read_f(&arest[15 * c + k - 1], v);
            }
        } else { /* :316-332 */
            g[OFF_RSH0] = attr_read_f(&adc[0], v); g[OFF_RSH0 + 1] = 0.0f; g[OFF_RSH0 + 2] = 0.0f; g[OFF_RSH0 + 3] = 0.0f;
            g[OFF_GSH0] = attr_read_f(&adc[1], v); g[OFF_GSH0 + 1] = 0.0f; g[OFF_GSH0 + 2] = 0.0f; g[OFF_GSH0 + 3] = 0.0f;
            g[OFF_BSH0] = attr_read_f(&adc[2], v); g[OFF_BSH0 + 1] = 0.0f; g[OFF_BSH0 + 2] = 0.0f; g[OFF_BSH0 + 3] = 0.0f;
        }
        float scale[3] = {expf(attr_read_f(&asc[0], v)), expf(attr_read_f(&asc[1], v)), expf(attr_read_f(&asc[2], v))}; /* :335-340 */
        float rot[4] = {attr_read_f(&arot[0], v), attr_read_f(&arot[1], v), attr_read_f(&arot[2], v), attr_read_f(&arot[3], v)};
        float V[9];
        cov_from_rot_scale(rot, scale, V);
        for (int k = 0; k < 9; k++) g[OFF_COV0 + k] = V[k]; /* :349-358 V[col][row] in order */
    }
    ply_free(&ply);
    return 0;
}

/* GaussianCloud::InitDebugCloud: gaussiancloud.cpp:505-578 (fixture generator; stride 244, hasFullSH=false) */
ORC_API void orc_debug_cloud(float* aos /* 16*61 floats */)
{
    const int NUM_SPLATS = 5;
    const float AXIS_LENGTH = 1.0f;
    const float DELTA = (AXIS_LENGTH / (float)NUM_SPLATS);
    const float COV_DIAG = 0.005f;
    const float SH_C0 = 0.28209479177387814f;
    const float SH_ONE = 1.0f / (2.0f * SH_C0);
    const float SH_ZERO = -1.0f / (2.0f * SH_C0);
    memset(aos, 0, 16 * FULL_FLOATS * sizeof(float));
    for (int axis = 0; axis < 3; axis++)
        for (int i = 0; i < NUM_SPLATS; i++) {
            float* g = aos + (axis * NUM_SPLATS + i) * FULL_FLOATS;
            g[OFF_POS + axis] = i * DELTA + DELTA + (axis == 2 ? 0.0001f : 0.0f);
            g[OFF_POS + 3] = 1.0f;
            g[OFF_RSH0] = axis == 0 ? SH_ONE : SH_ZERO;
            g[OFF_GSH0] = axis == 1 ? SH_ONE : SH_ZERO;
            g[OFF_BSH0] = axis == 2 ? SH_ONE : SH_ZERO;
            g[OFF_COV0 + 0] = COV_DIAG; g[OFF_COV1 + 1] = COV_DIAG; g[OFF_COV2 + 2] = COV_DIAG;
        }
    float* g = aos + 15 * FULL_FLOATS;
    g[OFF_POS + 3] = 1.0f;
    g[OFF_RSH0] = SH_ONE; g[OFF_GSH0] = SH_ONE; g[OFF_BSH0] = SH_ONE;
    g[OFF_COV0 + 0] = COV_DIAG; g[OFF_COV1 + 1] = COV_DIAG; g[OFF_COV2 + 2] = COV_DIAG;
}

/* ------------------------------------------------------------------------------------------------ */
/* per-frame constants                                                                               */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    float view[16]; /* inverse(cameraMat)          splatrenderer.cpp:161,327 */
    float mvp[16];  /* projMat * view              splatrenderer.cpp:175     */
    float proj[16];
    float eye[3];   /* cameraMat[3].xyz            splatrenderer.cpp:328     */
    float W, H, zn, zf;
} orc_frame;

ORC_API void orc_frame_constants(const float* cameraMat, const float* projMat, const float* viewport,
                                 const float* nearFar, orc_frame* fr)
{
    mat4_inverse(cameraMat, fr->view);
    mat4_mul(projMat, fr->view, fr->mvp);
    memcpy(fr->proj, projMat, 16 * sizeof(float));
    fr->eye[0] = M(cameraMat,3,0); fr->eye[1] = M(cameraMat,3,1); fr->eye[2] = M(cameraMat,3,2);
    fr->W = viewport[2]; fr->H = viewport[3];
    fr->zn = nearFar[0]; fr->zf = nearFar[1];
}

/* splat_vert.glsl:129-151 */
static inline float srgb_to_linear(float s)
{
    if (s <= 0.04045f) return s / 12.92f;
    return powf((s + 0.055f) / 1.055f, 2.4f);
}

/* splat_vert.glsl:51-127 ComputeRadianceFromSH; g = pointer to the AoS record; nsh = 4 (base) | 16 (FULL_SH) */
static void radiance_from_sh(const float* g, int full, float vx, float vy, float vz, float* rgb)
{
    float b[16];
    float vx2 = vx * vx, vy2 = vy * vy, vz2 = vz * vz;
    b[0] = 0.28209479177387814f;
    float k1 = 0.4886025119029199f;
    b[1] = -k1 * vy;
    b[2] = k1 * vz;
    b[3] = -k1 * vx;
    const float* sh0[3] = {g + OFF_RSH0, g + OFF_GSH0, g + OFF_BSH0};
    if (full) {
        float k2 = 1.0925484305920792f, k3 = 0.31539156525252005f, k4 = 0.5462742152960396f;
        b[4] = k2 * vy * vx;
        b[5] = -k2 * vy * vz;
        b[6] = k3 * (3.0f * vz2 - 1.0f);
        b[7] = -k2 * vx * vz;
        b[8] = k4 * (vx2 - vy2);
        float k5 = 0.5900435899266435f, k6 = 2.8906114426405543f, k7 = 0.4570457994644658f;
        float k8 = 0.37317633259011546f, k9 = 1.4453057213202771f;
        b[9] = -k5 * vy * (3.0f * vx2 - vy2);
        b[10] = k6 * vy * vx * vz;
        b[11] = -k7 * vy * (5.0f * vz2 - 1.0f);
        b[12] = k8 * vz * (5.0f * vz2 - 3.0f);
        b[13] = -k7 * vx * (5.0f * vz2 - 1.0f);
        b[14] = k9 * vz * (vx2 - vy2);
        b[15] = -k5 * vx * (vx2 - 3.0f * vy2);
        const float* shN[3][3] = {{g + OFF_RSH1, g + OFF_RSH2, g + OFF_RSH3},
                                  {g + OFF_GSH1, g + OFF_GSH2, g + OFF_GSH3},
                                  {g + OFF_BSH1, g + OFF_BSH2, g + OFF_BSH3}};
        for (int c = 0; c < 3; c++) {
            float acc = b[0] * sh0[c][0];
            for (int k = 1; k < 4; k++) acc = acc + b[k] * sh0[c][k];
            for (int k = 4; k < 16; k++) acc = acc + b[k] * shN[c][(k >> 2) - 1][k & 3];
            rgb[c] = 0.5f + acc;
        }
    } else {
        for (int c = 0; c < 3; c++) {
            float acc = b[0] * sh0[c][0];
            for (int k = 1; k < 4; k++) acc = acc + b[k] * sh0[c][k];
            rgb[c] = 0.5f + acc;
        }
    }
}

/* K1 + K5 + K6 for one gaussian */
/* fs: the frame of the last Sort (pre-sort cull + depth key, K1); fr: the frame of the draw (K5/K6).  The desktop path
 * passes the same matrices to both (app.cpp:1067-1068); the VR path sorts for eye 0 and draws both eyes (app.cpp:603-607). */
static void project_one(const float* g, int full, int srgb, const orc_frame* fs, const orc_frame* fr, orc_record* o)
{
    memset(o, 0, sizeof *o);
    float x = g[OFF_POS], y = g[OFF_POS + 1], z = g[OFF_POS + 2];
    o->alpha = g[OFF_POS + 3];
    /* ---- K1: presort_compute.glsl:42-56 */
    float p[4];
    mat4_xform_point(fs->mvp, x, y, z, p);
    float depth = p[3];
    float xx = p[0] / depth, yy = p[1] / depth;
    const float CLIP = 1.5f;
    int keep = (depth > 0.0f && xx < CLIP && xx > -CLIP && yy < CLIP && yy > -CLIP);
    /* uint() of a value >= 2^32 is undefined in GLSL and the quad is far-clipped anyway: treat as culled */
    keep = keep && (depth < fs->zf);
    if (keep) {
        float q = (depth / fs->zf) * 4294967296.0f; /* float(keyMax) rounds to 2^32 */
        o->key = 0xFFFFFFFFu - (uint32_t)q;
    }
    /* ---- K5: splat_vert.glsl:153-222 */
    float t[4], p4[4];
    mat4_xform_point(fr->view, x, y, z, t);
    mat4_xform_vec4(fr->proj, t, p4);
    /* ---- K6 guard band: splat_geom.glsl:46-54 */
    float ndcx = p4[0] / p4[3], ndcy = p4[1] / p4[3], ndcz = p4[2] / p4[3];
    if (ndcz < 0.25f || ndcx > 2.0f || ndcx < -2.0f || ndcy > 2.0f || ndcy < -2.0f) keep = 0;
    if (!keep) return;

    float WIDTH = fr->W, HEIGHT = fr->H, Z_NEAR = fr->zn, Z_FAR = fr->zf;
    float SX = M(fr->proj,0,0), SY = M(fr->proj,1,1), WZ = M(fr->proj,3,2);
    float tzSq = t[2] * t[2];
    float jsx = -(SX * WIDTH) / (2.0f * t[2]);
    float jsy = -(SY * HEIGHT) / (2.0f * t[2]);
    float jtx = (SX * t[0] * WIDTH) / (2.0f * tzSq);
    float jty = (SY * t[1] * HEIGHT) / (2.0f * tzSq);
    float jtz = ((Z_FAR - Z_NEAR) * WZ) / (2.0f * tzSq);
    float J[9] = {jsx, 0.0f, 0.0f, 0.0f, jsy, 0.0f, jtx, jty, jtz}; /* columns */
    float Wm[9];
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) M3(Wm, c, r) = M(fr->view, c, r);
    float V[9];
    for (int k = 0; k < 9; k++) V[k] = g[OFF_COV0 + k];
    float JW[9], JWt[9], Vp[9];
    mat3_mul(J, Wm, JW);
    mat3_transpose(JW, JWt);
    mat3_mul(JW, V, Vp);
    mat3_mul(Vp, JWt, Vp);
    /* mat2(V_prime) + low-pass: splat_vert.glsl:191-196 ; m[col][row] */
    float m00 = M3(Vp,0,0) + 0.3f, m01 = M3(Vp,0,1), m10 = M3(Vp,1,0), m11 = M3(Vp,1,1) + 0.3f;
    /* inverseMat2: splat_geom.glsl:22-32 */
    float det = m00 * m11 - m01 * m10;
    if (!(det > 0.0f) || !(det <= 3.4028234e38f)) return; /* degenerate: the reference would emit inf/NaN; oracle skips */
    o->i00 = m11 / det;
    o->i01 = -m01 / det;
    o->i10 = -m10 / det;
    o->i11 = m00 / det;
    /* centre: splat_vert.glsl:200-203 with viewport origin 0 (X0 = Y0 = 0) */
    float gx = p4[0] / p4[3], gy = p4[1] / p4[3];
    o->px = 0.5f * (WIDTH + (gx * WIDTH) + (2.0f * 0.0f));
    o->py = 0.5f * (HEIGHT + (gy * HEIGHT) + (2.0f * 0.0f));
    /* radiance: splat_vert.glsl:206-207 */
    float dx = x - fr->eye[0], dy = y - fr->eye[1], dz = z - fr->eye[2];
    float len = sqrtf((dx * dx + dy * dy) + dz * dz);
    float vx = dx / len, vy = dy / len, vz = dz / len;
    float rgb[3];
    radiance_from_sh(g, full, vx, vy, vz, rgb);
    if (srgb) for (int c = 0; c < 3; c++) rgb[c] = srgb_to_linear(rgb[c]);
    o->r = rgb[0]; o->g = rgb[1]; o->b = rgb[2];
    o->visible = 1;
}

ORC_API void orc_project(const float* aos, size_t n, size_t stride_bytes, int full_sh, int srgb,
                         const orc_frame* fr, orc_record* out)
{
    size_t sf = stride_bytes / 4;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; i++) project_one(aos + (size_t)i * sf, full_sh, srgb, fr, fr, &out[i]);
}

/* the VR draw of the eye that was NOT sorted for (app.cpp:603-607): list membership and key from the Sort's frame,
 * everything the vertex / geometry stages compute from the draw's frame.  orc_sort() on the result gives the
 * reference's draw order for that eye (the sorted list of the other one, minus what the geometry stage discards). */
ORC_API void orc_project_shared(const float* aos, size_t n, size_t stride_bytes, int full_sh, int srgb,
                                const orc_frame* fr_sort, const orc_frame* fr_draw, orc_record* out)
{
    size_t sf = stride_bytes / 4;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; i++) project_one(aos + (size_t)i * sf, full_sh, srgb, fr_sort, fr_draw, &out[i]);
}

/* ------------------------------------------------------------------------------------------------ */
/* sort: stable LSD, 4 x 8-bit digits on the u32 key, ascending (=> back to front); ties keep index   */
/* order (the reference's atomic compaction order is nondeterministic; the oracle fixes it: q15)      */
/* ------------------------------------------------------------------------------------------------ */
ORC_API size_t orc_sort(const orc_record* rec, size_t n, uint32_t* order /* n entries */)
{
    uint32_t* ka = (uint32_t*)malloc(n * 4 + 4), *kb = (uint32_t*)malloc(n * 4 + 4);
    uint32_t* va = (uint32_t*)malloc(n * 4 + 4), *vb = (uint32_t*)malloc(n * 4 + 4);
    size_t V = 0;
    for (size_t i = 0; i < n; i++)
        if (rec[i].visible) { ka[V] = rec[i].key; va[V] = (uint32_t)i; V++; }
    for (int pass = 0; pass < 4; pass++) {
        size_t hist[256] = {0};
        int shift = 8 * pass;
        for (size_t i = 0; i < V; i++) hist[(ka[i] >> shift) & 255]++;
        size_t sum = 0;
        for (int d = 0; d < 256; d++) { size_t c = hist[d]; hist[d] = sum; sum += c; }
        for (size_t i = 0; i < V; i++) {
            size_t dst = hist[(ka[i] >> shift) & 255]++;
            kb[dst] = ka[i]; vb[dst] = va[i];
        }
        uint32_t* t = ka; ka = kb; kb = t;
        t = va; va = vb; vb = t;
    }
    memcpy(order, va, V * 4);
    free(ka); free(kb); free(va); free(vb);
    return V;
}

/* ------------------------------------------------------------------------------------------------ */
/* composite: K7 + K8, strict back-to-front "over" in float32 (or binary16 storage)                   */
/* ------------------------------------------------------------------------------------------------ */
static inline float round_half(float v) { return (float)(_Float16)v; }

static inline void blend_pixel(const orc_record* s, float fx, float fy, float* px, int fp16)
{
    float dx = fx - s->px, dy = fy - s->py;            /* splat_frag.glsl:20 */
    /* dot(d, cov2Dinv * d), cov2Dinv = mat2(i00,i01 | i10,i11) columns  :23-25 */
    float mx = s->i00 * dx + s->i10 * dy;
    float my = s->i01 * dx + s->i11 * dy;
    float g = expf(-0.5f * (dx * mx + dy * my));
    float w = s->alpha * g;
    if (w <= (1.0f / 256.0f)) return;                 /* discard :38-41 */
    float sr = s->alpha * g * s->r, sg = s->alpha * g * s->g, sb = s->alpha * g * s->b; /* :27 */
    float om = 1.0f - w;                              /* GL_ONE, GL_ONE_MINUS_SRC_ALPHA  app.cpp:153-156 */
    px[0] = sr + om * px[0];
    px[1] = sg + om * px[1];
    px[2] = sb + om * px[2];
    px[3] = w + om * px[3];
    if (fp16) { px[0] = round_half(px[0]); px[1] = round_half(px[1]); px[2] = round_half(px[2]); px[3] = round_half(px[3]); }
}

/* brute force: every visible splat against every pixel; for small cases and for validating the binned path */
ORC_API void orc_composite_bruteforce(const orc_record* rec, const uint32_t* order, size_t V, int W, int H,
                                      int fp16, float* fb)
{
    for (size_t i = 0; i < (size_t)W * H; i++) { fb[4 * i] = 0.f; fb[4 * i + 1] = 0.f; fb[4 * i + 2] = 0.f; fb[4 * i + 3] = 1.f; } /* app.cpp:158-160 */
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; y++)
        for (size_t k = 0; k < V; k++) {
            const orc_record* s = &rec[order[k]];
            for (int x = 0; x < W; x++) blend_pixel(s, (float)x + 0.5f, (float)y + 0.5f, fb + 4 * ((size_t)y * W + x), fp16);
        }
}

/* binned: per 16x16 bin a list of splats (global order preserved), each restricted to the axis-aligned
 * box of its 3.5-sigma ellipse (inside the reference's quad, splat_geom.glsl:56-87) + 1 px.  Because
 * alpha<=1 the discard contour lies at Mahalanobis radius <= sqrt(2 ln 256) = 3.33 < 3.5, so the box
 * never removes a contributing pixel. */
#define BIN 16
ORC_API void orc_composite(const orc_record* rec, const uint32_t* order, size_t V, int W, int H, int fp16, float* fb)
{
    int bw = (W + BIN - 1) / BIN, bh = (H + BIN - 1) / BIN;
    size_t nb = (size_t)bw * bh;
    size_t* cnt = (size_t*)calloc(nb + 1, sizeof(size_t));
    int* box = (int*)malloc(V * 4 * sizeof(int));
    for (size_t k = 0; k < V; k++) {
        const orc_record* s = &rec[order[k]];
        float det = s->i00 * s->i11 - s->i01 * s->i10;
        float a = s->i11 / det, c = s->i00 / det; /* back to covariance diagonal */
        float ex = 3.5f * sqrtf(fabsf(a)) + 1.0f, ey = 3.5f * sqrtf(fabsf(c)) + 1.0f;
        float x0f = floorf(s->px - ex), x1f = ceilf(s->px + ex), y0f = floorf(s->py - ey), y1f = ceilf(s->py + ey);
        int x0 = x0f < 0.f ? 0 : (x0f > (float)W ? W : (int)x0f), x1 = x1f < 0.f ? 0 : (x1f > (float)W ? W : (int)x1f);
        int y0 = y0f < 0.f ? 0 : (y0f > (float)H ? H : (int)y0f), y1 = y1f < 0.f ? 0 : (y1f > (float)H ? H : (int)y1f);
        if (!(ex == ex) || !(ey == ey)) { x0 = 0; x1 = W; y0 = 0; y1 = H; } /* NaN extent: cover everything */
        box[4 * k] = x0; box[4 * k + 1] = x1; box[4 * k + 2] = y0; box[4 * k + 3] = y1;
        if (x1 <= x0 || y1 <= y0) continue;
        for (int by = y0 / BIN; by <= (y1 - 1) / BIN; by++)
            for (int bx = x0 / BIN; bx <= (x1 - 1) / BIN; bx++) cnt[(size_t)by * bw + bx + 1]++;
    }
    for (size_t b = 0; b < nb; b++) cnt[b + 1] += cnt[b];
    uint32_t* list = (uint32_t*)malloc((cnt[nb] ? cnt[nb] : 1) * sizeof(uint32_t));
    size_t* cur = (size_t*)malloc(nb * sizeof(size_t));
    memcpy(cur, cnt, nb * sizeof(size_t));
    for (size_t k = 0; k < V; k++) {
        int x0 = box[4 * k], x1 = box[4 * k + 1], y0 = box[4 * k + 2], y1 = box[4 * k + 3];
        if (x1 <= x0 || y1 <= y0) continue;
        for (int by = y0 / BIN; by <= (y1 - 1) / BIN; by++)
            for (int bx = x0 / BIN; bx <= (x1 - 1) / BIN; bx++) list[cur[(size_t)by * bw + bx]++] = (uint32_t)k;
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (long long b = 0; b < (long long)nb; b++) {
        int bx = (int)(b % bw), by = (int)(b / bw);
        int tx0 = bx * BIN, ty0 = by * BIN, tx1 = tx0 + BIN > W ? W : tx0 + BIN, ty1 = ty0 + BIN > H ? H : ty0 + BIN;
        for (int y = ty0; y < ty1; y++)
            for (int x = tx0; x < tx1; x++) { float* p = fb + 4 * ((size_t)y * W + x); p[0] = 0.f; p[1] = 0.f; p[2] = 0.f; p[3] = 1.f; }
        for (size_t j = cnt[b]; j < cnt[b + 1]; j++) {
            size_t k = list[j];
            const orc_record* s = &rec[order[k]];
            int x0 = box[4 * k] > tx0 ? box[4 * k] : tx0, x1 = box[4 * k + 1] < tx1 ? box[4 * k + 1] : tx1;
            int y0 = box[4 * k + 2] > ty0 ? box[4 * k + 2] : ty0, y1 = box[4 * k + 3] < ty1 ? box[4 * k + 3] : ty1;
            for (int y = y0; y < y1; y++)
                for (int x = x0; x < x1; x++) blend_pixel(s, (float)x + 0.5f, (float)y + 0.5f, fb + 4 * ((size_t)y * W + x), fp16);
        }
    }
    free(cnt); free(box); free(list); free(cur);
}

/* one whole frame = SplatRenderer::Sort + SplatRenderer::Render into a cleared float target.
 * stage_ms (optional, 3 floats): project, sort, composite wall ms. returns V (sortCount). */
ORC_API size_t orc_render(const float* aos, size_t n, size_t stride_bytes, int full_sh, int srgb,
                          const float* cameraMat, const float* projMat, const float* viewport, const float* nearFar,
                          int fp16, float* fb, orc_record* rec_out, uint32_t* order_out, double* stage_ms)
{
    orc_frame fr;
    orc_frame_constants(cameraMat, projMat, viewport, nearFar, &fr);
    orc_record* rec = rec_out ? rec_out : (orc_record*)malloc((n ? n : 1) * sizeof(orc_record));
    uint32_t* order = order_out ? order_out : (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
    double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
#ifdef _OPENMP
    t0 = omp_get_wtime();
#endif
    orc_project(aos, n, stride_bytes, full_sh, srgb, &fr, rec);
#ifdef _OPENMP
    t1 = omp_get_wtime();
#endif
    size_t V = orc_sort(rec, n, order);
#ifdef _OPENMP
    t2 = omp_get_wtime();
#endif
    orc_composite(rec, order, V, (int)fr.W, (int)fr.H, fp16, fb);
#ifdef _OPENMP
    t3 = omp_get_wtime();
#endif
    if (stage_ms) { stage_ms[0] = (t1 - t0) * 1e3; stage_ms[1] = (t2 - t1) * 1e3; stage_ms[2] = (t3 - t2) * 1e3; }
    if (!rec_out) free(rec);
    if (!order_out) free(order);
    return V;
}

/* stand-alone sort oracle for the u64 (tile|depth) key sort micro-benchmark: stable ascending */
ORC_API void orc_sort_pairs_u64(const uint64_t* keys, const uint32_t* vals, size_t n, int begin_bit, int end_bit,
                                uint64_t* keys_out, uint32_t* vals_out)
{
    uint64_t* ka = (uint64_t*)malloc(n * 8 + 8), *kb = (uint64_t*)malloc(n * 8 + 8);
    uint32_t* va = (uint32_t*)malloc(n * 4 + 4), *vb = (uint32_t*)malloc(n * 4 + 4);
    memcpy(ka, keys, n * 8); memcpy(va, vals, n * 4);
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        int bits = end_bit - shift < 8 ? end_bit - shift : 8;
        uint64_t mask = (1ull << bits) - 1;
        size_t hist[256] = {0};
        for (size_t i = 0; i < n; i++) hist[(ka[i] >> shift) & mask]++;
        size_t sum = 0;
        for (int d = 0; d < 256; d++) { size_t c = hist[d]; hist[d] = sum; sum += c; }
        for (size_t i = 0; i < n; i++) { size_t dst = hist[(ka[i] >> shift) & mask]++; kb[dst] = ka[i]; vb[dst] = va[i]; }
        uint64_t* t = ka; ka = kb; kb = t;
        uint32_t* tv = va; va = vb; vb = tv;
    }
    memcpy(keys_out, ka, n * 8); memcpy(vals_out, va, n * 4);
    free(ka); free(kb); free(va); free(vb);
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
ORC_API void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
ORC_API size_t orc_record_size(void) { return sizeof(orc_record); }
ORC_API size_t orc_frame_size(void) { return sizeof(orc_frame); }
