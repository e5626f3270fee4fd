/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/easu_ref.c header).
 *
 * A minimal OpenCL-C-on-the-host shim: just enough of the OpenCL C language (vector types with component access,
 * as_type / convert_type / vloadN / vstoreN, the work-item functions and the handful of math built-ins) to compile
 * the reference's OWN kernel source — /root/reference/LiveVisionKit/Functions/OpenCL/Sources/FSR.cl — with g++ and
 * run its __kernel functions work-item by work-item on the CPU.  Nothing in this file restates FSR.cl: the kernel
 * text is read from the reference where it lies (oracle/ref_build/build_ref.sh pipes it into the compiler) and the
 * result (oracle/_ref/libfsrcl_ref_*.so) is the reference's arithmetic, not ours.  It pins oracle/easu_ref.c (the
 * restatement) and, through it, the CUDA kernels.
 *
 * Semantics of the built-ins follow the OpenCL C 1.2 specification:
 *   max(x,y) = (x < y) ? y : x, min(x,y) = (y < x) ? y : x (6.12.4), fmax/fmin = IEEE maxNum/minNum (6.12.2),
 *   clamp(x,lo,hi) = min(max(x,lo),hi), convert_intN / convert_ucharN (no suffix) and _rtz round toward zero for
 *   float -> integer (6.2.3.3), native_recip is implementation-defined: 1.0f / x here (correctly rounded).
 * Multiply-add contraction is left to the compiler flags of the build (OpenCL C allows it by default):
 *   libfsrcl_ref_strict.so   -ffp-contract=off   no a*b+c is fused
 *   libfsrcl_ref_contract.so -ffp-contract=fast -mfma   g++ fuses wherever it can after inlining
 * The spread between the two is the tolerance an unpinned OpenCL device has (tests/test_fsr_ref_cpu.py reports it).
 */
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

typedef uint8_t uchar;
typedef uint32_t uint;

#define __kernel
#define __global

/* ---- vector types ---------------------------------------------------------------------------------------------- */
#define CL_VEC2(T, N)                                                                                                  \
    struct N                                                                                                           \
    {                                                                                                                  \
        T x, y;                                                                                                        \
        N() = default;                                                                                                 \
        explicit N(T a) : x(a), y(a) {}                                                                                \
        N(T a, T b) : x(a), y(b) {}                                                                                    \
    };
#define CL_VEC3(T, N)                                                                                                  \
    struct N                                                                                                           \
    {                                                                                                                  \
        T x, y, z;                                                                                                     \
        N() = default;                                                                                                 \
        explicit N(T a) : x(a), y(a), z(a) {}                                                                          \
        N(T a, T b, T c) : x(a), y(b), z(c) {}                                                                         \
    };
#define CL_VEC4(T, N, N2, N3)                                                                                          \
    struct N                                                                                                           \
    {                                                                                                                  \
        union                                                                                                          \
        {                                                                                                              \
            struct { T x, y, z, w; };                                                                                  \
            N2 xy;                                                                                                     \
            N3 xyz;                                                                                                    \
        };                                                                                                             \
        N() = default;                                                                                                 \
        explicit N(T a) : x(a), y(a), z(a), w(a) {}                                                                    \
        N(T a, T b, T c, T d) : x(a), y(b), z(c), w(d) {}                                                              \
    };
#define CL_VEC8(T, N)                                                                                                  \
    struct N                                                                                                           \
    {                                                                                                                  \
        union                                                                                                          \
        {                                                                                                              \
            T s[8];                                                                                                    \
            struct { T s0, s1, s2, s3, s4, s5, s6, s7; };                                                              \
        };                                                                                                             \
    };
#define CL_VEC16(T, N)                                                                                                 \
    struct N                                                                                                           \
    {                                                                                                                  \
        union                                                                                                          \
        {                                                                                                              \
            T s[16];                                                                                                   \
            struct { T s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, sA, sB, sC, sD, sE, sF; };                              \
        };                                                                                                             \
    };

CL_VEC2(float, float2) CL_VEC3(float, float3) CL_VEC4(float, float4, float2, float3) CL_VEC8(float, float8) CL_VEC16(float, float16)
CL_VEC2(int, int2) CL_VEC3(int, int3) CL_VEC4(int, int4, int2, int3)
CL_VEC2(uint, uint2) CL_VEC3(uint, uint3) CL_VEC4(uint, uint4, uint2, uint3)
CL_VEC2(uchar, uchar2) CL_VEC3(uchar, uchar3) CL_VEC4(uchar, uchar4, uchar2, uchar3) CL_VEC8(uchar, uchar8) CL_VEC16(uchar, uchar16)

/* ---- component-wise operators (generated for the N-ary struct forms; arrays for 8/16) ----------------------------- */
#define CL_OPS2(V, T, OP)                                                                                              \
    static inline V operator OP(V a, V b) { return V(a.x OP b.x, a.y OP b.y); }                                        \
    static inline V operator OP(V a, T b) { return V(a.x OP b, a.y OP b); }                                            \
    static inline V operator OP(T a, V b) { return V(a OP b.x, a OP b.y); }                                            \
    static inline V& operator OP##=(V& a, V b) { a = a OP b; return a; }                                               \
    static inline V& operator OP##=(V& a, T b) { a = a OP b; return a; }
#define CL_OPS3(V, T, OP)                                                                                              \
    static inline V operator OP(V a, V b) { return V(a.x OP b.x, a.y OP b.y, a.z OP b.z); }                            \
    static inline V operator OP(V a, T b) { return V(a.x OP b, a.y OP b, a.z OP b); }                                  \
    static inline V operator OP(T a, V b) { return V(a OP b.x, a OP b.y, a OP b.z); }                                  \
    static inline V& operator OP##=(V& a, V b) { a = a OP b; return a; }                                               \
    static inline V& operator OP##=(V& a, T b) { a = a OP b; return a; }
#define CL_OPS4(V, T, OP)                                                                                              \
    static inline V operator OP(V a, V b) { return V(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w); }                \
    static inline V operator OP(V a, T b) { return V(a.x OP b, a.y OP b, a.z OP b, a.w OP b); }                        \
    static inline V operator OP(T a, V b) { return V(a OP b.x, a OP b.y, a OP b.z, a OP b.w); }                        \
    static inline V& operator OP##=(V& a, V b) { a = a OP b; return a; }                                               \
    static inline V& operator OP##=(V& a, T b) { a = a OP b; return a; }
#define CL_ARITH(M, V, T) M(V, T, +) M(V, T, -) M(V, T, *)
CL_ARITH(CL_OPS2, float2, float) CL_ARITH(CL_OPS3, float3, float) CL_ARITH(CL_OPS4, float4, float)
CL_ARITH(CL_OPS2, int2, int) CL_ARITH(CL_OPS4, int4, int)
CL_ARITH(CL_OPS2, uint2, uint) CL_ARITH(CL_OPS3, uint3, uint) CL_ARITH(CL_OPS4, uint4, uint)
static inline float2 operator-(float2 a) { return float2(-a.x, -a.y); }
static inline float3 operator-(float3 a) { return float3(-a.x, -a.y, -a.z); }
static inline float4 operator-(float4 a) { return float4(-a.x, -a.y, -a.z, -a.w); }
static inline uint2 operator>>(uint2 a, int s) { return uint2(a.x >> s, a.y >> s); }
static inline uint3 operator>>(uint3 a, int s) { return uint3(a.x >> s, a.y >> s, a.z >> s); }
static inline uint4 operator>>(uint4 a, int s) { return uint4(a.x >> s, a.y >> s, a.z >> s, a.w >> s); }
static inline float8 operator*(float8 a, float b) { float8 r; for (int i = 0; i < 8; i++) r.s[i] = a.s[i] * b; return r; }
static inline float16 operator*(float16 a, float b) { float16 r; for (int i = 0; i < 16; i++) r.s[i] = a.s[i] * b; return r; }

/* ---- as_type: bit reinterpretation (6.2.4.2) --------------------------------------------------------------------- */
template <class To, class From> static inline To cl_bitcast(From f)
{
    static_assert(sizeof(To) == sizeof(From), "as_type needs equal sizes");
    To t;
    std::memcpy(&t, &f, sizeof(To));
    return t;
}
static inline float as_float(uint a) { return cl_bitcast<float>(a); }
static inline uint as_uint(float a) { return cl_bitcast<uint>(a); }
static inline float2 as_float2(uint2 a) { return float2(as_float(a.x), as_float(a.y)); }
static inline float3 as_float3(uint3 a) { return float3(as_float(a.x), as_float(a.y), as_float(a.z)); }
static inline float4 as_float4(uint4 a) { return float4(as_float(a.x), as_float(a.y), as_float(a.z), as_float(a.w)); }
static inline uint2 as_uint2(float2 a) { return uint2(as_uint(a.x), as_uint(a.y)); }
static inline uint3 as_uint3(float3 a) { return uint3(as_uint(a.x), as_uint(a.y), as_uint(a.z)); }
static inline uint4 as_uint4(float4 a) { return uint4(as_uint(a.x), as_uint(a.y), as_uint(a.z), as_uint(a.w)); }
static inline float2 as_float2(uchar8 a) { float2 r; std::memcpy(&r, a.s, 8); return r; }

/* ---- convert_type (6.2.3): float -> integer truncates; integer -> float is exact for these ranges ----------------- */
static inline int2 convert_int2(uint2 a) { return int2((int)a.x, (int)a.y); }
static inline int2 convert_int2_rtz(float2 a) { return int2((int)a.x, (int)a.y); }
static inline float2 convert_float2(int2 a) { return float2((float)a.x, (float)a.y); }
static inline float3 convert_float3(uchar3 a) { return float3((float)a.x, (float)a.y, (float)a.z); }
static inline float8 convert_float8(uchar8 a) { float8 r; for (int i = 0; i < 8; i++) r.s[i] = (float)a.s[i]; return r; }
static inline float16 convert_float16(uchar16 a) { float16 r; for (int i = 0; i < 16; i++) r.s[i] = (float)a.s[i]; return r; }
static inline uchar3 convert_uchar3(float3 a) { return uchar3((uchar)(int)a.x, (uchar)(int)a.y, (uchar)(int)a.z); }

/* ---- vloadN / vstoreN (6.12.7) ------------------------------------------------------------------------------------ */
static inline uchar3 vload3(size_t off, const uchar* p) { p += 3 * off; return uchar3(p[0], p[1], p[2]); }
static inline uchar8 vload8(size_t off, const uchar* p) { uchar8 r; std::memcpy(r.s, p + 8 * off, 8); return r; }
static inline uchar16 vload16(size_t off, const uchar* p) { uchar16 r; std::memcpy(r.s, p + 16 * off, 16); return r; }
static inline void vstore3(uchar3 v, size_t off, uchar* p) { p += 3 * off; p[0] = v.x; p[1] = v.y; p[2] = v.z; }

/* ---- math / common built-ins ------------------------------------------------------------------------------------- */
static inline float max(float a, float b) { return (a < b) ? b : a; }
static inline float min(float a, float b) { return (b < a) ? b : a; }
static inline float3 max(float3 a, float3 b) { return float3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
static inline float3 min(float3 a, float3 b) { return float3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
static inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
static inline float fmax(float a, float b) { return std::fmax(a, b); }
static inline float fmin(float a, float b) { return std::fmin(a, b); }
static inline float fabs(float a) { return std::fabs(a); }
static inline float2 floor(float2 a) { return float2(std::floor(a.x), std::floor(a.y)); }
static inline float native_recip(float a) { return 1.0f / a; }

/* ---- work-item functions: set by the NDRange driver below ---------------------------------------------------------- */
static thread_local size_t cl_local_id[2], cl_group_id[2];
static inline size_t get_local_id(uint d) { return cl_local_id[d]; }
static inline size_t get_group_id(uint d) { return cl_group_id[d]; }

/* Runs `body` for every work-item of a 2-D NDRange with 8x8 work-groups (what lvk::ocl::optimal_groups picks for 2-D
 * buffers, Kernels.cpp:62-68: global = ceil(size / 8) * 8 in both dimensions, so the padding items run too). */
template <class F> static inline void cl_run_groups(int gy0, int gy1, int groups_x, F body)
{
    for (int gy = gy0; gy < gy1; gy++)
        for (int gx = 0; gx < groups_x; gx++)
            for (int ly = 0; ly < 8; ly++)
                for (int lx = 0; lx < 8; lx++)
                {
                    cl_local_id[0] = (size_t)lx; cl_local_id[1] = (size_t)ly;
                    cl_group_id[0] = (size_t)gx; cl_group_id[1] = (size_t)gy;
                    body();
                }
}
