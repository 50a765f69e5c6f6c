// Point types and their vector-space arithmetic.
//
// A point type Pt is a plain struct of floats whose first three members are
// x, y, z (position); any further members (w, theta, phi, u, v, …) are extra
// degrees of freedom that are integrated alongside. float3, float4, Po_cell
// and everything made with MAKE_PT(Name, extra...) qualify. The B200 solver
// relies on exactly this layout: it views a Pt as sizeof(Pt)/4 float lanes
// when it moves state between the user's AoS arrays and its own cube-ordered
// SoA planes.
//
// Arithmetic contract (same observable rounding as the reference,
// include/dtypes.cuh:11-217):
//   a += b, a *= s        lane-wise, one rounding per lane
//   a -= b                a += (-1 * b)   (the negation is exact)
//   a /= s                a *= float(1.0 / double(s))  -- reciprocal in double
//   + - * /               copies built from the compound forms above
#pragma once

#include <type_traits>
#include <cuda_runtime.h>


// Opt-in trait: only types marked here get the generic operators below.
template<typename Pt>
struct Is_vector : public std::false_type {};

template<>
struct Is_vector<float3> : public std::true_type {};

template<>
struct Is_vector<float4> : public std::true_type {};


// ---- compound assignment for the two CUDA built-ins -----------------------
__device__ __host__ inline float3 operator+=(float3& a, const float3& b)
{
    a.x += b.x, a.y += b.y, a.z += b.z;
    return a;
}

__device__ __host__ inline float3 operator*=(float3& a, const float b)
{
    a.x *= b, a.y *= b, a.z *= b;
    return a;
}

__device__ __host__ inline float4 operator+=(float4& a, const float4& b)
{
    a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
    return a;
}

__device__ __host__ inline float4 operator*=(float4& a, const float b)
{
    a.x *= b, a.y *= b, a.z *= b, a.w *= b;
    return a;
}


// ---- MAKE_PT ---------------------------------------------------------------
// YB_EACH(F, a, b, c) expands to F(a) F(b) F(c) for up to 80 arguments. It
// works by re-scanning: every YB_EVAL level multiplies the number of macro
// expansion passes, and YB_EACH_STEP defers its own re-invocation by one pass.
#define YB_PARENS ()
#define YB_EVAL(...) YB_EVAL4(YB_EVAL4(YB_EVAL4(__VA_ARGS__)))
#define YB_EVAL4(...) YB_EVAL3(YB_EVAL3(YB_EVAL3(__VA_ARGS__)))
#define YB_EVAL3(...) YB_EVAL2(YB_EVAL2(YB_EVAL2(__VA_ARGS__)))
#define YB_EVAL2(...) YB_EVAL1(YB_EVAL1(YB_EVAL1(__VA_ARGS__)))
#define YB_EVAL1(...) __VA_ARGS__

#define YB_SECOND(a, b, ...) b
#define YB_IS_END_PROBE(...) YB_SECOND(__VA_ARGS__, 0, ~)
#define YB_IS_END(x) YB_IS_END_PROBE(YB_END_MARK_##x)
#define YB_END_MARK_YB_STOP ~, 1

#define YB_CAT(a, b) YB_CAT_(a, b)
#define YB_CAT_(a, b) a##b
#define YB_EACH_STEP(F, x, ...) \
    YB_CAT(YB_EACH_BRANCH_, YB_IS_END(x))(F, x, __VA_ARGS__)
#define YB_EACH_BRANCH_1(F, x, ...)
#define YB_EACH_BRANCH_0(F, x, ...) \
    F(x) YB_EACH_AGAIN YB_PARENS(F, __VA_ARGS__)
#define YB_EACH_AGAIN() YB_EACH_STEP
#define YB_EACH(F, ...) YB_EVAL(YB_EACH_STEP(F, __VA_ARGS__, YB_STOP, ~))

#define YB_LANE_ADD(member) a.member += b.member;
#define YB_LANE_SCALE(member) a.member *= b;

// MAKE_PT(Cell, theta, phi) declares
//     struct Cell { float x, y, z, theta, phi; };
// with lane-wise += and *= found by argument-dependent lookup, and registers
// the type with Is_vector so that the generic operators apply.
#define MAKE_PT(Pt, ...)                                                 \
    struct Pt {                                                          \
        float x, y, z, __VA_ARGS__;                                      \
                                                                         \
        friend __device__ __host__ inline Pt operator+=(                 \
            Pt& a, const Pt& b)                                          \
        {                                                                \
            YB_EACH(YB_LANE_ADD, x, y, z, __VA_ARGS__)                   \
            return a;                                                    \
        }                                                                \
        friend __device__ __host__ inline Pt operator*=(                 \
            Pt& a, const float b)                                        \
        {                                                                \
            YB_EACH(YB_LANE_SCALE, x, y, z, __VA_ARGS__)                 \
            return a;                                                    \
        }                                                                \
    };                                                                   \
                                                                         \
    template<>                                                           \
    struct Is_vector<Pt> : public std::true_type {}

// Cell with a polarity in spherical coordinates, see polarity.cuh.
MAKE_PT(Po_cell, theta, phi);


// ---- derived operators -----------------------------------------------------
#define YB_VECTOR_OP \
    __device__ __host__ inline \
        typename std::enable_if<Is_vector<Pt>::value, Pt>::type

template<typename Pt>
YB_VECTOR_OP operator*(const Pt& a, const float b)
{
    Pt scaled = a;
    scaled *= b;
    return scaled;
}

template<typename Pt>
YB_VECTOR_OP operator*(const float b, const Pt& a)
{
    Pt scaled = a;
    scaled *= b;
    return scaled;
}

template<typename Pt>
YB_VECTOR_OP operator+(const Pt& a, const Pt& b)
{
    Pt total = a;
    total += b;
    return total;
}

// Subtraction is "add the negated operand"; -1 * b is exact in IEEE
// arithmetic, so this rounds like a.x - b.x.
template<typename Pt>
YB_VECTOR_OP operator-=(Pt& a, const Pt& b)
{
    a += -1 * b;
    return a;
}

template<typename Pt>
YB_VECTOR_OP operator-(const Pt& a, const Pt& b)
{
    Pt difference = a;
    difference -= b;
    return difference;
}

template<typename Pt>
YB_VECTOR_OP operator-(const Pt& a)
{
    return -1 * a;
}

// Division multiplies by the reciprocal. The reciprocal is formed in double
// precision and narrowed once when it is handed to *= (whose scalar is float).
template<typename Pt>
YB_VECTOR_OP operator/=(Pt& a, const float b)
{
    a *= 1. / b;
    return a;
}

template<typename Pt>
YB_VECTOR_OP operator/(const Pt& a, const float b)
{
    Pt quotient = a;
    quotient /= b;
    return quotient;
}
