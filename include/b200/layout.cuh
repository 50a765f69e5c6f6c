// Internal: how point state is laid out in HBM while a step is in flight.
//
// User-visible state (Solution::d_X, d_old_v, …) is array-of-structs in the
// ORIGINAL cell order, because model code indexes it by cell id between steps
// (SURVEY.md 8b). For the pairwise sweep the solver keeps a second copy in
// CUBE ORDER (sorted by cube id, ascending original id inside a cube), split
// into planes so that the hot loop only touches what it needs:
//
//   pos4[k] = { x, y, z, bits(original id) }                16 B per cell
//       the only data the 27-cube candidate scan reads; staged to shared
//       memory with 1-D bulk copies (cube-ordered spans are contiguous).
//   aux[k]  = { extra lanes of Pt ..., old_v.x, old_v.y, old_v.z, pad }
//       padded to a multiple of 4 floats (16 B for float3/float4, 32 B = one
//       sector for Po_cell and the 7-float branching Cell); read only for
//       candidates that pass the distance test.
//   cube[k] = cube id of slot k                               4 B per cell
#pragma once

#include <cuda_runtime.h>

namespace yb {

template<typename Pt>
struct Layout {
    static_assert(sizeof(Pt) % sizeof(float) == 0 && sizeof(Pt) >= 12,
        "point types are structs of floats starting with x, y, z");
    static constexpr int lanes = sizeof(Pt) / sizeof(float);
    static constexpr int extras = lanes - 3;
    // extras + 3 velocity lanes, rounded up to whole float4s
    static constexpr int aux_lanes = ((extras + 3 + 3) / 4) * 4;
    static constexpr int aux_vec4 = aux_lanes / 4;
    static constexpr int v_lane = extras;  // first velocity lane inside aux
};

// Lane views. All indices are compile-time constants after unrolling, so the
// structs stay in registers.
template<typename Pt>
__device__ __host__ __forceinline__ float& lane(Pt& X, int k)
{
    return reinterpret_cast<float*>(&X)[k];
}

template<typename Pt>
__device__ __host__ __forceinline__ const float& lane(const Pt& X, int k)
{
    return reinterpret_cast<const float*>(&X)[k];
}

// Strided scalar loads/stores of one AoS element: Pt is only 4-byte aligned
// in general (float3, MAKE_PT types), so no vector access here.
template<typename Pt>
__device__ __forceinline__ Pt load_pt(const Pt* __restrict__ base, int i)
{
    Pt X;
    const float* src = reinterpret_cast<const float*>(base + i);
#pragma unroll
    for (int k = 0; k < Layout<Pt>::lanes; k++) lane(X, k) = __ldg(src + k);
    return X;
}

template<typename Pt>
__device__ __forceinline__ Pt load_pt_rw(const Pt* base, int i)
{
    Pt X;
    const float* src = reinterpret_cast<const float*>(base + i);
#pragma unroll
    for (int k = 0; k < Layout<Pt>::lanes; k++) lane(X, k) = src[k];
    return X;
}

template<typename Pt>
__device__ __forceinline__ void store_pt(Pt* base, int i, const Pt& X)
{
    float* dst = reinterpret_cast<float*>(base + i);
#pragma unroll
    for (int k = 0; k < Layout<Pt>::lanes; k++) dst[k] = lane(X, k);
}

__host__ __device__ constexpr int ceil_div(int a, int b)
{
    return (a + b - 1) / b;
}

}  // namespace yb
