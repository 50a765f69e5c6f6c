// Extension: seeded initial conditions generated and relaxed on the device.
//
// The reference's generators (inits.cuh:34-125) draw every coordinate from
// the host's rand(), seeded from std::random_device, copy all n_max points to
// the device, and relaxed_sphere then integrates up to 3000 steps. Results
// are not repeatable and at 10^6..10^8 cells the host loop and the copies
// dominate the set-up. The functions below place the same distributions
// (same radius and volume formulas) with a counter-based generator --
// Philox4_32_10 keyed by (seed, cell index) -- so a tissue depends only on
// the seed, not on launch geometry or on how many cells are generated at
// once, and relax it with take_step<relu_force> without leaving the device.
//
//   seeded_sphere(dist_to_nb, points, seed, n_0 = 0)
//   seeded_cuboid(dist_to_nb, minimum, maximum, points, seed, n_0 = 0)
//   relaxed_seeded_sphere(dist_to_nb, points, seed, n_0 = 0, relax_steps = -1)
//   relaxed_seeded_cuboid(dist_to_nb, minimum, maximum, points, seed,
//                         n_0 = 0, relax_steps = -1)
//
// Like the reference's generators they fill cells n_0 .. *h_n - 1, touch only
// x, y, z, and leave host and device copies in agreement.
#pragma once

#include <assert.h>
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <math.h>

#include "../cudebug.cuh"
#include "grid_build.cuh"

template<typename Pt, template<typename> class Solver>
class Solution;

template<typename Pt>
__device__ Pt relu_force(Pt Xi, Pt r, float dist, int i, int j);

namespace yb {

// Three uniforms in (0, 1] for cell i of the tissue `seed`.
__device__ __forceinline__ float4 cell_uniforms(unsigned long long seed, int i)
{
    curandStatePhilox4_32_10_t state;
    curand_init(seed, static_cast<unsigned long long>(i), 0, &state);
    return curand_uniform4(&state);
}

template<typename Pt>
__global__ void __launch_bounds__(256) seed_ball(
    Pt* d_X, int n_0, int n, float r_max, unsigned long long seed)
{
    for (int i = n_0 + blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += gridDim.x * blockDim.x) {
        const float4 u = cell_uniforms(seed, i);
        const float r = r_max * cbrtf(u.x);
        const float cos_theta = 2.f * u.y - 1.f;
        const float sin_theta = sqrtf(fmaxf(1.f - cos_theta * cos_theta, 0.f));
        float sin_phi, cos_phi;
        sincospif(2.f * u.z, &sin_phi, &cos_phi);
        d_X[i].x = r * sin_theta * cos_phi;
        d_X[i].y = r * sin_theta * sin_phi;
        d_X[i].z = r * cos_theta;
    }
}

template<typename Pt>
__global__ void __launch_bounds__(256) seed_box(Pt* d_X, int n_0, int n,
    float3 minimum, float3 dimension, unsigned long long seed)
{
    for (int i = n_0 + blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += gridDim.x * blockDim.x) {
        const float4 u = cell_uniforms(seed, i);
        d_X[i].x = minimum.x + dimension.x * u.x;
        d_X[i].y = minimum.y + dimension.y * u.y;
        d_X[i].z = minimum.z + dimension.z * u.z;
    }
}

template<typename Pt>
__global__ void __launch_bounds__(256) scale_positions(
    Pt* d_X, const int* __restrict__ d_n, int n_max, float scale)
{
    const int n = live_cells(d_n, n_max);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += gridDim.x * blockDim.x) {
        d_X[i].x *= scale;
        d_X[i].y *= scale;
        d_X[i].z *= scale;
    }
}

inline int seeding_blocks(int n)
{
    const int blocks = ceil_div(n > 0 ? n : 1, 256);
    const int limit = 32 * sm_count();
    return blocks < limit ? blocks : limit;
}

// Steps the reference spends on relaxing n cells (inits.cuh:96-112, 128-144).
inline int sphere_relaxation_steps(int n)
{
    return n <= 100 ? 500 : n <= 1000 ? 1000 : n <= 6000 ? 2000 : 3000;
}
inline int cuboid_relaxation_steps(int n)
{
    return n <= 3000 ? 1000 : n <= 12000 ? 2000 : 3000;
}

template<typename Pt, template<typename> class Solver>
void publish_count_and_mirror(Solution<Pt, Solver>& points)
{
    YB_CUDA(cudaMemcpyAsync(points.d_n, points.h_n, sizeof(int),
        cudaMemcpyHostToDevice, points.stream));
    YB_CUDA(cudaStreamSynchronize(points.stream));
    points.copy_to_host();
}

}  // namespace yb


// Uniformly filled ball of radius (n / 0.64)^(1/3) dist_to_nb / 2, the formula
// of random_sphere (inits.cuh:41-48).
template<typename Pt, template<typename> class Solver>
void seeded_sphere(float dist_to_nb, Solution<Pt, Solver>& points,
    unsigned long long seed, unsigned int n_0 = 0)
{
    const int n = *points.h_n;
    assert(static_cast<int>(n_0) < n && n <= points.n_max);
    const float r_max = static_cast<float>(
        pow((n - static_cast<int>(n_0)) / 0.64, 1. / 3) * dist_to_nb / 2);
    yb::seed_ball<Pt><<<yb::seeding_blocks(n - n_0), 256, 0, points.stream>>>(
        points.d_X, static_cast<int>(n_0), n, r_max, seed);
    yb::publish_count_and_mirror(points);
}

// Uniformly filled box holding as many cells as random sphere packing allows
// (inits.cuh:52-75); sets *h_n.
template<typename Pt, template<typename> class Solver>
void seeded_cuboid(float dist_to_nb, float3 minimum, float3 maximum,
    Solution<Pt, Solver>& points, unsigned long long seed, unsigned int n_0 = 0)
{
    const float3 dimension{
        maximum.x - minimum.x, maximum.y - minimum.y, maximum.z - minimum.z};
    const double cube_volume =
        static_cast<double>(dimension.x) * dimension.y * dimension.z;
    const double sphere_volume = 4. / 3 * M_PI * pow(dist_to_nb / 2, 3);
    const int n_new = static_cast<int>(cube_volume / sphere_volume * 0.64);
    assert(static_cast<int>(n_0) + n_new <= points.n_max);
    *points.h_n = static_cast<int>(n_0) + n_new;
    yb::seed_box<Pt><<<yb::seeding_blocks(n_new), 256, 0, points.stream>>>(
        points.d_X, static_cast<int>(n_0), *points.h_n, minimum, dimension, seed);
    yb::publish_count_and_mirror(points);
}

// Ball at neighbour distance 0.6, relaxed with relu_force (equilibrium
// distance 0.8) and rescaled to dist_to_nb -- relaxed_sphere (inits.cuh:96-112)
// without host round trips. relax_steps < 0 takes the reference's step count.
template<typename Pt, template<typename> class Solver>
void relaxed_seeded_sphere(float dist_to_nb, Solution<Pt, Solver>& points,
    unsigned long long seed, unsigned int n_0 = 0, int relax_steps = -1)
{
    seeded_sphere(0.6f, points, seed, n_0);
    const int n = *points.h_n;
    if (relax_steps < 0) relax_steps = yb::sphere_relaxation_steps(n);
    for (int i = 0; i < relax_steps; i++)
        points.template take_step<relu_force>(0.1f);
    yb::scale_positions<Pt><<<yb::seeding_blocks(n), 256, 0, points.stream>>>(
        points.d_X, points.d_n, points.n_max, dist_to_nb / 0.8f);
    yb::publish_count_and_mirror(points);
}

template<typename Pt, template<typename> class Solver>
void relaxed_seeded_cuboid(float dist_to_nb, float3 minimum, float3 maximum,
    Solution<Pt, Solver>& points, unsigned long long seed, unsigned int n_0 = 0,
    int relax_steps = -1)
{
    const float scale = dist_to_nb / 0.8f;
    const float3 lo{minimum.x / scale, minimum.y / scale, minimum.z / scale};
    const float3 hi{maximum.x / scale, maximum.y / scale, maximum.z / scale};
    seeded_cuboid(0.8f, lo, hi, points, seed, n_0);
    const int n = *points.h_n;
    if (relax_steps < 0) relax_steps = yb::cuboid_relaxation_steps(n);
    for (int i = 0; i < relax_steps; i++)
        points.template take_step<relu_force>(0.1f);
    yb::scale_positions<Pt><<<yb::seeding_blocks(n), 256, 0, points.stream>>>(
        points.d_X, points.d_n, points.n_max, scale);
    yb::publish_count_and_mirror(points);
}
