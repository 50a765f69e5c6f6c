// The named models served by the C ABI (include/yalla_b200.h).
//
// Everything in this file is USER-LEVEL ya||a code: pairwise functors, division
// kernels and generic-force callbacks written against the public header API
// only, following the reference examples cited per model. The file is compiled
// twice -- against this repo's include/ (product) and against
// /root/reference/include (baseline, oracle/_ref) -- so both builds integrate
// exactly the same model source.
#pragma once

#include <curand_kernel.h>

#include "dtypes.cuh"
#include "inits.cuh"
#include "links.cuh"
#include "polarity.cuh"
#include "property.cuh"
#include "solvers.cuh"
#include "utils.cuh"
#ifdef YALLA_B200  // extensions of this repo's headers
#include "b200/division.cuh"
#endif

// Point types have to be declared at namespace scope (MAKE_PT specialises
// Is_vector). 7-float cell of examples/branching.cu:57, and a 4-lane type for
// the lane-count dispatch of the grid/link entry points.
MAKE_PT(Branching_cell, theta, phi, u, v);
MAKE_PT(Lanes4_cell, w);

namespace models {

using Cell = Branching_cell;

// ---- float3 springs ------------------------------------------------------------
// Parameters live in __device__ variables so that configs can change them
// without re-instantiating the force kernel (set through yb_sim_set_param).
__device__ float d_spring_length = 0.5f;

// All-pairs linear spring, examples/springs.cu:14-21.
__device__ float3 spring(float3 Xi, float3 r, float dist, int i, int j)
{
    float3 dF{0};
    if (i == j) return dF;

    dF = r * (d_spring_length - dist) / dist;
    return dF;
}

// Spring cut off at distance 1, tests/test_solvers.cu:44-53.
__device__ float3 clipped_spring(float3 Xi, float3 r, float dist, int i, int j)
{
    float3 dF{0};
    if (i == j) return dF;

    if (dist >= 1) return dF;

    dF = r * (d_spring_length - dist) / dist;
    return dF;
}


// ---- Po_cell: epithelium and growth --------------------------------------------
enum Cell_types { mesenchyme, epithelium };

__device__ Cell_types* d_type;
__device__ int* d_mes_nbs;  // number of mesenchymal neighbours
__device__ int* d_epi_nbs;

// ReLU forces plus bending resistance, examples/epithelium.cu:16-31.
__device__ Po_cell layer_force(Po_cell Xi, Po_cell r, float dist, int i, int j)
{
    Po_cell dF{0};
    if (i == j) return dF;

    if (dist > 1) return dF;

    auto F = fmaxf(0.7 - dist, 0) * 2 - fmaxf(dist - 0.8, 0);
    dF.x = r.x * F / dist;
    dF.y = r.y * F / dist;
    dF.z = r.z * F / dist;

    dF += bending_force(Xi, r, dist) * 0.2;
    return dF;
}

// Mesenchyme enveloped by epithelium, examples/passive_growth.cu:29-57. Counts
// neighbours by type as a side effect (one thread owns i).
__device__ Po_cell relu_w_epithelium(
    Po_cell Xi, Po_cell r, float dist, int i, int j)
{
    Po_cell dF{0};
    if (i == j) return dF;

    if (dist > 1) return dF;

    float F;
    if (d_type[i] == d_type[j]) {
        F = fmaxf(0.7 - dist, 0) * 2 - fmaxf(dist - 0.8, 0);
    } else {
        F = fmaxf(0.8 - dist, 0) * 2 - fmaxf(dist - 0.9, 0);
    }
    dF.x = r.x * F / dist;
    dF.y = r.y * F / dist;
    dF.z = r.z * F / dist;

    if (d_type[j] == mesenchyme)
        d_mes_nbs[i] += 1;
    else
        d_epi_nbs[i] += 1;

    if (d_type[i] == mesenchyme or d_type[j] == mesenchyme) return dF;

    dF += bending_force(Xi, r, dist) * 0.15;
    return dF;
}

// Cell division, examples/passive_growth.cu:59-91: mesenchymal cells divide
// at `rate`, epithelial cells when they have more mesenchymal than epithelial
// neighbours; the daughter is appended at slot atomicAdd(d_n_cells, 1).
// snapshot_count freezes the count first, so that all blocks agree on which
// cells existed when the step began.
__global__ void snapshot_count(const int* d_n_cells, int* d_n_at_launch)
{
    *d_n_at_launch = *d_n_cells;
}

// The number of cells at launch is read from d_n_cells by the first thread of
// every block (no host round trip to size the launch); blocks beyond it exit.
__global__ void proliferate(float rate, float mean_dist, int n_max,
    curandState* d_state, Po_cell* d_X, float3* d_old_v, int* d_n_cells,
    const int* d_n_at_launch)
{
    const int n_cells = *d_n_at_launch;
    auto i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells) return;  // Dividing new cells is problematic!

    switch (d_type[i]) {
        case mesenchyme: {
            auto rnd = curand_uniform(&d_state[i]);
            if (rnd > rate) return;
            break;
        }
        case epithelium: {
            if (d_epi_nbs[i] > d_mes_nbs[i]) return;
        }
    }

    auto n = atomicAdd(d_n_cells, 1);
    if (n >= n_max) {  // full: undo instead of writing out of bounds
        atomicSub(d_n_cells, 1);
        return;
    }
    auto theta = acosf(2. * curand_uniform(&d_state[i]) - 1);
    auto phi = curand_uniform(&d_state[i]) * 2 * M_PI;
    d_X[n].x = d_X[i].x + mean_dist / 4 * sinf(theta) * cosf(phi);
    d_X[n].y = d_X[i].y + mean_dist / 4 * sinf(theta) * sinf(phi);
    d_X[n].z = d_X[i].z + mean_dist / 4 * cosf(theta);
    d_X[n].theta = d_X[i].theta;
    d_X[n].phi = d_X[i].phi;
    d_type[n] = d_type[i];
    d_mes_nbs[n] = 0;
    d_epi_nbs[n] = 0;
    d_old_v[n] = d_old_v[i];
}


// The same division rule for Cell_division (b200/division.cuh, product build
// only): mesenchyme divides with probability d_prolif_rate per step, an
// epithelial cell whenever it has no more epithelial than mesenchymal
// neighbours; the daughter inherits the type and starts with empty counters.
#ifdef YALLA_B200
__device__ float d_prolif_rate = 0.f;

__device__ float growth_division_rate(int i, const Po_cell&)
{
    if (d_type[i] == mesenchyme) return d_prolif_rate;
    return d_epi_nbs[i] > d_mes_nbs[i] ? 0.f : 1.f;
}

__device__ void growth_inherit(int mother, int daughter)
{
    d_type[daughter] = d_type[mother];
    d_mes_nbs[daughter] = 0;
    d_epi_nbs[daughter] = 0;
}
#endif


// ---- 7-float Cell: branching ------------------------------------------------------
// Turing parameters of examples/branching.cu:21-31.
const auto lambda = 0.0075;
const auto D_u = 0.001;
const auto D_v = 0.2;
const auto f_v = 1.0;
const auto f_u = 80.0;
const auto g_u = 80.0;
const auto m_u = 0.25;
const auto m_v = 0.75;
const auto s_u = 0.05;

// Meinhardt reaction in the self-interaction, diffusion + adhesion + bending
// between neighbours, atomic neighbour counters; examples/branching.cu:60-110.
__device__ Cell epi_turing_mes_noturing(Cell Xi, Cell r, float dist, int i, int j)
{
    Cell dF{0};

    if (i == j) {
        if (d_type[i] == epithelium) {
            dF.u = lambda *
                   ((f_u * Xi.u * Xi.u) / (1 + f_v * Xi.v) - m_u * Xi.u + s_u);
            dF.v = lambda * (g_u * Xi.u * Xi.u - m_v * Xi.v);

            // Prevent negative values
            if (-dF.u > Xi.u) dF.u = 0.0f;
            if (-dF.v > Xi.v) dF.v = 0.0f;
        }
        return dF;
    }

    if (dist > 1.0f) return dF;

    float F;
    if (d_type[i] == d_type[j]) {
        F = fmaxf(0.7 - dist, 0) * 2 - fmaxf(dist - 0.8, 0);
    } else {
        F = fmaxf(0.8 - dist, 0) * 2 - fmaxf(dist - 0.9, 0);
    }
    dF.x = r.x * F / dist;
    dF.y = r.y * F / dist;
    dF.z = r.z * F / dist;

    if (d_type[i] == epithelium && d_type[j] == epithelium) {
        dF.u = -D_u * r.u;
        dF.v = -D_v * r.v;

        if (-dF.u > Xi.u) dF.u = 0.0f;
        if (-dF.v > Xi.v) dF.v = 0.0f;

        dF += bending_force(Xi, r, dist) * 0.2;
    } else {
        dF.v = -D_v * r.v;  // Diffuses into mesenchyme to induce proliferation
    }

    if (d_type[j] == epithelium)
        atomicAdd(&d_epi_nbs[i], 1);
    else
        atomicAdd(&d_mes_nbs[i], 1);

    return dF;
}



// ---- branching + growth + protrusions (BASELINE.json configs[3]) --------------------
// The branching cell with cell division (examples/branching.cu:113-138, without
// the lineage bookkeeping) and one protrusion per cell that a kernel rewires
// every step (examples/intercalation_w_gradient.cu:119-173), pulling through
// link_forces. The morphogen v plays the part of that example's w (cells far
// from the source align their protrusion along its gradient), u that of f.
// Launches are sized for the capacity and read the live count on the device,
// so that one iteration needs nothing from the host.
const auto prots_per_cell = 1;
const auto r_protrusion = 2.0f;

__global__ void set_link_count(const int* d_n_cells, int* d_n_links)
{
    *d_n_links = *d_n_cells * prots_per_cell;
}

__global__ void update_protrusions(const int* d_n_cells,
    const Grid* __restrict__ d_grid, const Cell* __restrict d_X,
    curandState* d_state, Link* d_link)
{
    const int n_cells = *d_n_cells;
    auto i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells * prots_per_cell) return;

    auto j = static_cast<int>((i + 0.5) / prots_per_cell);
    auto rand_nb_cube =
        d_grid->d_cube_id[j] +
        d_nhood[min(static_cast<int>(curand_uniform(&d_state[i]) * 27), 26)];
    if (rand_nb_cube < 0 or rand_nb_cube >= d_grid->n_cubes) return;
    auto cells_in_cube =
        d_grid->d_cube_end[rand_nb_cube] - d_grid->d_cube_start[rand_nb_cube];
    if (cells_in_cube < 1) return;

    auto a = d_grid->d_point_id[j];
    auto b =
        d_grid->d_point_id[d_grid->d_cube_start[rand_nb_cube] +
                           min(static_cast<int>(
                                   curand_uniform(&d_state[i]) * cells_in_cube),
                               cells_in_cube - 1)];
    if (a == b) return;

    if ((d_type[a] != mesenchyme) or (d_type[b] != mesenchyme)) return;

    auto link = &d_link[a * prots_per_cell + i % prots_per_cell];

    auto old_r = d_X[link->a] - d_X[link->b];
    auto old_dist = norm3df(old_r.x, old_r.y, old_r.z);
    auto new_r = d_X[a] - d_X[b];
    auto new_dist = norm3df(new_r.x, new_r.y, new_r.z);
    if (new_dist > r_protrusion) return;

    auto not_initialized = link->a == link->b;
    auto noise = curand_uniform(&d_state[i]);
    auto superficial = d_X[a].v + d_X[b].v > 0.3f;
    auto parallel_to_v_gradient = false;
    auto normal_to_u_gradient = false;
    if (superficial) {
        normal_to_u_gradient =
            fabs(new_r.u / new_dist) < fabs(old_r.u / old_dist) * (1.f - noise);
    } else {
        parallel_to_v_gradient =
            fabs(new_r.v / new_dist) > fabs(old_r.v / old_dist) * (1.f - noise);
    }

    if (not_initialized or parallel_to_v_gradient or normal_to_u_gradient) {
        link->a = a;
        link->b = b;
    }
}

// Division rule of examples/branching.cu:113-138 (mesenchyme at mes_rate, here
// without the v threshold; epithelium at epi_rate where it has room and touches
// mesenchyme); daughters halve u and v with their mothers.
__global__ void proliferate_branching(float mes_rate, float epi_rate,
    float mean_distance, int n_max, curandState* d_state, Cell* d_X,
    float3* d_old_v, int* d_n_cells, const int* d_n_at_launch)
{
    const int n_cells = *d_n_at_launch;
    auto i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells) return;  // Dividing new cells is problematic!

    auto rnd = curand_uniform(&d_state[i]);
    switch (d_type[i]) {
        case mesenchyme: {
            if (rnd > mes_rate) return;

            break;
        }
        case epithelium: {
            if (d_epi_nbs[i] > 5) return;

            if (d_mes_nbs[i] <= 0) return;

            if (rnd > epi_rate) return;
        }
    }

    auto n = atomicAdd(d_n_cells, 1);
    if (n >= n_max) {  // full: undo instead of writing out of bounds
        atomicSub(d_n_cells, 1);
        return;
    }
    auto theta = acosf(2. * curand_uniform(&d_state[i]) - 1);
    auto phi = curand_uniform(&d_state[i]) * 2 * M_PI;
    d_X[n].x = d_X[i].x + mean_distance / 4 * sinf(theta) * cosf(phi);
    d_X[n].y = d_X[i].y + mean_distance / 4 * sinf(theta) * sinf(phi);
    d_X[n].u = d_X[i].u / 2;
    d_X[n].z = d_X[i].z + mean_distance / 4 * cosf(theta);
    d_X[i].u = d_X[i].u / 2;
    d_X[n].v = d_X[i].v / 2;
    d_X[i].v = d_X[i].v / 2;
    d_X[n].theta = d_X[i].theta;
    d_X[n].phi = d_X[i].phi;
    d_type[n] = d_type[i];
    d_mes_nbs[n] = 0;
    d_epi_nbs[n] = 0;
    d_old_v[n] = d_old_v[i];
}

}  // namespace models
