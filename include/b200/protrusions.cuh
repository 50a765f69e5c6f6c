// Extension: rewiring of protrusions as a library operation, reproducible bit
// for bit.
//
// Every model of the reference with protrusions carries its own
// `update_protrusions` kernel (examples/sorting_prot.cu:33-74,
// intercalation_w_gradient.cu:119-173, limb_bud.cu, ...): per link, one curand
// XORWOW state; pick a random one of the 27 cubes around the cell in a Grid
// built with cube size r_protrusion, pick a random cell in it, and let a
// model-specific rule decide whether the pair replaces the link's current one.
// Around it the host reads the cell count back twice per step (set_d_n,
// Grid::build) and sizes the launch with it. Protrusion_update does the same
// job with
//   * a counter-based generator, Philox4_32_10 keyed by (seed, link slot, call
//     number): the draws of a link do not depend on launch geometry or on what
//     other links drew, and a run can be repeated exactly;
//   * all counts on the device: the links' count follows the cells' count, the
//     grid is built with Grid::build_live, the launch is sized for the
//     capacity -- nothing waits for the host, everything is capturable;
//   * the whole cube (the examples' `end - start` leaves out its last cell).
//
//   // true: the candidate pair (a, b) replaces `current` (a == b: no link yet)
//   __device__ bool rule(const Pt* d_X, int a, int b, Link current, float noise);
//   Protrusion_update<Pt> protrusions_update{n_max, prots_per_cell, seed, grid_size};
//   protrusions_update.template rewire<rule>(cells, protrusions, r_protrusion);
//
// Link slot a * prots_per_cell + k belongs to cell a, as in the examples.
#pragma once

#include <cuda_runtime.h>
#include <curand_kernel.h>

#include "../cudebug.cuh"
#include "../links.cuh"
#include "../solvers.cuh"
#include "grid_build.cuh"
#include "layout.cuh"

template<typename Pt>
using Protrusion_rule = bool(
    const Pt* __restrict__ d_X, int a, int b, Link current, float noise);

// The plainest rule: take the first partner that comes along, keep it.
template<typename Pt>
__device__ bool protrusion_if_none(
    const Pt* __restrict__ d_X, int a, int b, Link current, float noise)
{
    return current.a == current.b;
}

namespace yb {

__global__ void follow_cell_count(
    const int* d_n_cells, int prots_per_cell, int n_links_max, int* d_n_links)
{
    const long long n = static_cast<long long>(*d_n_cells) * prots_per_cell;
    *d_n_links = n < n_links_max ? static_cast<int>(n) : n_links_max;
}

template<typename Pt, Protrusion_rule<Pt> rule>
__global__ void __launch_bounds__(128) rewire_protrusions(
    const int* __restrict__ d_n_cells, int n_max_cells, int prots_per_cell,
    const int* __restrict__ cube_id, const int* __restrict__ point_id,
    const int* __restrict__ cube_start, const int* __restrict__ cube_end,
    int grid_size, const Pt* __restrict__ d_X, float r_protrusion,
    unsigned long long seed, unsigned long long call, Link* d_link)
{
    const int n_cells = live_cells(d_n_cells, n_max_cells);
    const long long n_slots = static_cast<long long>(n_cells) * prots_per_cell;
    const int n_cubes = grid_size * grid_size * grid_size;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
         i < n_slots; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        curandStatePhilox4_32_10_t state;
        curand_init(seed, static_cast<unsigned long long>(i), 4 * call, &state);
        const float4 draw = curand_uniform4(&state);  // each in (0, 1]

        // sorted slot j of the grid <-> cell a; the k-th protrusion of a
        const int j = static_cast<int>(i / prots_per_cell);
        const int k = static_cast<int>(i % prots_per_cell);
        const int which = min(static_cast<int>(draw.x * 27.f), 26);
        const int cube = __ldg(cube_id + j) + (which % 3 - 1) +
                         ((which / 3) % 3 - 1) * grid_size +
                         (which / 9 - 1) * grid_size * grid_size;
        if (cube < 0 || cube >= n_cubes) continue;
        const int start = __ldg(cube_start + cube);
        const int in_cube = __ldg(cube_end + cube) - start + 1;  // -2 - -1 + 1 = 0
        if (start < 0 || in_cube < 1) continue;

        const int a = __ldg(point_id + j);
        const int b = __ldg(point_id + start +
                            min(static_cast<int>(draw.y * in_cube), in_cube - 1));
        if (a == b) continue;
        const Pt r = load_pt(d_X, a) - load_pt(d_X, b);
        if (norm3df(r.x, r.y, r.z) > r_protrusion) continue;

        Link* link = d_link + static_cast<long long>(a) * prots_per_cell + k;
        if (rule(d_X, a, b, *link, draw.z)) *link = Link{a, b};
    }
}

}  // namespace yb


template<typename Pt>
class Protrusion_update {
public:
    Protrusion_update(int n_max_cells, int prots_per_cell,
        unsigned long long seed, int grid_size = 50)
        : n_max_cells{n_max_cells}, prots_per_cell{prots_per_cell}, seed{seed},
          grid{n_max_cells, grid_size}
    {}

    // One round of rewiring, enqueued on the solver's stream. `protrusions`
    // needs room for n_max_cells * prots_per_cell links.
    template<Protrusion_rule<Pt> rule, template<typename> class Solver>
    void rewire(Solution<Pt, Solver>& cells, Links& protrusions, float r_protrusion)
    {
        assert(protrusions.n_max >= n_max_cells * prots_per_cell);
        const cudaStream_t s = cells.stream;
        yb::follow_cell_count<<<1, 1, 0, s>>>(
            cells.d_n, prots_per_cell, protrusions.n_max, protrusions.d_n);
        grid.stream = s;
        grid.build_live(cells.d_n, cells.d_X, r_protrusion);
        const long long slots =
            static_cast<long long>(n_max_cells) * prots_per_cell;
        const int blocks = static_cast<int>(
            (slots + 127) / 128 < yb::sm_count() * 32 ? (slots + 127) / 128
                                                      : yb::sm_count() * 32);
        yb::rewire_protrusions<Pt, rule><<<blocks > 0 ? blocks : 1, 128, 0, s>>>(
            cells.d_n, n_max_cells, prots_per_cell, grid.d_cube_id,
            grid.d_point_id, grid.d_cube_start, grid.d_cube_end, grid.grid_size,
            cells.d_X, r_protrusion, seed, calls, protrusions.d_link);
        YB_CUDA(cudaGetLastError());
        protrusions.mark_changed();
        calls++;
    }

    unsigned long long calls = 0;  // the "step" part of the generator key
    Grid grid;                     // the last rewiring's neighbour grid

private:
    const int n_max_cells, prots_per_cell;
    const unsigned long long seed;
};
