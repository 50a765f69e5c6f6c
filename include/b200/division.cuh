// Extension: cell division as a library operation, reproducible bit for bit.
//
// Every growing model of the reference carries its own `proliferate` kernel
// (examples/passive_growth.cu:59-91, branching.cu:118-160, ...): one curand
// XORWOW state per cell, the daughter appended at atomicAdd(d_n, 1). Which
// daughter lands in which slot depends on the order the atomics retire, so no
// two runs produce the same arrays, and a decomposed run cannot reproduce a
// single-GPU run. Cell_division does the same job with
//   * a counter-based generator, Philox4_32_10 keyed by (seed, cell, call
//     number): the decision and the division axis of cell i do not depend on
//     launch geometry, on other cells, or on how many calls other tissues made;
//   * a stable one-pass compaction (decoupled look-back, grid_build.cuh): the
//     k-th dividing cell in index order gets slot n + k.
// The cell count stays on the device; nothing here waits for the host.
//
//   __device__ float rate(int i, const Pt& X);        // P(divide) per call
//   __device__ void inherit(int mother, int daughter); // copy model properties
//   Cell_division<Pt> division{n_max, seed};
//   division.template divide<rate, inherit>(cells, mean_dist);
//
// The daughter is placed mean_dist / 4 away from the mother in a uniformly
// random direction and inherits every other member of Pt and the mother's
// old velocity, exactly what the reference's kernels do.
#pragma once

#include <cuda_runtime.h>
#include <curand_kernel.h>

#include "../cudebug.cuh"
#include "grid_build.cuh"
#include "layout.cuh"

template<typename Pt, template<typename> class Solver>
class Solution;

template<typename Pt>
using Division_rate = float(int i, const Pt& X);
using Division_inherit = void(int mother, int daughter);

__device__ inline void inherit_nothing(int, int) {}

namespace yb {

template<typename Pt, Division_rate<Pt> rate, Division_inherit inherit>
__global__ void __launch_bounds__(SCAN_THREADS) divide_cells(int* d_n, int n_max,
    Pt* d_X, float3* d_old_v, float mean_dist, unsigned long long seed,
    unsigned long long call, Step_ctl* scan_ctl, unsigned long long* status,
    int n_tiles, int* n_before, int* n_dropped)
{
    constexpr int SUB = SCAN_TILE / SCAN_THREADS, WARPS = SCAN_THREADS / 32;
    __shared__ int s_tile;
    __shared__ int s_count[SUB][WARPS];  // then: exclusive prefixes
    __shared__ int s_total, s_tile_prefix;

    const int t = threadIdx.x, lane_id = t & 31, warp_id = t >> 5;
    if (t == 0) s_tile = atomicAdd(&scan_ctl->scan_next_tile, 1);
    __syncthreads();
    const int tile = s_tile;
    const unsigned epoch =
        static_cast<unsigned>(*(volatile int*)&scan_ctl->scan_epoch) & 0x3fffffffu;
    // the count this call started with: d_n itself only changes after every
    // tile has read it (the last tile to finish publishes the new count)
    const int n = live_cells(d_n, n_max);
    const int first = tile * SCAN_TILE;

    unsigned divides = 0;  // bit u: cell first + u * THREADS + t divides
    float4 draw[SUB];
#pragma unroll
    for (int u = 0; u < SUB; u++) {
        const int i = first + u * SCAN_THREADS + t;
        int yes = 0;
        if (i < n) {
            curandStatePhilox4_32_10_t state;
            curand_init(seed, static_cast<unsigned long long>(i), 4 * call, &state);
            draw[u] = curand_uniform4(&state);
            yes = draw[u].x <= rate(i, d_X[i]);
        }
        divides |= unsigned(yes) << u;
        const unsigned mask = __ballot_sync(0xffffffffu, yes);
        if (lane_id == 0) s_count[u][warp_id] = __popc(mask);
    }
    __syncthreads();
    if (warp_id == 0) {
        constexpr int PER_LANE = SUB * WARPS / 32;
        int* counts = &s_count[0][0];
        int mine[PER_LANE], sum = 0;
#pragma unroll
        for (int q = 0; q < PER_LANE; q++) {
            mine[q] = counts[lane_id * PER_LANE + q];
            sum += mine[q];
        }
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane_id >= d) incl += up;
        }
        int running = incl - sum;
#pragma unroll
        for (int q = 0; q < PER_LANE; q++) {
            counts[lane_id * PER_LANE + q] = running;
            running += mine[q];
        }
        const int aggregate = __shfl_sync(0xffffffffu, incl, 31);
        const int exclusive = scan_lookback(status, tile, epoch, aggregate, lane_id);
        if (lane_id == 0) s_total = aggregate, s_tile_prefix = exclusive;
    }
    __syncthreads();
    const int tile_prefix = s_tile_prefix;

#pragma unroll
    for (int u = 0; u < SUB; u++) {
        const int i = first + u * SCAN_THREADS + t;
        const int yes = (divides >> u) & 1;
        const unsigned mask = __ballot_sync(0xffffffffu, yes);
        const int rank = tile_prefix + s_count[u][warp_id] +
                         __popc(mask & ((1u << lane_id) - 1u));
        const int daughter = n + rank;
        if (!yes || daughter >= n_max) continue;  // full: the rest is dropped
        const float cos_theta = 2.f * draw[u].y - 1.f;
        const float sin_theta = sqrtf(fmaxf(1.f - cos_theta * cos_theta, 0.f));
        float sin_phi, cos_phi;
        sincospif(2.f * draw[u].z, &sin_phi, &cos_phi);
        Pt X = d_X[i];
        X.x += mean_dist / 4 * sin_theta * cos_phi;
        X.y += mean_dist / 4 * sin_theta * sin_phi;
        X.z += mean_dist / 4 * cos_theta;
        d_X[daughter] = X;
        d_old_v[daughter] = d_old_v[i];
        inherit(i, daughter);
    }

    // the tile that holds the last cell knows how many cells divide
    const int last_tile = n > 0 ? (n - 1) / SCAN_TILE : 0;
    if (t == 0 && tile == last_tile) {
        const int wanted = tile_prefix + s_total;
        const int made = min(wanted, n_max - n);
        *n_before = n;
        *n_dropped = wanted - made;
        // published by the last tile to finish, below
        scan_ctl->n_snapshot = n + made;
    }
    if (t == 0) {
        __threadfence();
        if (atomicAdd(&scan_ctl->scan_tiles_done, 1) == n_tiles - 1) {
            __threadfence();
            *d_n = *(volatile int*)&scan_ctl->n_snapshot;
            scan_ctl->scan_next_tile = 0;
            scan_ctl->scan_tiles_done = 0;
            scan_ctl->scan_epoch = static_cast<int>((epoch + 1u) & 0x3fffffffu);
            __threadfence();
        }
    }
}

}  // namespace yb


template<typename Pt>
class Cell_division {
public:
    Cell_division(int n_max, unsigned long long seed) : n_max{n_max}, seed{seed}
    {
        n_tiles = yb::ceil_div(n_max > 0 ? n_max : 1, yb::SCAN_TILE);
        YB_CUDA(cudaMalloc(&status, n_tiles * sizeof(unsigned long long)));
        YB_CUDA(cudaMemset(status, 0, n_tiles * sizeof(unsigned long long)));
        YB_CUDA(cudaMalloc(&scan_ctl, sizeof(yb::Step_ctl)));
        yb::Step_ctl fresh{};
        fresh.scan_epoch = 1;
        YB_CUDA(cudaMemcpy(scan_ctl, &fresh, sizeof(fresh), cudaMemcpyHostToDevice));
        YB_CUDA(cudaMalloc(&d_report, 2 * sizeof(int)));
        YB_CUDA(cudaMemset(d_report, 0, 2 * sizeof(int)));
    }
    Cell_division(const Cell_division&) = delete;
    Cell_division& operator=(const Cell_division&) = delete;
    ~Cell_division()
    {
        cudaFree(d_report);
        cudaFree(scan_ctl);
        cudaFree(status);
    }

    // One round of divisions, enqueued on the solver's stream.
    template<Division_rate<Pt> rate, Division_inherit inherit = inherit_nothing,
        template<typename> class Solver>
    void divide(Solution<Pt, Solver>& cells, float mean_dist)
    {
        yb::divide_cells<Pt, rate, inherit>
            <<<n_tiles, yb::SCAN_THREADS, 0, cells.stream>>>(cells.d_n, n_max,
                cells.d_X, cells.d_old_v, mean_dist, seed, calls, scan_ctl, status,
                n_tiles, d_report, d_report + 1);
        YB_CUDA(cudaGetLastError());
        calls++;
    }

    // Blocking: cell count before the last call and divisions it had to drop
    // because the tissue was full.
    void last_call(int* n_before, int* n_dropped) const
    {
        int report[2];
        YB_CUDA(cudaMemcpy(report, d_report, sizeof(report), cudaMemcpyDeviceToHost));
        if (n_before) *n_before = report[0];
        if (n_dropped) *n_dropped = report[1];
    }
    unsigned long long calls = 0;  // the "step" part of the generator key

private:
    const int n_max;
    const unsigned long long seed;
    int n_tiles;
    unsigned long long* status = nullptr;
    yb::Step_ctl* scan_ctl = nullptr;
    int* d_report = nullptr;
};
