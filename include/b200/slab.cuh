// Internal: packing and unpacking kernels for the slab domain decomposition
// (Heun_solver::slab_* in solvers.cuh, driven by yalla_b200/dd.py).
//
// A slab owns the cells with z_lo <= z < z_hi. Everything a step needs from the
// host is fixed at set-up time; cell counts stay on the device, so a decomposed
// step never synchronises with the host:
//
//   exchange buffers   [header: 4 floats, header[0] = bits(count)]
//                      [count records of (lanes + 3) floats: Pt, old_v.xyz]
//   always sent at full capacity (NVLink makes that cheaper than a host round
//   trip for the count); the receiver reads the count from the header.
//
// Packing is a stable single-pass stream compaction (slab_select), so the
// order of ghosts and migrants, and with it every floating-point sum on the
// receiving side, is reproducible.
#pragma once

#include <cuda_runtime.h>
#include <stdlib.h>

#include "grid_build.cuh"
#include "layout.cuh"

namespace yb {

constexpr int SLAB_HEADER = 4;  // floats in front of the records

// Records are 8-byte aligned when they hold an even number of floats (the
// header is 16 bytes): stored as float2s then -- exchange buffers may live in a
// neighbour's memory, where every store is an NVLink transaction.
template<typename Pt>
__device__ __forceinline__ void write_record(
    float* record, const Pt* P, const float3* v, int i)
{
    using L = Layout<Pt>;
    constexpr int W = L::lanes + 3;
    const float* x = reinterpret_cast<const float*>(P + i);
    const float* w = reinterpret_cast<const float*>(v + i);
    float r[W];
#pragma unroll
    for (int k = 0; k < L::lanes; k++) r[k] = x[k];
    r[L::lanes + 0] = w[0];
    r[L::lanes + 1] = w[1];
    r[L::lanes + 2] = w[2];
    if (W % 2 == 0) {
        float2* out = reinterpret_cast<float2*>(record);
#pragma unroll
        for (int k = 0; k < W / 2; k++) out[k] = make_float2(r[2 * k], r[2 * k + 1]);
    } else {
#pragma unroll
        for (int k = 0; k < W; k++) record[k] = r[k];
    }
}

template<typename Pt>
__device__ __forceinline__ void read_record(
    const float* record, Pt* P, float3* v, int i)
{
    using L = Layout<Pt>;
    float* x = reinterpret_cast<float*>(P + i);
    float* w = reinterpret_cast<float*>(v + i);
#pragma unroll
    for (int k = 0; k < L::lanes; k++) x[k] = record[k];
    w[0] = record[L::lanes + 0];
    w[1] = record[L::lanes + 1];
    w[2] = record[L::lanes + 2];
}

// Which owned cells go to the lower / upper neighbour, and where: a
// single-pass stable stream compaction with
// decoupled look-back (grid_build.cuh) over tiles of SCAN_TILE cells. Every
// tile counts its cells bound for the lower and for the upper neighbour, warps
// 0 and 1 resolve the two running prefixes at the same time, and the tile then
// writes its records straight to their final positions. In a migration round
// every cell goes exactly one way, so the rank of a cell that stays is its
// index minus the cells before it that leave: the stayers are compacted in the
// same pass. The order of records is ascending cell index.
//
// Cell identity is not tracked across a decomposed run, so a migration round
// may also PERMUTE the cells that stay: with `order` (the cube-ordered pos4 of
// the last force evaluation, original index in .w) the pass walks the cells in
// cube order and compacts them in that order. From the second step on the
// slab's own storage order is then nearly cube order, which turns the
// scattered record writes of the next grid builds (place_cells) and the
// scattered force stores of the sweeps into near-sequential traffic. A third
// running count (entries of `order` that are ghosts) gives the ranks.
constexpr int SELECT_SUB = SCAN_TILE / SCAN_THREADS;  // cells per thread

template<typename Pt>
__global__ void __launch_bounds__(SCAN_THREADS) slab_select(Step_ctl* ctl,
    Step_ctl* scan_ctl, const Pt* __restrict__ P, const float3* __restrict__ v,
    float lo_edge, float hi_edge, int has_lower, int has_upper, int migration,
    float* __restrict__ send_lo, float* __restrict__ send_hi, int capacity,
    Pt* __restrict__ X_tmp, float3* __restrict__ v_tmp, int* n_stay,
    unsigned long long* status_lo, unsigned long long* status_hi, int n_tiles,
    const float4* __restrict__ order, const int* __restrict__ d_n_total,
    int n_max, unsigned long long* status_ghost)
{
    constexpr int W = Layout<Pt>::lanes + 3;
    constexpr int WARPS = SCAN_THREADS / 32;
    __shared__ int s_tile;
    __shared__ int s_count[3][SELECT_SUB][WARPS];  // then: exclusive prefixes
    __shared__ int s_total[3];
    __shared__ int s_tile_prefix[3];

    const int t = threadIdx.x;
    const int lane_id = t & 31, warp_id = t >> 5;
    if (t == 0) s_tile = atomicAdd(&scan_ctl->scan_next_tile, 1);
    __syncthreads();
    const int tile = s_tile;
    const unsigned epoch =
        static_cast<unsigned>(*(volatile int*)&scan_ctl->scan_epoch) & 0x3fffffffu;
    const int n_owned = ctl->n_owned;
    const bool permute = order != nullptr;
    // entries to walk: the owned cells, or every slot of `order` (with ghosts)
    const int n = permute ? live_cells(d_n_total, n_max) : n_owned;
    const int first = tile * SCAN_TILE;

    // sub-block u holds entries first + u * THREADS + t: coalesced, and ranks
    // in (u, warp, lane) order are ranks in entry order
    unsigned flags = 0;  // bit 2u: goes down, bit 2u + 1: goes up
    unsigned ghosts = 0;  // bit u: the entry is a ghost (permuted walk only)
    int cell[SELECT_SUB];
#pragma unroll
    for (int u = 0; u < SELECT_SUB; u++) {
        const int q = first + u * SCAN_THREADS + t;
        int lo = 0, hi = 0, ghost = 0;
        cell[u] = q;
        if (q < n) {
            if (permute) {
                cell[u] = __float_as_int(__ldg(&order[q].w));
                ghost = cell[u] >= n_owned;
            }
            if (!ghost) {
                const float z =
                    __ldg(reinterpret_cast<const float*>(P + cell[u]) + 2);
                lo = has_lower && z < lo_edge;
                hi = has_upper && z >= hi_edge;
            }
        }
        flags |= (unsigned(lo) << (2 * u)) | (unsigned(hi) << (2 * u + 1));
        ghosts |= unsigned(ghost) << u;
        const unsigned lo_mask = __ballot_sync(0xffffffffu, lo);
        const unsigned hi_mask = __ballot_sync(0xffffffffu, hi);
        const unsigned ghost_mask = __ballot_sync(0xffffffffu, ghost);
        if (lane_id == 0) {
            s_count[0][u][warp_id] = __popc(lo_mask);
            s_count[1][u][warp_id] = __popc(hi_mask);
            s_count[2][u][warp_id] = __popc(ghost_mask);
        }
    }
    __syncthreads();
    // exclusive scan of the SUB x WARPS counts, one warp per running count
    if (warp_id < (permute ? 3 : 2)) {
        constexpr int ENTRIES = SELECT_SUB * WARPS, PER_LANE = ENTRIES / 32;
        int* counts = &s_count[warp_id][0][0];
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
        unsigned long long* const status =
            warp_id == 0 ? status_lo : (warp_id == 1 ? status_hi : status_ghost);
        const int exclusive =
            scan_lookback(status, tile, epoch, aggregate, lane_id);
        if (lane_id == 0) {
            s_total[warp_id] = aggregate;
            s_tile_prefix[warp_id] = exclusive;
        }
    }
    __syncthreads();
    const int prefix_lo = s_tile_prefix[0], prefix_hi = s_tile_prefix[1];
    const int prefix_ghost = permute ? s_tile_prefix[2] : 0;

#pragma unroll
    for (int u = 0; u < SELECT_SUB; u++) {
        const int q = first + u * SCAN_THREADS + t;
        const int i = cell[u];
        const int lo = (flags >> (2 * u)) & 1, hi = (flags >> (2 * u + 1)) & 1;
        const int ghost = (ghosts >> u) & 1;
        const unsigned lo_mask = __ballot_sync(0xffffffffu, lo);
        const unsigned hi_mask = __ballot_sync(0xffffffffu, hi);
        const unsigned ghost_mask = __ballot_sync(0xffffffffu, ghost);
        const unsigned below = (1u << lane_id) - 1u;
        const int a = prefix_lo + s_count[0][u][warp_id] + __popc(lo_mask & below);
        const int b = prefix_hi + s_count[1][u][warp_id] + __popc(hi_mask & below);
        const int g = permute ? prefix_ghost + s_count[2][u][warp_id] +
                                    __popc(ghost_mask & below)
                              : 0;
        if (q >= n || ghost) continue;
        if (lo && a < capacity)
            write_record(send_lo + SLAB_HEADER + size_t(a) * W, P, v, i);
        if (hi && b < capacity)
            write_record(send_hi + SLAB_HEADER + size_t(b) * W, P, v, i);
        if (migration && !lo && !hi) {
            const int to = q - a - b - g;
            store_pt(X_tmp, to, load_pt(P, i));
            v_tmp[to] = v[i];
        }
    }

    // the tile that holds the last entry knows the totals
    const int last_tile = n > 0 ? (n - 1) / SCAN_TILE : 0;
    if (t == 0 && tile == last_tile) {
        const int n_lo = prefix_lo + s_total[0], n_hi = prefix_hi + s_total[1];
        if (n_lo > capacity || n_hi > capacity)
            atomicAdd(&ctl->out_of_grid, 1 << 20);
        send_lo[0] = __int_as_float(min(n_lo, capacity));
        send_hi[0] = __int_as_float(min(n_hi, capacity));
        if (migration) *n_stay = n_owned - n_lo - n_hi;
    }
    // The last tile to finish re-arms the control words for the next launch.
    if (t == 0) {
        __threadfence();
        if (atomicAdd(&scan_ctl->scan_tiles_done, 1) == n_tiles - 1) {
            scan_ctl->scan_next_tile = 0;
            scan_ctl->scan_tiles_done = 0;
            scan_ctl->scan_epoch = static_cast<int>((epoch + 1u) & 0x3fffffffu);
            __threadfence();
        }
    }
}

// Ghosts: append the received records behind the owned cells of P (X or X1)
// and set the total cell count.
template<typename Pt>
__global__ void __launch_bounds__(256) slab_append_ghosts(Step_ctl* ctl,
    Pt* P, float3* v, const float* __restrict__ recv_lo,
    const float* __restrict__ recv_hi, int has_lower, int has_upper, int n_max,
    int* d_n)
{
    constexpr int W = Layout<Pt>::lanes + 3;
    const int n = ctl->n_owned;
    int n_lo = has_lower ? __float_as_int(recv_lo[0]) : 0;
    int n_hi = has_upper ? __float_as_int(recv_hi[0]) : 0;
    n_lo = max(0, min(n_lo, n_max - n));
    n_hi = max(0, min(n_hi, n_max - n - n_lo));
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_lo + n_hi;
         r += gridDim.x * blockDim.x) {
        const float* record =
            r < n_lo ? recv_lo + SLAB_HEADER + size_t(r) * W
                     : recv_hi + SLAB_HEADER + size_t(r - n_lo) * W;
        read_record(record, P, v, n + r);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *d_n = n + n_lo + n_hi;
        ctl->n_ghosts = n_lo + n_hi;
    }
}

// Migration, step 2: owned cells := stayers, arrivals from below, from above.
template<typename Pt>
__global__ void __launch_bounds__(256) slab_merge(const Step_ctl* ctl,
    const int* __restrict__ n_stay_in, const Pt* __restrict__ X_tmp,
    const float3* __restrict__ v_tmp, const float* __restrict__ recv_lo,
    const float* __restrict__ recv_hi, int has_lower, int has_upper, int n_max,
    Pt* X, float3* v, int* new_count)
{
    constexpr int W = Layout<Pt>::lanes + 3;
    const int n_stay = *n_stay_in;
    int n_lo = has_lower ? __float_as_int(recv_lo[0]) : 0;
    int n_hi = has_upper ? __float_as_int(recv_hi[0]) : 0;
    n_lo = max(0, min(n_lo, n_max - n_stay));
    n_hi = max(0, min(n_hi, n_max - n_stay - n_lo));
    const int total = n_stay + n_lo + n_hi;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < total;
         r += gridDim.x * blockDim.x) {
        if (r < n_stay) {
            store_pt(X, r, load_pt(X_tmp, r));
            v[r] = v_tmp[r];
        } else if (r < n_stay + n_lo) {
            read_record(recv_lo + SLAB_HEADER + size_t(r - n_stay) * W, X, v, r);
        } else {
            read_record(
                recv_hi + SLAB_HEADER + size_t(r - n_stay - n_lo) * W, X, v, r);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *new_count = total;
}

// Runs after slab_merge (separate launch: everyone has read the old count).
__global__ void slab_commit_count(Step_ctl* ctl, const int* new_count, int* d_n)
{
    ctl->n_owned = *new_count;
    *d_n = *new_count;
}

// drift[stage] = reduced sums / reduced count, as operator/= would
// (dtypes.cuh:204-208).
__global__ void slab_set_drift(Step_ctl* ctl, int stage, const float* sums4)
{
    const float inv_n = static_cast<float>(1. / sums4[3]);
    ctl->drift[stage][0] = sums4[0] * inv_n;
    ctl->drift[stage][1] = sums4[1] * inv_n;
    ctl->drift[stage][2] = sums4[2] * inv_n;
}

// Scratch of the compaction: look-back status words of both directions.
struct Slab_scratch {
    int capacity = 0;     // records per exchange buffer
    int n_tiles = 0;      // tiles of SCAN_TILE cells covering n_max
    unsigned long long* status[3] = {nullptr, nullptr, nullptr};
    Step_ctl* scan_ctl = nullptr;  // scan bookkeeping only
    int* n_stay = nullptr;
    int* new_count = nullptr;
    float z_lo = 0.f, z_hi = 0.f, halo = 0.f;
    int has_lower = 0, has_upper = 0;
    // migration rounds re-store the owned cells in cube order
    // (YALLA_B200_SLAB_PERMUTE=0 keeps their index order)
    bool permute = [] {
        const char* env = getenv("YALLA_B200_SLAB_PERMUTE");
        return !(env && env[0] == '0');
    }();

    void allocate(int n_max)
    {
        n_tiles = ceil_div(n_max > 0 ? n_max : 1, SCAN_TILE);
        for (int k = 0; k < 3; k++) {
            YB_CUDA(cudaMalloc(&status[k], n_tiles * sizeof(unsigned long long)));
            YB_CUDA(cudaMemset(status[k], 0, n_tiles * sizeof(unsigned long long)));
        }
        YB_CUDA(cudaMalloc(&scan_ctl, sizeof(Step_ctl)));
        Step_ctl fresh{};
        fresh.scan_epoch = 1;
        YB_CUDA(cudaMemcpy(
            scan_ctl, &fresh, sizeof(fresh), cudaMemcpyHostToDevice));
        YB_CUDA(cudaMalloc(&n_stay, sizeof(int)));
        YB_CUDA(cudaMalloc(&new_count, sizeof(int)));
    }
    void release()
    {
        cudaFree(status[0]);
        cudaFree(status[1]);
        cudaFree(status[2]);
        cudaFree(scan_ctl);
        cudaFree(n_stay);
        cudaFree(new_count);
    }
};

}  // namespace yb
