// Links between cells, to model protrusions (cf.
// https://dx.doi.org/doi:10.1073/pnas.97.19.10448 and
// https://dx.doi.org/doi:10.1371/journal.pcbi.1004952), and flat walls.
//
// API as in the reference (include/links.cuh): Link, Links, Link_force,
// linear_force, link_forces<Pt[, force]>, Wall_force, xy_wall_relu_force,
// wall_forces, link_wall_forces.
//
// link_forces with the default linear_force is the hot part. The reference
// launches one thread per link that does six float atomicAdds into d_dX
// (links.cuh:99-125) and reads the link count back to the host twice per call.
// Here, when called from inside a solver step (the normal case: it is passed
// as, or from, the generic-forces callback), the sum is an atomic-free
// segmented reduction:
//   1. bin_link_ends    bucket both ends of every live link by cell id
//   2. scan_bins        per-cell offsets (same single-pass scan as the grid)
//   3. place_link_ends  write (2 * link + side) entries into the cell's segment
//   4. pull_on_cells    one thread per cell adds its entries' forces in
//                       ascending link order and updates d_dX once.
// The link count stays on the device and the result no longer depends on the
// order in which atomics happen to land. Custom Link_force functors do their
// own (atomic) updates, exactly as in the reference.
#pragma once

#include <assert.h>
#include <curand_kernel.h>
#include <stdlib.h>
#include <time.h>
#include <functional>

#include "b200/grid_build.cuh"
#include "b200/layout.cuh"
#include "cudebug.cuh"
#include "utils.cuh"


struct Link {
    int a, b;
};

using Check_link = std::function<bool(int a, int b)>;

inline bool every_link(int a, int b) { return true; }


namespace yb {

// Seed for generators that the reference seeds with time(NULL). Setting the
// environment variable YALLA_B200_SEED makes runs repeatable.
inline int wall_clock_seed()
{
    const char* fixed = getenv("YALLA_B200_SEED");
    if (fixed && fixed[0]) return atoi(fixed);
    return static_cast<int>(time(NULL));
}

}  // namespace yb


class Links {
public:
    Link* h_link;
    Link* d_link;
    int* h_n = (int*)malloc(sizeof(int));
    int* d_n;
    const int n_max;
    curandState* d_state;
    float strength;

    Links(int n_max, float strength = 1.f / 5)
        : n_max{n_max}, strength{strength}
    {
        const size_t links = static_cast<size_t>(n_max > 0 ? n_max : 1);
        h_link = static_cast<Link*>(malloc(links * sizeof(Link)));
        YB_CUDA(cudaMalloc(&d_link, links * sizeof(Link)));
        YB_CUDA(cudaMalloc(&d_n, sizeof(int)));
        YB_CUDA(cudaMalloc(&d_state, links * sizeof(curandState)));
        *h_n = n_max;
        set_d_n(n_max);
        // all links start as a == b == 0, i.e. inactive
        YB_CUDA(cudaMemset(d_link, 0, links * sizeof(Link)));
        for (int i = 0; i < n_max; i++) h_link[i] = Link{0, 0};
        setup_rand_states<<<(n_max + 32 - 1) / 32, 32>>>(
            n_max, yb::wall_clock_seed(), d_state);
    }
    Links(const Links&) = delete;
    Links& operator=(const Links&) = delete;
    ~Links()
    {
        release_segments();
        cudaFree(d_state);
        cudaFree(d_n);
        cudaFree(d_link);
        free(h_link);
        free(h_n);
    }

    void set_d_n(int n)
    {
        assert(n <= n_max);
        YB_CUDA(cudaMemcpy(d_n, &n, sizeof(int), cudaMemcpyHostToDevice));
    }
    int get_d_n()
    {
        int n;
        YB_CUDA(cudaMemcpy(&n, d_n, sizeof(int), cudaMemcpyDeviceToHost));
        assert(n <= n_max);
        return n;
    }
    // Deactivate (a = b = 0) every link for which check(a, b) holds.
    void reset(Check_link check = every_link)
    {
        copy_to_host();
        for (int i = 0; i < n_max; i++) {
            if (check(h_link[i].a, h_link[i].b)) h_link[i] = Link{0, 0};
        }
        copy_to_device();
    }
    void copy_to_device()
    {
        assert(*h_n <= n_max);
        YB_CUDA(cudaMemcpy(d_link, h_link,
            static_cast<size_t>(n_max) * sizeof(Link), cudaMemcpyHostToDevice));
        YB_CUDA(cudaMemcpy(d_n, h_n, sizeof(int), cudaMemcpyHostToDevice));
    }
    void copy_to_host()
    {
        YB_CUDA(cudaMemcpy(h_link, d_link,
            static_cast<size_t>(n_max) * sizeof(Link), cudaMemcpyDeviceToHost));
        YB_CUDA(cudaMemcpy(h_n, d_n, sizeof(int), cudaMemcpyDeviceToHost));
        assert(*h_n <= n_max);
    }

    // ---- scratch of the segmented reduction (internal) ----------------------
    struct Segments {
        int cell_capacity = 0;
        int n_tiles = 0;
        int* count = nullptr;    // per cell, zero between calls
        int* offset = nullptr;   // per cell (+1, + padding)
        int* arrival = nullptr;  // 2 per link
        int* entry = nullptr;    // 2 per link: 2 * link + side, by cell
        unsigned long long* status = nullptr;
        yb::Step_ctl* ctl = nullptr;
    } segments;

    void reserve_segments(int n_cells)
    {
        if (n_cells <= segments.cell_capacity) return;
        release_segments();
        const size_t bins = static_cast<size_t>(yb::scan_padded(n_cells + 1));
        const size_t ends = 2 * static_cast<size_t>(n_max > 0 ? n_max : 1);
        segments.cell_capacity = n_cells;
        segments.n_tiles = static_cast<int>(bins / yb::SCAN_TILE);
        YB_CUDA(cudaMalloc(&segments.count, bins * sizeof(int)));
        YB_CUDA(cudaMemset(segments.count, 0, bins * sizeof(int)));
        YB_CUDA(cudaMalloc(&segments.offset, bins * sizeof(int)));
        YB_CUDA(cudaMalloc(&segments.arrival, ends * sizeof(int)));
        YB_CUDA(cudaMalloc(&segments.entry, ends * sizeof(int)));
        YB_CUDA(cudaMalloc(&segments.status,
            segments.n_tiles * sizeof(unsigned long long)));
        YB_CUDA(cudaMemset(segments.status, 0,
            segments.n_tiles * sizeof(unsigned long long)));
        YB_CUDA(cudaMalloc(&segments.ctl, sizeof(yb::Step_ctl)));
        yb::Step_ctl fresh{};
        fresh.scan_epoch = 1;
        YB_CUDA(cudaMemcpy(
            segments.ctl, &fresh, sizeof(fresh), cudaMemcpyHostToDevice));
    }

private:
    void release_segments()
    {
        cudaFree(segments.ctl);
        cudaFree(segments.status);
        cudaFree(segments.entry);
        cudaFree(segments.arrival);
        cudaFree(segments.offset);
        cudaFree(segments.count);
        segments = Segments{};
    }
};


// A link force adds its contribution for link (a, b) to d_dX[a] and d_dX[b]
// itself; several links may touch the same cell concurrently, so it must use
// atomics (reference: links.cuh:94-111).
template<typename Pt>
using Link_force = void(const Pt* __restrict__ d_X, const int a, const int b,
    const float strength, Pt* d_dX);

// Constant-magnitude pull of the two ends towards each other.
template<typename Pt>
__device__ void linear_force(const Pt* __restrict__ d_X, const int a,
    const int b, const float strength, Pt* d_dX)
{
    const Pt r = d_X[a] - d_X[b];
    const float dist = norm3df(r.x, r.y, r.z);

    atomicAdd(&d_dX[a].x, -strength * r.x / dist);
    atomicAdd(&d_dX[a].y, -strength * r.y / dist);
    atomicAdd(&d_dX[a].z, -strength * r.z / dist);
    atomicAdd(&d_dX[b].x, strength * r.x / dist);
    atomicAdd(&d_dX[b].y, strength * r.y / dist);
    atomicAdd(&d_dX[b].z, strength * r.z / dist);
}

// One thread per live link; links with a == b are inactive. The name `link` is
// the reference's (links.cuh:113-125).
template<typename Pt, Link_force<Pt> force>
__global__ void link(const Pt* __restrict__ d_X, Pt* d_dX,
    const Link* __restrict__ d_link, const int* __restrict__ d_n_links,
    int n_links_max, float strength)
{
    const int n_links = yb::live_cells(d_n_links, n_links_max);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_links;
         i += gridDim.x * blockDim.x) {
        const int a = d_link[i].a;
        const int b = d_link[i].b;
        if (a == b) continue;
        force(d_X, a, b, strength, d_dX);
    }
}


namespace yb {

// 1. count the live ends per cell and remember each end's arrival rank
__global__ void __launch_bounds__(256) bin_link_ends(
    const Link* __restrict__ d_link, const int* __restrict__ d_n_links,
    int n_links_max, int n_cells, int* count, int* __restrict__ arrival)
{
    const int n_links = live_cells(d_n_links, n_links_max);
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < n_links;
         l += gridDim.x * blockDim.x) {
        const int a = d_link[l].a, b = d_link[l].b;
        const bool active = a != b && a >= 0 && b >= 0 && a < n_cells &&
                            b < n_cells;
        arrival[2 * l] = active ? atomicAdd(count + a, 1) : -1;
        arrival[2 * l + 1] = active ? atomicAdd(count + b, 1) : -1;
    }
}

// 3. entries of a cell, in arrival order
__global__ void __launch_bounds__(256) place_link_ends(
    const Link* __restrict__ d_link, const int* __restrict__ d_n_links,
    int n_links_max, const int* __restrict__ offset,
    const int* __restrict__ arrival, int* __restrict__ entry)
{
    const int n_links = live_cells(d_n_links, n_links_max);
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < n_links;
         l += gridDim.x * blockDim.x) {
        const int at_a = arrival[2 * l], at_b = arrival[2 * l + 1];
        if (at_a < 0) continue;
        entry[__ldg(offset + d_link[l].a) + at_a] = 2 * l;
        entry[__ldg(offset + d_link[l].b) + at_b] = 2 * l + 1;
    }
}

// 4. one thread per cell: visit its entries by ascending link index (repeated
//    minimum search -- segments hold a handful of entries) and add
//    -/+ strength * r / |r| for the a / b end, like linear_force does.
template<typename Pt>
__global__ void __launch_bounds__(128) pull_on_cells(int n_cells,
    const Pt* __restrict__ d_X, const Link* __restrict__ d_link,
    const int* __restrict__ offset, const int* __restrict__ entry,
    float strength, Pt* d_dX)
{
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cells;
         c += gridDim.x * blockDim.x) {
        const int start = __ldg(offset + c), end = __ldg(offset + c + 1);
        if (end <= start) continue;
        float* out = reinterpret_cast<float*>(d_dX + c);
        float fx = out[0], fy = out[1], fz = out[2];
        int last = -1;
        for (int done = start; done < end; done++) {
            int next = 0x7fffffff;
            for (int q = start; q < end; q++) {
                const int e = __ldg(entry + q);
                if (e > last && e < next) next = e;
            }
            last = next;
            const Link l = d_link[next >> 1];
            const Pt r = load_pt(d_X, l.a) - load_pt(d_X, l.b);
            const float dist = norm3df(r.x, r.y, r.z);
            if (next & 1) {  // this cell is the b end
                fx += strength * r.x / dist;
                fy += strength * r.y / dist;
                fz += strength * r.z / dist;
            } else {
                fx += -strength * r.x / dist;
                fy += -strength * r.y / dist;
                fz += -strength * r.z / dist;
            }
        }
        out[0] = fx, out[1] = fy, out[2] = fz;
    }
}

template<typename Pt>
void segmented_link_forces(Links& links, const Stage_context& stage,
    const Pt* d_X, Pt* d_dX)
{
    links.reserve_segments(stage.n_max_cells);
    auto& seg = links.segments;
    const cudaStream_t s = stage.stream;
    const int sms = sm_count();
    const int link_blocks = stride_grid(links.n_max, 256, sms);
    const int n_tiles = ceil_div(stage.n_cells + 1, SCAN_TILE);
    bin_link_ends<<<link_blocks, 256, 0, s>>>(links.d_link, links.d_n,
        links.n_max, stage.n_cells, seg.count, seg.arrival);
    scan_bins<<<n_tiles, SCAN_THREADS, 0, s>>>(
        seg.count, seg.offset, n_tiles, seg.status, seg.ctl);
    place_link_ends<<<link_blocks, 256, 0, s>>>(links.d_link, links.d_n,
        links.n_max, seg.offset, seg.arrival, seg.entry);
    pull_on_cells<Pt><<<stride_grid(stage.n_cells, 128, sms), 128, 0, s>>>(
        stage.n_cells, d_X, links.d_link, seg.offset, seg.entry,
        links.strength, d_dX);
    YB_CUDA(cudaGetLastError());
}

}  // namespace yb


// Forces of all links on their ends, added to d_dX.
template<typename Pt, Link_force<Pt> force>
void link_forces(Links& links, const Pt* __restrict__ d_X, Pt* d_dX)
{
    const yb::Stage_context* stage = yb::current_stage();
    if (force == &linear_force<Pt> && stage != nullptr && stage->n_cells > 0) {
        yb::segmented_link_forces(links, *stage, d_X, d_dX);
        return;
    }
    const cudaStream_t s = stage ? stage->stream : 0;
    link<Pt, force>
        <<<yb::stride_grid(links.n_max, 128, yb::sm_count()), 128, 0, s>>>(
            d_X, d_dX, links.d_link, links.d_n, links.n_max, links.strength);
    YB_CUDA(cudaGetLastError());
}

template<typename Pt>
void link_forces(Links& links, const Pt* __restrict__ d_X, Pt* d_dX)
{
    link_forces<Pt, linear_force<Pt>>(links, d_X, d_dX);
}


// ---- walls ---------------------------------------------------------------------
// A wall is a plane normal to an axis whose position along that axis is carried
// by a "wall node", one of the cells. Cells interact with the wall through
// their distance to the plane (reference: links.cuh:142-228).
template<typename Pt>
using Wall_force = void(const Pt* __restrict__ d_X, const int i,
    const int wall_idx, Pt* d_dX, int* d_nints);

// One wall normal to the z axis.
template<typename Pt>
__device__ void xy_wall_relu_force(const Pt* __restrict__ d_X, const int i,
    const int wall_idx, Pt* d_dX, int* d_nints)
{
    const float z_wall = d_X[wall_idx].z;
    const float dist_wall = fabs(d_X[i].z - z_wall);
    if (dist_wall < 1.0f) {
        const auto F = fmaxf(0.8 - dist_wall, 0) - fmaxf(dist_wall - 0.8, 0);
        d_dX[i].z += F;

        atomicAdd(&d_dX[wall_idx].z, -F);
        atomicAdd(&d_nints[wall_idx], 1);
    }
}

template<typename Pt, Wall_force<Pt> force>
__global__ void wall(const Pt* __restrict__ d_X, Pt* d_dX, int n_max,
    int wall_idx, int* d_nints)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_max || i == wall_idx) return;
    force(d_X, i, wall_idx, d_dX, d_nints);
}

// Average the force on nodes that collected interactions. Launched <<<1, 1>>>
// like in the reference (links.cuh:208, :226), i.e. it normalises node 0.
template<typename Pt>
__global__ void update_wall_node(
    Pt* d_dX, int n_max, int wall_idx, int* d_nints)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_max) return;
    if (d_nints[i] > 0) {
        d_dX[i].x *= 1 / float(d_nints[i]);
        d_dX[i].y *= 1 / float(d_nints[i]);
        d_dX[i].z *= 1 / float(d_nints[i]);
    }
}

namespace yb {
// Interaction counters for the wall kernels: one zeroed int per possible node
// index, kept for the lifetime of the process (the reference cudaMallocs two
// ints per call and never frees them).
inline int* wall_counters(int n_entries, cudaStream_t s)
{
    static thread_local int* d_nints = nullptr;
    static thread_local int capacity = 0;
    if (n_entries > capacity) {
        if (d_nints) cudaFree(d_nints);
        capacity = n_entries;
        YB_CUDA(cudaMalloc(&d_nints, capacity * sizeof(int)));
    }
    YB_CUDA(cudaMemsetAsync(d_nints, 0, n_entries * sizeof(int), s));
    return d_nints;
}
}  // namespace yb

// Use this when there is a wall node, but no links.
template<typename Pt, Wall_force<Pt> force>
void wall_forces(
    const int n, const Pt* __restrict__ d_X, Pt* d_dX, const int wall_idx)
{
    const yb::Stage_context* stage = yb::current_stage();
    const cudaStream_t s = stage ? stage->stream : 0;
    const int entries = (wall_idx + 1 > 2 ? wall_idx + 1 : 2);
    int* d_nints = yb::wall_counters(entries, s);
    wall<Pt, force>
        <<<(n + 32 - 1) / 32, 32, 0, s>>>(d_X, d_dX, n, wall_idx, d_nints);
    update_wall_node<<<1, 1, 0, s>>>(d_dX, n, wall_idx, d_nints);
}

// Use this instead of link_forces when there is a wall node and links.
template<typename Pt, Link_force<Pt> l_force, Wall_force<Pt> w_force>
void link_wall_forces(Links& links, const int n, const Pt* __restrict__ d_X,
    Pt* d_dX, const int wall_idx)
{
    link_forces<Pt, l_force>(links, d_X, d_dX);
    wall_forces<Pt, w_force>(n, d_X, d_dX, wall_idx);
}
