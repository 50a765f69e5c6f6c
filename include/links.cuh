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
// segmented reduction over a per-cell index of the link ends (a CSR):
//   build   bin_link_ends      bucket both ends of every live link by cell id
//           scan_bins          per-cell offsets (the grid's single-pass scan)
//           place_link_ends    (2 * link + side) entries into the cell's segment
//           order_link_ends    every segment into ascending link order (hubs,
//           order_hub_segments  cells with > 32 ends, by one CTA each)
//   pull    pull_on_cells      one thread per cell adds its entries' forces in
//           pull_on_hubs       ascending link order and updates d_dX once.
// The index is built once per step (the second Heun stage reuses the first
// stage's) or, with Links::cache_topology, once per change of the links. All
// counts stay on the device, every launch is capturable into a CUDA graph, and
// the result does not depend on the order in which atomics happen to land.
// Custom Link_force functors do their own (atomic) updates, exactly as in the
// reference.
#pragma once

#include <assert.h>
#include <curand_kernel.h>
#include <stdlib.h>
#include <time.h>
#include <functional>
#include <memory>

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


namespace yb {

// Per-cell index of the link ends (internal). Shared between the Links object
// and the replay hooks of captured steps, so it can outlive either.
struct Link_segments {
    int cell_capacity = 0;
    int n_tiles = 0;
    int* count = nullptr;    // per cell, zero between builds
    int* offset = nullptr;   // per cell (+1, + padding)
    int* arrival = nullptr;  // 2 per link (scratch of the hub sort afterwards)
    int* entry = nullptr;    // 2 per link: 2 * link + side, grouped by cell
    int* hubs = nullptr;     // cells with more than LINK_HUB ends
    int* hub_ctl = nullptr;  // [0] number of hubs, [1] links with a bad end
    unsigned long long* status = nullptr;
    Step_ctl* ctl = nullptr;
    // bookkeeping of the cache
    bool owner_alive = true;
    bool dirty = true;              // d_link changed since the last build
    const void* built_by = nullptr;  // solver + step of the last build
    unsigned long long built_in_step = 0;

    Link_segments() = default;
    Link_segments(const Link_segments&) = delete;
    Link_segments& operator=(const Link_segments&) = delete;
    ~Link_segments() { release(); }

    void reserve(int n_cells, int n_links_max)
    {
        if (n_cells <= cell_capacity) return;
        release();
        const size_t bins = static_cast<size_t>(scan_padded(n_cells + 1));
        const size_t ends = 2 * static_cast<size_t>(n_links_max > 0 ? n_links_max : 1);
        cell_capacity = n_cells;
        n_tiles = static_cast<int>(bins / SCAN_TILE);
        YB_CUDA(cudaMalloc(&count, bins * sizeof(int)));
        YB_CUDA(cudaMemset(count, 0, bins * sizeof(int)));
        YB_CUDA(cudaMalloc(&offset, bins * sizeof(int)));
        YB_CUDA(cudaMalloc(&arrival, ends * sizeof(int)));
        YB_CUDA(cudaMalloc(&entry, ends * sizeof(int)));
        // a hub has more than LINK_HUB = 32 ends, so there are fewer than
        // ends / 32 of them: the list cannot overflow
        YB_CUDA(cudaMalloc(&hubs, (ends / 32 + 1) * sizeof(int)));
        YB_CUDA(cudaMalloc(&hub_ctl, 4 * sizeof(int)));
        YB_CUDA(cudaMemset(hub_ctl, 0, 4 * sizeof(int)));
        YB_CUDA(cudaMalloc(&status, n_tiles * sizeof(unsigned long long)));
        YB_CUDA(cudaMemset(status, 0, n_tiles * sizeof(unsigned long long)));
        YB_CUDA(cudaMalloc(&ctl, sizeof(Step_ctl)));
        Step_ctl fresh{};
        fresh.scan_epoch = 1;
        YB_CUDA(cudaMemcpy(ctl, &fresh, sizeof(fresh), cudaMemcpyHostToDevice));
        dirty = true;
        built_by = nullptr;
    }
    void release()
    {
        cudaFree(ctl);
        cudaFree(status);
        cudaFree(hub_ctl);
        cudaFree(hubs);
        cudaFree(entry);
        cudaFree(arrival);
        cudaFree(offset);
        cudaFree(count);
        ctl = nullptr, status = nullptr, hub_ctl = nullptr, hubs = nullptr;
        entry = nullptr, arrival = nullptr, offset = nullptr, count = nullptr;
        cell_capacity = 0;
    }
};

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
        : n_max{n_max}, strength{strength},
          segments{std::make_shared<yb::Link_segments>()}
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
        segments->owner_alive = false;  // captured steps notice and re-capture
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
        mark_changed();
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
        mark_changed();
    }
    void copy_to_host()
    {
        YB_CUDA(cudaMemcpy(h_link, d_link,
            static_cast<size_t>(n_max) * sizeof(Link), cudaMemcpyDeviceToHost));
        YB_CUDA(cudaMemcpy(h_n, d_n, sizeof(int), cudaMemcpyDeviceToHost));
        assert(*h_n <= n_max);
    }

    // Extension: by default link_forces re-indexes the links once per solver
    // step, because model kernels rewire d_link on the device between steps.
    // With cache_topology = true the index is kept until the links change
    // through copy_to_device / set_d_n / reset or the model says so with
    // mark_changed() (e.g. after its own rewiring kernel).
    bool cache_topology = false;
    void mark_changed() { segments->dirty = true; }

    // Extension: links with an end outside [0, n_max of the cells) seen by the
    // last link_forces call inside a step (they are skipped). Blocks.
    int links_out_of_range()
    {
        if (segments->hub_ctl == nullptr) return 0;
        int bad = 0;
        YB_CUDA(cudaMemcpy(&bad, segments->hub_ctl + 1, sizeof(int),
            cudaMemcpyDeviceToHost));
        return bad;
    }

    std::shared_ptr<yb::Link_segments> segments;  // internal
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

constexpr int LINK_HUB = 32;        // segments longer than this go to one CTA each
constexpr int LINK_HUB_THREADS = 256;
constexpr int LINK_HUB_SORT = 4096;  // entries a CTA sorts in shared memory

// build 1: count the ends per cell and remember each end's arrival rank. Like
// the reference's kernel, every link with a != b pulls; a link with an end
// outside the cell arrays is skipped and counted (the reference would write
// out of bounds).
__global__ void __launch_bounds__(256) bin_link_ends(
    const Link* __restrict__ d_link, const int* __restrict__ d_n_links,
    int n_links_max, int n_cells_max, int* count, int* __restrict__ arrival,
    int* hub_ctl)
{
    const int n_links = live_cells(d_n_links, n_links_max);
    if (blockIdx.x == 0 && threadIdx.x == 0) hub_ctl[0] = 0;  // no hubs yet
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < n_links;
         l += gridDim.x * blockDim.x) {
        const int a = d_link[l].a, b = d_link[l].b;
        const bool in_range = a >= 0 && b >= 0 && a < n_cells_max &&
                              b < n_cells_max;
        if (a != b && !in_range) atomicAdd(hub_ctl + 1, 1);
        const bool active = a != b && in_range;
        arrival[2 * l] = active ? atomicAdd(count + a, 1) : -1;
        arrival[2 * l + 1] = active ? atomicAdd(count + b, 1) : -1;
    }
}

// build 3: entries of a cell, in arrival order
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

// build 4: ascending link order inside every segment, so that the sums below
// are accumulated in an order that does not depend on atomic timing. Short
// segments (a handful of entries) by insertion sort, one thread each; long
// ones are listed for order_hub_segments.
__global__ void __launch_bounds__(256) order_link_ends(int n_cells_max,
    const int* __restrict__ offset, int* entry, int* __restrict__ hubs,
    int* hub_ctl)
{
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cells_max;
         c += gridDim.x * blockDim.x) {
        const int start = __ldg(offset + c), end = __ldg(offset + c + 1);
        const int m = end - start;
        if (m < 2) continue;
        if (m > LINK_HUB) {
            hubs[atomicAdd(hub_ctl, 1)] = c;
            continue;
        }
        for (int k = start + 1; k < end; k++) {
            const int e = entry[k];
            int q = k - 1;
            while (q >= start && entry[q] > e) {
                entry[q + 1] = entry[q];
                q--;
            }
            entry[q + 1] = e;
        }
    }
}

// build 5: one CTA per hub. Up to LINK_HUB_SORT entries are sorted in shared
// memory (bitonic); longer segments by ranking every entry against all others
// into the (by now unused) arrival array and copying back.
__global__ void __launch_bounds__(LINK_HUB_THREADS) order_hub_segments(
    const int* __restrict__ offset, int* entry, int* scratch,
    const int* __restrict__ hubs, const int* __restrict__ hub_ctl)
{
    __shared__ int s_e[LINK_HUB_SORT];
    const int n_hubs = hub_ctl[0];
    for (int h = blockIdx.x; h < n_hubs; h += gridDim.x) {
        const int c = hubs[h];
        const int start = offset[c], m = offset[c + 1] - start;
        if (m <= LINK_HUB_SORT) {
            int padded = 1;
            while (padded < m) padded <<= 1;
            for (int q = threadIdx.x; q < padded; q += blockDim.x)
                s_e[q] = q < m ? entry[start + q] : 0x7fffffff;
            __syncthreads();
            for (int k = 2; k <= padded; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int q = threadIdx.x; q < padded; q += blockDim.x) {
                        const int partner = q ^ j;
                        if (partner > q) {
                            const bool up = (q & k) == 0;
                            const int x = s_e[q], y = s_e[partner];
                            if ((x > y) == up) s_e[q] = y, s_e[partner] = x;
                        }
                    }
                    __syncthreads();
                }
            }
            for (int q = threadIdx.x; q < m; q += blockDim.x)
                entry[start + q] = s_e[q];
        } else {
            for (int q = threadIdx.x; q < m; q += blockDim.x) {
                const int e = entry[start + q];
                int rank = 0;
                for (int p = 0; p < m; p++) rank += entry[start + p] < e;
                scratch[start + rank] = e;  // entries are distinct
            }
            __syncthreads();
            for (int q = threadIdx.x; q < m; q += blockDim.x)
                entry[start + q] = scratch[start + q];
        }
        __syncthreads();
    }
}

// -/+ strength * r / |r| on the a / b end of the link, like linear_force.
template<typename Pt>
__device__ __forceinline__ void add_link_pull(const Pt* __restrict__ d_X,
    const Link* __restrict__ d_link, int e, float strength, float& fx,
    float& fy, float& fz)
{
    const Link l = d_link[e >> 1];
    const Pt r = load_pt(d_X, l.a) - load_pt(d_X, l.b);
    const float dist = norm3df(r.x, r.y, r.z);
    if (e & 1) {  // this cell is the b end
        fx += strength * r.x / dist;
        fy += strength * r.y / dist;
        fz += strength * r.z / dist;
    } else {
        fx += -strength * r.x / dist;
        fy += -strength * r.y / dist;
        fz += -strength * r.z / dist;
    }
}

// pull 1: one thread per cell walks its (ordered) entries and updates d_dX once.
template<typename Pt>
__global__ void __launch_bounds__(128) pull_on_cells(int n_cells_max,
    const Pt* __restrict__ d_X, const Link* __restrict__ d_link,
    const int* __restrict__ offset, const int* __restrict__ entry,
    float strength, Pt* d_dX)
{
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cells_max;
         c += gridDim.x * blockDim.x) {
        const int start = __ldg(offset + c), end = __ldg(offset + c + 1);
        if (end <= start || end - start > LINK_HUB) continue;
        float* out = reinterpret_cast<float*>(d_dX + c);
        float fx = out[0], fy = out[1], fz = out[2];
        for (int q = start; q < end; q++)
            add_link_pull(d_X, d_link, __ldg(entry + q), strength, fx, fy, fz);
        out[0] = fx, out[1] = fy, out[2] = fz;
    }
}

// pull 2: one CTA per hub; thread t adds the entries t, t + 256, ... in that
// order and the partial sums are combined by a fixed tree.
template<typename Pt>
__global__ void __launch_bounds__(LINK_HUB_THREADS) pull_on_hubs(
    const Pt* __restrict__ d_X, const Link* __restrict__ d_link,
    const int* __restrict__ offset, const int* __restrict__ entry,
    const int* __restrict__ hubs, const int* __restrict__ hub_ctl,
    float strength, Pt* d_dX)
{
    __shared__ float s_f[3][LINK_HUB_THREADS];
    const int n_hubs = hub_ctl[0];
    for (int h = blockIdx.x; h < n_hubs; h += gridDim.x) {
        const int c = hubs[h];
        const int start = offset[c], end = offset[c + 1];
        float fx = 0.f, fy = 0.f, fz = 0.f;
        for (int q = start + threadIdx.x; q < end; q += blockDim.x)
            add_link_pull(d_X, d_link, entry[q], strength, fx, fy, fz);
        s_f[0][threadIdx.x] = fx;
        s_f[1][threadIdx.x] = fy;
        s_f[2][threadIdx.x] = fz;
        __syncthreads();
        for (int d = LINK_HUB_THREADS / 2; d > 0; d >>= 1) {
            if (threadIdx.x < d) {
                s_f[0][threadIdx.x] += s_f[0][threadIdx.x + d];
                s_f[1][threadIdx.x] += s_f[1][threadIdx.x + d];
                s_f[2][threadIdx.x] += s_f[2][threadIdx.x + d];
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            float* out = reinterpret_cast<float*>(d_dX + c);
            out[0] += s_f[0][0], out[1] += s_f[1][0], out[2] += s_f[2][0];
        }
        __syncthreads();
    }
}

// (Re)build the per-cell index of the link ends on stream s.
inline void index_link_ends(Links& links, Link_segments& seg, int n_cells_max,
    cudaStream_t s)
{
    const int sms = sm_count();
    const int link_blocks = stride_grid(links.n_max, 256, sms);
    const int n_tiles = ceil_div(n_cells_max + 1, SCAN_TILE);
    bin_link_ends<<<link_blocks, 256, 0, s>>>(links.d_link, links.d_n,
        links.n_max, n_cells_max, seg.count, seg.arrival, seg.hub_ctl);
    scan_bins<<<n_tiles, SCAN_THREADS, 0, s>>>(
        seg.count, seg.offset, n_tiles, seg.status, seg.ctl);
    place_link_ends<<<link_blocks, 256, 0, s>>>(links.d_link, links.d_n,
        links.n_max, seg.offset, seg.arrival, seg.entry);
    order_link_ends<<<stride_grid(n_cells_max, 256, sms), 256, 0, s>>>(
        n_cells_max, seg.offset, seg.entry, seg.hubs, seg.hub_ctl);
    order_hub_segments<<<32, LINK_HUB_THREADS, 0, s>>>(
        seg.offset, seg.entry, seg.arrival, seg.hubs, seg.hub_ctl);
    YB_CUDA(cudaGetLastError());
}

template<typename Pt>
void segmented_link_forces(Links& links, const Stage_context& stage,
    const Pt* d_X, Pt* d_dX)
{
    std::shared_ptr<Link_segments> seg = links.segments;
    const int n_cells_max = stage.n_max_cells;
    seg->reserve(n_cells_max, links.n_max);
    const cudaStream_t s = stage.stream;

    // Is the index of the last build still good?
    //  * cached topology: until the links are marked as changed;
    //  * otherwise: within one solver step (stage 1 reuses stage 0's index).
    // When a solver records its own step graph (hooks != nullptr) a cached
    // index is kept up to date OUTSIDE the graph, by the replay hook; inside a
    // caller's capture there is nobody to do that, so the per-step rule holds.
    const bool cached = links.cache_topology &&
                        !(stage.capturing && stage.hooks == nullptr);
    const bool same_step = stage.stage == 1 && seg->built_by == stage.solver &&
                           seg->built_in_step == stage.step_serial;
    if (cached) {
        if (seg->dirty) {
            index_link_ends(links, *seg, n_cells_max,
                stage.capturing ? stage.eager_stream : s);
            seg->dirty = false;
        }
    } else if (!same_step || stage.solver == nullptr) {
        index_link_ends(links, *seg, n_cells_max, s);
        seg->dirty = true;  // good for this step only
    }
    seg->built_by = stage.solver;
    seg->built_in_step = stage.step_serial;

    const int sms = sm_count();
    pull_on_cells<Pt><<<stride_grid(n_cells_max, 128, sms), 128, 0, s>>>(
        n_cells_max, d_X, links.d_link, seg->offset, seg->entry,
        links.strength, d_dX);
    pull_on_hubs<Pt><<<32, LINK_HUB_THREADS, 0, s>>>(d_X, links.d_link,
        seg->offset, seg->entry, seg->hubs, seg->hub_ctl, links.strength, d_dX);
    YB_CUDA(cudaGetLastError());

    if (stage.capturing && stage.hooks != nullptr && stage.stage == 0) {
        // What the recorded launches depend on: the Links object, its
        // strength, and the caching mode. A cached index is refreshed here,
        // on the stream the graph is about to be launched on.
        Links* owner = &links;
        const float strength = links.strength;
        const bool was_cached = cached;
        stage.hooks->push_back(
            [seg, owner, strength, was_cached, n_cells_max](cudaStream_t launch) {
                if (!seg->owner_alive) return false;
                if (owner->strength != strength ||
                    owner->cache_topology != was_cached)
                    return false;
                if (was_cached && seg->dirty) {
                    index_link_ends(*owner, *seg, n_cells_max, launch);
                    seg->dirty = false;
                }
                return true;
            });
    }
}

}  // namespace yb


// Forces of all links on their ends, added to d_dX.
template<typename Pt, Link_force<Pt> force>
void link_forces(Links& links, const Pt* __restrict__ d_X, Pt* d_dX)
{
    const yb::Stage_context* stage = yb::current_stage();
    if (force == &linear_force<Pt> && stage != nullptr && stage->n_max_cells > 0) {
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
