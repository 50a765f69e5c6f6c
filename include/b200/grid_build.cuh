// Internal: building the neighbour grid -- cube ids, the bucket sort that puts
// cells into cube order, and the per-cube offsets.
//
// Reference semantics being reproduced (solvers.cuh:350-417): every cell gets
// the id of the cube of edge `cube_size` it sits in, cells are STABLY sorted
// by cube id (so inside a cube they appear by ascending original index), and
// cube_start/cube_end give the inclusive slot range of every cube.
//
// The reference gets there with compute_cube_id + thrust::sort_by_key over all
// 32 key bits + two n_cubes-sized fills + a boundary-detection kernel. Here the
// key space is tiny (n_cubes = grid_size^3 <= 2^24..2^27) and comparable to
// the number of cells, so the sort is a ONE-digit radix sort whose digit is
// the whole cube id:
//
//   1. bin_cells      key[i] = cube(X[i]); arrival[i] = atomicAdd(count[key], 1)
//   2. scan_bins      offset[c] = exclusive prefix sum of count (single pass,
//                     decoupled look-back); re-zeroes count for the next build
//   3. place_ids      slot_id[offset[key[i]] + arrival[i]] = i   (unstable:
//                     arrival order inside a cube depends on atomic timing)
//   4. reorder_cells  each slot ranks its id among the ids of its cube and
//                     moves the cell's STATE (not just an index) to
//                     offset[c] + rank  -> stable, deterministic cube order.
//
// Step 4 costs O(sum over cubes of m^2) id comparisons, which is always
// dominated by the pairwise sweep over the same cube (27 * m * m' distance
// tests), and it is where positions/extras/velocities are gathered from the
// user's AoS arrays into the cube-ordered planes of layout.cuh.
//
// offset[] (n_cubes + 1 entries) replaces cube_start/cube_end inside the
// solver: the cells of cubes a..b (inclusive) are slots
// [offset[a], offset[b + 1]), with no special case for empty cubes. The public
// Grid class still materialises the reference's four arrays bit-exactly.
#pragma once

#include <cuda_runtime.h>
#include <functional>
#include <vector>

#include "../cudebug.cuh"
#include "layout.cuh"

namespace yb {

inline int sm_count()
{
    static int n_sms = [] {
        int device = 0, count = 0;
        YB_CUDA(cudaGetDevice(&device));
        YB_CUDA(cudaDeviceGetAttribute(
            &count, cudaDevAttrMultiProcessorCount, device));
        return count;
    }();
    return n_sms;
}


// Make sure a kernel's code is resident (CUDA loads modules lazily, on the first
// launch, which can wait for the device to drain).
template<typename Kernel>
inline void load_kernel(Kernel kernel)
{
    cudaFuncAttributes attributes;
    YB_CUDA(cudaFuncGetAttributes(&attributes, kernel));
}


// ---- device-resident control block ---------------------------------------
// One per solver/grid; lives in device memory so that a captured CUDA graph
// can be replayed without the host knowing n, the scan epoch, or the drift.
struct Step_ctl {
    int scan_next_tile;   // dynamic tile ticket of the running scan
    int scan_tiles_done;  // completion counter of the running scan
    int scan_epoch;       // increments once per scan launch (validates status)
    int sweep_blocks_done;  // last-block election in the pairwise sweep
    int out_of_grid;      // cells whose cube id had to be clamped (diagnostic)
    int n_snapshot;       // n used by the step in flight (diagnostic)
    int list_overflow;    // a neighbour list of the split sweep was too short
    int n_ghosts;         // ghosts appended by the last halo round (diagnostic)
    float drift[2][4];    // per Heun stage: mean (or fixed-point) dX.xyz
    // Domain decomposition (b200/slab.cuh): cells with an id >= n_owned are
    // ghosts -- neighbours only; in force while external_drift is set.
    int n_owned;
    int external_drift;   // 1: decomposed run, the sweep leaves drift[] alone
    int pad2, pad3;
    float drift_sum[2][4];  // per stage: sum of dX.xyz over owned cells, count
};

// Set by the solver around the generic-forces callback, so that forces called
// from it (link_forces, wall_forces, ...) know how many cells there are, which
// stream the step runs on, and whether their launches are being recorded into
// a CUDA graph.
//
// A step with generic forces can only be replayed from a graph if the forces
// are "capturable" (Heun_solver::capture_generic_forces): they enqueue the same
// work on `stream` every time and take the live cell count from d_n_cells --
// n_cells is then n_max_cells, an upper bound. Work that must NOT be recorded
// (rebuilding a cache, allocating) goes to eager_stream, the stream the graph
// is launched on afterwards; a force may register a hook that runs on that
// stream before every replay and returns false if the recorded launches are
// no longer valid (the graph is then captured again).
using Replay_hook = std::function<bool(cudaStream_t)>;

struct Stage_context {
    int n_cells;
    int n_max_cells;
    cudaStream_t stream;
    const int* d_n_cells = nullptr;  // live count in device memory
    int stage = 0;                   // Heun stage: 0 predictor, 1 corrector
    bool capturing = false;
    const void* solver = nullptr;    // with step_serial: identifies the step
    unsigned long long step_serial = 0;
    cudaStream_t eager_stream = 0;
    std::vector<Replay_hook>* hooks = nullptr;
};

inline Stage_context*& current_stage()
{
    static thread_local Stage_context* context = nullptr;
    return context;
}

// ---- cube ids ---------------------------------------------------------------
// The solver's grid. By default the reference's: grid_size^3 cubes around the
// origin (solvers.cuh:357-360). One domain of a decomposed tissue restricts it
// to the box of cubes it can touch -- nx * ny * nz cubes starting at cube
// (grid_size / 2 - x_half, ...) of that cubic grid -- so that the per-cube
// tables, and the scan over them, stay proportional to the domain.
struct Grid_box {
    int nx, ny, nz;                // cubes along each axis
    int x_half, y_half, z_half;    // ix = floor(x / cube_size) + x_half, ...
    int n_cubes;                   // nx * ny * nz
    int restricted;                // 0: the reference's cubic grid

    static Grid_box cubic(int grid_size)
    {
        return Grid_box{grid_size, grid_size, grid_size, grid_size / 2,
            grid_size / 2, grid_size / 2, grid_size * grid_size * grid_size, 0};
    }
    bool operator==(const Grid_box& o) const
    {
        return nx == o.nx && ny == o.ny && nz == o.nz && x_half == o.x_half &&
               y_half == o.y_half && z_half == o.z_half;
    }
};

// Integer restatement of solvers.cuh:357-360. The reference evaluates
// floor(x / cs) + gs / 2 + (...) * gs + (...) * gs * gs in FP32, which is exact
// (hence equal to this) while all partial sums stay below 2^24, i.e. for
// grid_size <= 256. x / cs must stay an IEEE division: the
// cube of a cell within an ulp of a face depends on it.
// Ids outside [0, n_cubes) trip D_ASSERT in the reference; here they are
// clamped into the grid and counted. In a restricted box every axis is
// checked on its own (a cell beyond the box in x must not wrap into the next
// row); the cubic grid keeps the reference's linear rule.
__device__ __forceinline__ int cube_of(float x, float y, float z,
    float cube_size, const Grid_box& box, int* out_of_grid)
{
    long long ix = static_cast<long long>(floorf(x / cube_size)) + box.x_half;
    long long iy = static_cast<long long>(floorf(y / cube_size)) + box.y_half;
    long long iz = static_cast<long long>(floorf(z / cube_size)) + box.z_half;
    if (box.restricted) {
        const bool outside = ix < 0 || ix >= box.nx || iy < 0 || iy >= box.ny ||
                             iz < 0 || iz >= box.nz;
        if (outside) {
            if (out_of_grid) atomicAdd(out_of_grid, 1);
            ix = ix < 0 ? 0 : (ix >= box.nx ? box.nx - 1 : ix);
            iy = iy < 0 ? 0 : (iy >= box.ny ? box.ny - 1 : iy);
            iz = iz < 0 ? 0 : (iz >= box.nz ? box.nz - 1 : iz);
        }
    }
    long long id = ix + iy * box.nx + iz * box.nx * box.ny;
    if (id < 0 || id >= box.n_cubes) {
        if (out_of_grid) atomicAdd(out_of_grid, 1);  // nullptr: just recompute
        id = id < 0 ? 0 : box.n_cubes - 1;
    }
    return static_cast<int>(id);
}

// The faces of a brick of a decomposed tissue, moved inwards by the halo width:
// a cell outside [lo, hi) on an axis is within the halo of that face and has to
// be copied to the neighbour behind it (b200/domain.cuh). Bit 2a: below lo[a],
// bit 2a + 1: at or above hi[a].
struct Halo_faces {
    float lo[3], hi[3];
};

__device__ __forceinline__ unsigned char halo_flags_of(
    float x, float y, float z, const Halo_faces& faces)
{
    unsigned flags = 0;
    flags |= (x < faces.lo[0] ? 1u : 0u) | (x >= faces.hi[0] ? 2u : 0u);
    flags |= (y < faces.lo[1] ? 4u : 0u) | (y >= faces.hi[1] ? 8u : 0u);
    flags |= (z < faces.lo[2] ? 16u : 0u) | (z >= faces.hi[2] ? 32u : 0u);
    return static_cast<unsigned char>(flags);
}

__device__ __forceinline__ int live_cells(const int* d_n, int n_max)
{
    const int n = *d_n;
    return n < 0 ? 0 : (n > n_max ? n_max : n);
}

// Step 1, stand-alone form (the Heun predictor fuses the same two lines into
// its update kernel for the second stage).
template<typename Pt>
__global__ void __launch_bounds__(256) bin_cells(const int* __restrict__ d_n,
    int n_max, const Pt* __restrict__ d_X, float cube_size, Grid_box box,
    int* __restrict__ key, int* __restrict__ arrival, int* count, Step_ctl* ctl)
{
    const int n = live_cells(d_n, n_max);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += gridDim.x * blockDim.x) {
        const float* p = reinterpret_cast<const float*>(d_X + i);
        const int c = cube_of(__ldg(p), __ldg(p + 1), __ldg(p + 2), cube_size,
            box, &ctl->out_of_grid);
        key[i] = c;
        arrival[i] = atomicAdd(count + c, 1);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->n_snapshot = n;
}


// The same pass over a part of the cells of a decomposed tissue: part 0 = the
// cells this domain owns (before the ghosts of the stage have arrived, *d_n is
// not final then), part 1 = the ghosts behind them. Arrival numbers differ from
// one pass over all cells, the cube order built from them does not (in-cube
// order is by ascending index, see stable_rank).
template<typename Pt>
__global__ void __launch_bounds__(256) bin_cells_part(const int* __restrict__ d_n,
    int n_max, const Pt* __restrict__ d_X, float cube_size, Grid_box box,
    int* __restrict__ key, int* __restrict__ arrival, int* count, Step_ctl* ctl,
    int part)
{
    const int n_owned = min(ctl->n_owned, n_max);
    const int first = part == 0 ? 0 : n_owned;
    const int last = part == 0 ? n_owned : live_cells(d_n, n_max);
    for (int i = first + blockIdx.x * blockDim.x + threadIdx.x; i < last;
         i += gridDim.x * blockDim.x) {
        const float* p = reinterpret_cast<const float*>(d_X + i);
        const int c = cube_of(__ldg(p), __ldg(p + 1), __ldg(p + 2), cube_size,
            box, &ctl->out_of_grid);
        key[i] = c;
        arrival[i] = atomicAdd(count + c, 1);
    }
    if (part == 1 && blockIdx.x == 0 && threadIdx.x == 0) ctl->n_snapshot = last;
}


// ---- step 2: single-pass exclusive scan with decoupled look-back ------------
// Tiles of SCAN_TILE bins; tile ids are handed out dynamically so a tile only
// ever waits on tiles that are already running. A tile publishes
// (epoch, AGGREGATE, sum) as soon as it knows its own sum and upgrades to
// (epoch, PREFIX, inclusive prefix) once warp 0 has walked back to the nearest
// published prefix. The epoch makes stale words from earlier launches invalid,
// so the status array is never cleared.
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_SUB = 2;  // int4 vectors per thread
constexpr int SCAN_TILE = SCAN_THREADS * 4 * SCAN_SUB;
constexpr unsigned SCAN_AGGREGATE = 1u, SCAN_PREFIX = 2u;

__host__ __device__ constexpr int scan_padded(int n_entries)
{
    return ceil_div(n_entries, SCAN_TILE) * SCAN_TILE;
}

__device__ __forceinline__ unsigned long long scan_word(
    unsigned epoch, unsigned state, int value)
{
    return (static_cast<unsigned long long>((epoch << 2) | state) << 32) |
           static_cast<unsigned>(value);
}

__device__ __forceinline__ void scan_publish(
    unsigned long long* slot, unsigned long long word)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(slot), "l"(word)
                 : "memory");
}

__device__ __forceinline__ unsigned long long scan_peek(
    const unsigned long long* slot)
{
    unsigned long long word;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];"
                 : "=l"(word)
                 : "l"(slot)
                 : "memory");
    return word;
}

// Decoupled look-back of one warp: publish this tile's aggregate, walk back to
// the nearest tile that already knows its inclusive prefix, publish ours, and
// return the exclusive prefix of the tile (valid in every lane).
__device__ __forceinline__ int scan_lookback(unsigned long long* status,
    int tile, unsigned epoch, int aggregate, int lane_id)
{
    int exclusive = 0;
    if (tile == 0) {
        if (lane_id == 0)
            scan_publish(status, scan_word(epoch, SCAN_PREFIX, aggregate));
        return 0;
    }
    if (lane_id == 0)
        scan_publish(status + tile, scan_word(epoch, SCAN_AGGREGATE, aggregate));
    int look = tile - 1;  // lane 0 inspects `look`, lane l `look - l`
    while (true) {
        const int mine = look - lane_id;
        unsigned state = SCAN_PREFIX;
        int value = 0;
        if (mine >= 0) {
            unsigned long long word;
            do {
                word = scan_peek(status + mine);
                state = static_cast<unsigned>(word >> 32);
            } while ((state >> 2) != epoch || (state & 3u) == 0u);
            state &= 3u;
            value = static_cast<int>(static_cast<unsigned>(word));
        }
        // nearest predecessor that already knows its inclusive prefix
        const unsigned has_prefix =
            __ballot_sync(0xffffffffu, state == SCAN_PREFIX);
        // lanes up to and including it contribute; the whole window does if
        // nobody in it has a prefix yet
        const int stop = has_prefix ? __ffs(has_prefix) - 1 : 31;
        int contrib = (lane_id <= stop && mine >= 0) ? value : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1)
            contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
        exclusive += contrib;
        if (has_prefix != 0u) break;
        look -= 32;
    }
    if (lane_id == 0)
        scan_publish(
            status + tile, scan_word(epoch, SCAN_PREFIX, exclusive + aggregate));
    return exclusive;
}

// count: padded to scan_padded(n_entries), zero beyond n_entries; zeroed again
// on exit. offset: same padding; offset[c] = sum of count[0..c).
__global__ void __launch_bounds__(SCAN_THREADS) scan_bins(int* count,
    int* __restrict__ offset, int n_tiles, unsigned long long* status,
    Step_ctl* ctl)
{
    __shared__ int s_tile;
    __shared__ int s_warp_sum[SCAN_SUB][SCAN_THREADS / 32];
    __shared__ int s_tile_prefix;

    const int t = threadIdx.x;
    const int lane_id = t & 31, warp_id = t >> 5;
    if (t == 0) s_tile = atomicAdd(&ctl->scan_next_tile, 1);
    __syncthreads();
    const int tile = s_tile;
    const unsigned epoch =
        static_cast<unsigned>(*(volatile int*)&ctl->scan_epoch) & 0x3fffffffu;

    int4* tile_in = reinterpret_cast<int4*>(count + size_t(tile) * SCAN_TILE);
    int4 v[SCAN_SUB];
    int incl[SCAN_SUB];
#pragma unroll
    for (int u = 0; u < SCAN_SUB; u++) {
        v[u] = tile_in[u * SCAN_THREADS + t];
        tile_in[u * SCAN_THREADS + t] = make_int4(0, 0, 0, 0);
        int s = v[u].x + v[u].y + v[u].z + v[u].w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, s, d);
            if (lane_id >= d) s += up;
        }
        incl[u] = s;
        if (lane_id == 31) s_warp_sum[u][warp_id] = s;
    }
    __syncthreads();

    // exclusive prefix of this thread's vector inside the tile
    int before[SCAN_SUB];
    int running = 0;
#pragma unroll
    for (int u = 0; u < SCAN_SUB; u++) {
        int warps_before = 0, sub_total = 0;
#pragma unroll
        for (int w = 0; w < SCAN_THREADS / 32; w++) {
            const int ws = s_warp_sum[u][w];
            if (w < warp_id) warps_before += ws;
            sub_total += ws;
        }
        before[u] = running + warps_before + incl[u] -
                    (v[u].x + v[u].y + v[u].z + v[u].w);
        running += sub_total;
    }
    const int aggregate = running;

    if (warp_id == 0) {
        const int exclusive =
            scan_lookback(status, tile, epoch, aggregate, lane_id);
        if (lane_id == 0) s_tile_prefix = exclusive;
    }
    __syncthreads();
    const int tile_prefix = s_tile_prefix;

    int4* tile_out = reinterpret_cast<int4*>(offset + size_t(tile) * SCAN_TILE);
#pragma unroll
    for (int u = 0; u < SCAN_SUB; u++) {
        int4 o;
        o.x = tile_prefix + before[u];
        o.y = o.x + v[u].x;
        o.z = o.y + v[u].y;
        o.w = o.z + v[u].z;
        tile_out[u * SCAN_THREADS + t] = o;
    }

    // The last tile to finish re-arms the control words for the next launch.
    if (t == 0) {
        __threadfence();
        if (atomicAdd(&ctl->scan_tiles_done, 1) == n_tiles - 1) {
            ctl->scan_next_tile = 0;
            ctl->scan_tiles_done = 0;
            ctl->scan_epoch = static_cast<int>((epoch + 1u) & 0x3fffffffu);
            __threadfence();
        }
    }
}


// ---- step 3 -------------------------------------------------------------------
__global__ void __launch_bounds__(256) place_ids(const int* __restrict__ d_n,
    int n_max, const int* __restrict__ key, const int* __restrict__ arrival,
    const int* __restrict__ offset, int* __restrict__ slot_id)
{
    const int n = live_cells(d_n, n_max);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += gridDim.x * blockDim.x)
        slot_id[__ldg(offset + __ldg(key + i)) + __ldg(arrival + i)] = i;
}

// Rank of `id` among the ids that arrived in the same cube: the number of
// smaller ids. Gives ascending original index inside every cube, i.e. exactly
// the order a stable sort of (cube id, identity permutation) produces.
__device__ __forceinline__ int stable_rank(
    const int* __restrict__ slot_id, int start, int end, int id)
{
    int rank = 0;
    for (int q = start; q < end; q++) rank += (__ldg(slot_id + q) < id);
    return rank;
}

// ---- step 4, solver form: move state into cube order -------------------------
template<typename Pt>
__global__ void __launch_bounds__(256) reorder_cells(
    const int* __restrict__ d_n, int n_max, const Pt* __restrict__ d_X,
    const float3* __restrict__ d_old_v, const int* __restrict__ key,
    const int* __restrict__ offset, const int* __restrict__ slot_id,
    float4* __restrict__ pos4, float4* __restrict__ aux,
    int* __restrict__ cube_sorted)
{
    using L = Layout<Pt>;
    const int n = live_cells(d_n, n_max);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += gridDim.x * blockDim.x) {
        const int id = __ldg(slot_id + k);
        const int c = __ldg(key + id);
        const int start = __ldg(offset + c);
        const int end = __ldg(offset + c + 1);
        const int dst = start + stable_rank(slot_id, start, end, id);

        const Pt X = load_pt(d_X, id);
        pos4[dst] = make_float4(X.x, X.y, X.z, __int_as_float(id));

        float a[L::aux_lanes];
#pragma unroll
        for (int e = 0; e < L::extras; e++) a[e] = lane(X, 3 + e);
        const float* v = reinterpret_cast<const float*>(d_old_v + id);
        a[L::v_lane + 0] = __ldg(v + 0);
        a[L::v_lane + 1] = __ldg(v + 1);
        a[L::v_lane + 2] = __ldg(v + 2);
#pragma unroll
        for (int e = L::v_lane + 3; e < L::aux_lanes; e++) a[e] = 0.f;
#pragma unroll
        for (int q = 0; q < L::aux_vec4; q++)
            aux[size_t(dst) * L::aux_vec4 + q] = make_float4(
                a[4 * q], a[4 * q + 1], a[4 * q + 2], a[4 * q + 3]);

        cube_sorted[dst] = c;
    }
}

// ---- steps 3 + 4 for large tissues: carry the state along ---------------------
// place_ids + reorder_cells gather X, old_v and the key by original id: about
// four random 32-byte sectors per cell, which is what bounds the build once
// the state no longer fits the L2 (>= 10 M cells). The pair below reads the
// state in original order instead -- sequentially -- and scatters one packed
// record {x, y, z, bits(id), aux...} per cell to slot offset[cube] + arrival:
// cube order, but arrival order inside a cube. settle_cells then puts every
// cube into ascending id (ranks among 2-4 neighbouring records: sequential
// reads, near-sequential writes) and writes the planes the sweep reads.
template<typename Pt>
__global__ void __launch_bounds__(256) place_cells(
    const int* __restrict__ d_n, int n_max, const Pt* __restrict__ d_X,
    const float3* __restrict__ d_old_v, const int* __restrict__ key,
    const int* __restrict__ arrival, const int* __restrict__ offset,
    float4* __restrict__ staged)
{
    using L = Layout<Pt>;
    constexpr int REC = 1 + L::aux_vec4;  // float4s per staged record
    const int n = live_cells(d_n, n_max);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += gridDim.x * blockDim.x) {
        const int slot = __ldg(offset + __ldg(key + i)) + __ldg(arrival + i);
        const Pt X = load_pt(d_X, i);
        float4* record = staged + size_t(slot) * REC;
        record[0] = make_float4(X.x, X.y, X.z, __int_as_float(i));

        float a[L::aux_lanes];
#pragma unroll
        for (int e = 0; e < L::extras; e++) a[e] = lane(X, 3 + e);
        const float* v = reinterpret_cast<const float*>(d_old_v + i);
        a[L::v_lane + 0] = __ldg(v + 0);
        a[L::v_lane + 1] = __ldg(v + 1);
        a[L::v_lane + 2] = __ldg(v + 2);
#pragma unroll
        for (int e = L::v_lane + 3; e < L::aux_lanes; e++) a[e] = 0.f;
#pragma unroll
        for (int q = 0; q < L::aux_vec4; q++)
            record[1 + q] = make_float4(
                a[4 * q], a[4 * q + 1], a[4 * q + 2], a[4 * q + 3]);
    }
}

template<typename Pt>
__global__ void __launch_bounds__(256) settle_cells(
    const int* __restrict__ d_n, int n_max, const float4* __restrict__ staged,
    const int* __restrict__ offset, float cube_size, Grid_box box,
    float4* __restrict__ pos4, float4* __restrict__ aux,
    int* __restrict__ cube_sorted)
{
    using L = Layout<Pt>;
    constexpr int REC = 1 + L::aux_vec4;
    const int n = live_cells(d_n, n_max);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += gridDim.x * blockDim.x) {
        const float4 me = __ldg(staged + size_t(k) * REC);
        // same arithmetic as the binning, so the same cube
        const int c = cube_of(me.x, me.y, me.z, cube_size, box, nullptr);
        const int start = __ldg(offset + c), end = __ldg(offset + c + 1);
        const int id = __float_as_int(me.w);
        int rank = 0;
        for (int q = start; q < end; q++)
            rank += __float_as_int(__ldg(
                        reinterpret_cast<const float*>(staged + size_t(q) * REC) + 3)) < id;
        const int dst = start + rank;
        pos4[dst] = me;
#pragma unroll
        for (int q = 0; q < L::aux_vec4; q++)
            aux[size_t(dst) * L::aux_vec4 + q] =
                __ldg(staged + size_t(k) * REC + 1 + q);
        cube_sorted[dst] = c;
    }
}

// ---- step 4, public-Grid form: the reference's four arrays, bit for bit ------
__global__ void __launch_bounds__(256) fill_cube_ranges(
    int n_cubes, int* __restrict__ cube_start, int* __restrict__ cube_end)
{
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cubes;
         c += gridDim.x * blockDim.x) {
        cube_start[c] = -1;  // empty markers of solvers.cuh:411-412
        cube_end[c] = -2;
    }
}

__global__ void __launch_bounds__(256) publish_grid(
    const int* __restrict__ d_n, int n_max,
    const int* __restrict__ key, const int* __restrict__ offset,
    const int* __restrict__ slot_id, int* __restrict__ d_cube_id,
    int* __restrict__ d_point_id, int* __restrict__ d_cube_start,
    int* __restrict__ d_cube_end)
{
    const int n = live_cells(d_n, n_max);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += gridDim.x * blockDim.x) {
        const int id = __ldg(slot_id + k);
        const int c = __ldg(key + id);
        const int start = __ldg(offset + c);
        const int end = __ldg(offset + c + 1);
        const int rank = stable_rank(slot_id, start, end, id);
        d_cube_id[start + rank] = c;
        d_point_id[start + rank] = id;
        if (rank == 0) d_cube_start[c] = start;
        if (rank == end - start - 1) d_cube_end[c] = end - 1;  // inclusive
    }
}

// Launch width for grid-stride kernels over n_max elements: enough CTAs to
// fill the machine a few times over, never more than the data needs.
inline int stride_grid(int n_max, int threads, int n_sms, int ctas_per_sm = 32)
{
    const int wanted = ceil_div(n_max > 0 ? n_max : 1, threads);
    const int cap = n_sms * ctas_per_sm;
    return wanted < cap ? wanted : cap;
}

}  // namespace yb
