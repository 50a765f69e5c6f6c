// Internal: the pairwise-interaction kernels.
//
// sweep_cubes  -- Grid solver: for every cell, all cells in the 27 surrounding
//                 cubes closer than cube_size (reference: compute_cube,
//                 solvers.cuh:432-463, with add_rhs :147-161 and the zero
//                 fills :232-234 fused in).
// sweep_tiles  -- Tile solver: all pairs (reference: compute_tile :286-322).
//
// Semantics kept from the reference:
//  * one thread owns one cell i for the whole sweep, so user functors may
//    update per-i counters without atomics;
//  * the functor sees ORIGINAL cell ids and r = Xi - Xj over all lanes of Pt;
//  * Grid: a pair is handed to the functor iff !(norm3df(r) >= cube_size),
//    including the self pair (dist == 0), visited in the reference's order:
//    the 9 (y,z) rows in d_nhood order, ascending slot inside a row -- so the
//    per-cell sums are accumulated in the same order as the reference's;
//  * dX[i] (+)= F, then dX.xyz += sum_v / sum_friction if sum_friction > 0.
//
// What is different is how the data gets to the ALUs. Cells are physically in
// cube order (layout.cuh), so for a CTA working on 128 consecutive slots the
// candidates of each of the 9 rows form ONE contiguous span of pos4[] -- cube
// ids are linear with x fastest, so cubes c-1, c, c+1 are adjacent, and so are
// the rows of neighbouring cells. Per chunk the CTA:
//   1. looks up the 9 spans in offset[] (18 loads for the whole CTA),
//   2. pulls them into shared memory with 9 cp.async.bulk copies tracked by
//      one mbarrier (1-D TMA; pos4 records are 16 B so every span is aligned),
//   3. phase 1: every thread scans its own sub-ranges of the staged spans with
//      a cheap squared-distance test and appends the survivors to a short
//      per-thread list in shared memory,
//   4. phase 2: every thread walks its list: exact norm3df cut-off, functor,
//      friction. Separating the phases means a warp executes the (expensive,
//      user-defined) functor for ~max-over-lanes(#neighbours) iterations
//      instead of ~#candidates -- about 16 instead of 65 in a relaxed tissue.
// Spans longer than the staging buffer are consumed in several rounds and
// lists longer than LIST_CAP in several batches; neither changes the order.
//
// The kernel is persistent (grid = SMs x resident CTAs, chunks handed out
// round-robin), so each CTA can keep a running sum of dX for the drift
// correction; the last CTA to finish adds the per-CTA partials in a fixed
// order. That replaces thrust::reduce + the device->host copy of
// solvers.cuh:242 and keeps the result independent of scheduling.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "grid_build.cuh"
#include "layout.cuh"

namespace yb {

// Sizes of the staging window, of the per-thread list, and the CTAs per SM the
// register budget is set for -- chosen per point type (measured on B200 at 1 M
// cells, profiles/r01_sweep_tuning.md):
//  * plain positions (float3/float4, cheap functors): 1024 / 24 and 8 resident
//    CTAs -- the sweep is issue bound and more warps help (step 0.538 -> 0.503);
//  * points with extra lanes (polarities, concentrations: heavy functors that
//    gather their own per-cell arrays): 1280 / 24 and 6 CTAs, which leaves
//    ~60 KB of the SM's 256 KB as L1 for those gathers (growth step 2.25 ms
//    with 1536 / 32 -> 1.82 ms; 8 CTAs: 2.21 ms).
constexpr int SWEEP_THREADS = 128;
constexpr int SWEEP_ROWS = 9;

#ifndef YB_SWEEP_HEAVY_STAGE  // tuning overrides (profiles/r01_sweep_tuning.md)
#define YB_SWEEP_HEAVY_STAGE 1280
#endif
#ifndef YB_SWEEP_HEAVY_LIST
#define YB_SWEEP_HEAVY_LIST 24
#endif
#ifndef YB_SWEEP_HEAVY_CTAS
#define YB_SWEEP_HEAVY_CTAS 6
#endif

template<int LANES>
struct Sweep_config {
    static constexpr int stage_cap =
        LANES <= 4 ? 1024 : YB_SWEEP_HEAVY_STAGE;  // <= 4095
    static constexpr int list_cap = LANES <= 4 ? 24 : YB_SWEEP_HEAVY_LIST;
    static constexpr int min_ctas = LANES <= 4 ? 8 : YB_SWEEP_HEAVY_CTAS;
    static constexpr size_t smem =
        size_t(stage_cap) * sizeof(float4) +
        size_t(SWEEP_ROWS) * SWEEP_THREADS * sizeof(uint32_t) +
        size_t(list_cap) * SWEEP_THREADS * sizeof(uint16_t);
};
// Squared-distance pre-filter: keep everything the exact test could accept.
// The rounding error of dx*dx+dy*dy+dz*dz is a few 1e-7 relative; 1e-5 is
// generous and costs practically no extra list entries.
constexpr float SWEEP_PREFILTER_SLACK = 1.00001f;

// Cube trimming (sweep_cubes): squared gaps are compared against this, in units
// of cube_size^2; the gaps themselves are shortened by SWEEP_TRIM_MARGIN cube
// sizes, far more than the rounding of x / cube_size at |x| <= 512 cubes.
constexpr float SWEEP_TRIM_LIMIT = 1.0001f;
constexpr float SWEEP_TRIM_MARGIN = 1e-3f;

// ---- mbarrier / bulk-copy primitives (PTX; SASS: SYNCS.*, UBLKCP) ------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(arrivals)
                 : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// global -> shared, 16-byte granular, completion counted on the mbarrier
__device__ __forceinline__ void bulk_g2s(
    void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Entries in a thread's list column, from the address of its next free entry.
__device__ __forceinline__ int listed_at(uint32_t next_entry, uint32_t column)
{
    return int(next_entry - column) / int(SWEEP_THREADS * sizeof(uint16_t));
}

// Offset (in cube ids) of neighbour row r = 0..8 relative to a cell's cube.
// Same ordering as the reference's d_nhood table (solvers.cuh:472-484): the
// y-shift cycles 0, -1, +1 fastest, the z-shift 0, -1, +1 slowest; inside a row
// x runs -1, 0, +1, which is simply "the three adjacent ids".
__device__ __forceinline__ int row_shift(int r, const Grid_box& box)
{
    const int ry = r % 3, rz = r / 3;
    const int dy = ry == 0 ? 0 : (ry == 1 ? -1 : 1);
    const int dz = rz == 0 ? 0 : (rz == 1 ? -1 : 1);
    return dy * box.nx + dz * box.nx * box.ny;
}

// Cube ids stay far below 2^30 (grid_size <= 1024), so id +- one z-layer fits
// an int.
__device__ __forceinline__ int clamp_cube(int c, int n_cubes)
{
    return c < 0 ? 0 : (c > n_cubes ? n_cubes : c);
}

// Squared distance (in cube sizes) from a cell to the neighbouring cubes along
// each axis, as lower bounds: [axis][0] = 0 (same layer), [1] towards -1, [2]
// towards +1. The fractions come from the very quotients cube_of() floors, so
// they are consistent with the binning. A cell whose cube id was clamped (out
// of the grid) is not where its id says: no trimming for it.
__device__ __forceinline__ void trim_gaps(const float4& me, float cube_size,
    const Grid_box& box, int my_cube, float (*gap2)[3])
{
    const float q[3] = {me.x / cube_size, me.y / cube_size, me.z / cube_size};
    float fl[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        fl[c] = floorf(q[c]);
        const float f = q[c] - fl[c];
        const float below = fmaxf(f - SWEEP_TRIM_MARGIN, 0.f);
        const float above = fmaxf(1.f - f - SWEEP_TRIM_MARGIN, 0.f);
        gap2[c][0] = 0.f;
        gap2[c][1] = below * below;
        gap2[c][2] = above * above;
    }
    const long long id = (static_cast<long long>(fl[0]) + box.x_half) +
                         (static_cast<long long>(fl[1]) + box.y_half) * box.nx +
                         (static_cast<long long>(fl[2]) + box.z_half) *
                             box.nx * box.ny;
    if (id != my_cube) {
#pragma unroll
        for (int c = 0; c < 3; c++) gap2[c][1] = gap2[c][2] = 0.f;
    }
}

template<typename Pt>
__device__ __forceinline__ Pt assemble_pt(
    const float4& pos, const float4* __restrict__ aux_of_cell)
{
    using L = Layout<Pt>;
    Pt X;
    lane(X, 0) = pos.x;
    lane(X, 1) = pos.y;
    lane(X, 2) = pos.z;
    if (L::extras > 0) {
        float a[L::aux_lanes];
#pragma unroll
        for (int q = 0; q < ceil_div(L::extras, 4); q++) {
            const float4 v = __ldg(aux_of_cell + q);
            a[4 * q] = v.x, a[4 * q + 1] = v.y;
            a[4 * q + 2] = v.z, a[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int e = 0; e < L::extras; e++) lane(X, 3 + e) = a[e];
    }
    return X;
}

// |r| exactly as the reference computes it (norm3df), except that the self pair
// -- r = 0, one entry of every cell's list -- does not drag its warp through
// norm3df's out-of-range path (a divergent call for one lane in about 40 % of
// the warp iterations): it is handed a unit vector and the result replaced by
// the 0 norm3df would have returned.
__device__ __forceinline__ float pair_distance(float rx, float ry, float rz, bool self)
{
    const float d = norm3df(self ? 1.f : rx, ry, rz);
    return self ? 0.f : d;
}

template<typename Pt>
__device__ __forceinline__ float3 velocity_of(
    const float4* __restrict__ aux_of_cell)
{
    using L = Layout<Pt>;
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const int l = L::v_lane + c;
        const float4 q = __ldg(aux_of_cell + l / 4);
        v[c] = (l % 4 == 0) ? q.x : (l % 4 == 1) ? q.y : (l % 4 == 2) ? q.z : q.w;
    }
    return float3{v[0], v[1], v[2]};
}

// Deterministic CTA-wide sum of three floats; result valid in thread 0.
template<int THREADS>
__device__ __forceinline__ float3 block_sum3(
    float x, float y, float z, float (*s_red)[THREADS / 32])
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        x += __shfl_xor_sync(0xffffffffu, x, d);
        y += __shfl_xor_sync(0xffffffffu, y, d);
        z += __shfl_xor_sync(0xffffffffu, z, d);
    }
    const int warp_id = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        s_red[0][warp_id] = x, s_red[1][warp_id] = y, s_red[2][warp_id] = z;
    }
    __syncthreads();
    float3 total{0.f, 0.f, 0.f};
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < THREADS / 32; w++) {
            total.x += s_red[0][w], total.y += s_red[1][w], total.z += s_red[2][w];
        }
    }
    return total;
}

// How the drift of a stage is chosen (Heun_solver::set_fixed*).
enum Drift_mode : int {
    DRIFT_MEAN = 0,         // centre of mass stays put
    DRIFT_POINT = 1,        // one cell stays put
    DRIFT_POINT_XY_MEAN_Z = 2  // set_fixed_xy, first stage only
};

// Called by the last CTA of a sweep, after every dX of the stage is written.
template<typename Pt>
__device__ __forceinline__ void publish_drift(float3 mean, int mode,
    int fix_point, const Pt* d_dX, int stage, Step_ctl* ctl)
{
    float3 drift = mean;
    if (mode != DRIFT_MEAN) {
        const float* p = reinterpret_cast<const float*>(d_dX + fix_point);
        drift.x = __ldcg(p + 0);
        drift.y = __ldcg(p + 1);
        if (mode == DRIFT_POINT) drift.z = __ldcg(p + 2);
    }
    ctl->drift[stage][0] = drift.x;
    ctl->drift[stage][1] = drift.y;
    ctl->drift[stage][2] = drift.z;
}

// Last CTA standing: add up the per-CTA partial sums in CTA order and publish
// the drift for this Heun stage. sum / n follows the reference's operator/=
// (dtypes.cuh:204-208): multiply by float(1.0 / double(float(n))).
template<int THREADS, typename Pt>
__device__ __forceinline__ void finish_drift(float3 my_partial,
    float* __restrict__ partials, int n, int stage, int drift_mode,
    int fix_point, const Pt* d_dX, Step_ctl* ctl,
    float (*s_red)[THREADS / 32])
{
    __shared__ bool s_is_last;
    if (threadIdx.x == 0) {
        partials[3 * blockIdx.x + 0] = my_partial.x;
        partials[3 * blockIdx.x + 1] = my_partial.y;
        partials[3 * blockIdx.x + 2] = my_partial.z;
        __threadfence();
        s_is_last = atomicAdd(&ctl->sweep_blocks_done, 1) == int(gridDim.x) - 1;
    }
    __syncthreads();
    if (!s_is_last) return;
    __threadfence();
    float x = 0.f, y = 0.f, z = 0.f;
    for (int b = threadIdx.x; b < int(gridDim.x); b += THREADS) {
        x += __ldcg(partials + 3 * b + 0);
        y += __ldcg(partials + 3 * b + 1);
        z += __ldcg(partials + 3 * b + 2);
    }
    const float3 total = block_sum3<THREADS>(x, y, z, s_red);
    if (threadIdx.x == 0) {
        // with ghosts present only the owned cells count
        const int owned = ctl->external_drift ? min(ctl->n_owned, n) : n;
        ctl->drift_sum[stage][0] = total.x;
        ctl->drift_sum[stage][1] = total.y;
        ctl->drift_sum[stage][2] = total.z;
        ctl->drift_sum[stage][3] = static_cast<float>(owned);
        if (!ctl->external_drift) {
            const float inv_n =
                static_cast<float>(1. / static_cast<float>(owned));
            float3 mean{0.f, 0.f, 0.f};
            if (owned > 0)
                mean = float3{total.x * inv_n, total.y * inv_n, total.z * inv_n};
            publish_drift(mean, owned > 0 ? drift_mode : DRIFT_MEAN, fix_point,
                d_dX, stage, ctl);
        }
        ctl->sweep_blocks_done = 0;
        ctl->list_overflow = 0;  // (the fused kernel just took the stage over)
    }
}


// ---- Grid solver sweep -------------------------------------------------------
// SEEDED: d_dX already holds the generic forces of this stage and is added to;
// otherwise d_dX is write-only (no zero fill anywhere).
template<typename Pt, Pt (*pw_int)(Pt, Pt, float, int, int),
    float (*pw_friction)(Pt, Pt, float, int, int), bool SEEDED>
__global__ void __launch_bounds__(
    SWEEP_THREADS, Sweep_config<Layout<Pt>::lanes>::min_ctas) sweep_cubes(
    const int* __restrict__ d_n, int n_max, const float4* __restrict__ pos4,
    const float4* __restrict__ aux, const int* __restrict__ cube_sorted,
    const int* __restrict__ offset, float cube_size, Grid_box box, Pt* d_dX,
    float* __restrict__ partials, int stage, int drift_mode, int fix_point,
    Step_ctl* ctl, int only_if_overflow)
{
    using L = Layout<Pt>;
    // behind list_cubes + interact_lists: only if a neighbour list overflowed
    if (only_if_overflow && *(volatile int*)&ctl->list_overflow == 0) return;
    constexpr int SWEEP_STAGE_CAP = Sweep_config<L::lanes>::stage_cap;
    constexpr int SWEEP_LIST_CAP = Sweep_config<L::lanes>::list_cap;
    extern __shared__ __align__(16) unsigned char sweep_smem[];
    float4* s_pos = reinterpret_cast<float4*>(sweep_smem);
    uint32_t* s_range = reinterpret_cast<uint32_t*>(s_pos + SWEEP_STAGE_CAP);
    uint16_t* s_list =
        reinterpret_cast<uint16_t*>(s_range + SWEEP_ROWS * SWEEP_THREADS);
    __shared__ int s_row_lo[SWEEP_ROWS];     // first global slot of row span
    __shared__ int s_row_v[SWEEP_ROWS + 1];  // start in the concatenated spans
    __shared__ float s_red[3][SWEEP_THREADS / 32];
    __shared__ __align__(8) uint64_t s_bar;

    const int t = threadIdx.x;
    const int n = live_cells(d_n, n_max);
    const int n_chunks = ceil_div(n, SWEEP_THREADS);
    const float reach2 = cube_size * cube_size * SWEEP_PREFILTER_SLACK;
    const int n_owned = ctl->external_drift ? ctl->n_owned : n_max;

    if (t == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    uint32_t parity = 0;
    float3 cta_sum{0.f, 0.f, 0.f};

    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int first_slot = chunk * SWEEP_THREADS;
        const int k = first_slot + t;
        const bool live = k < n;

        if (t < SWEEP_ROWS) {
            const int last_slot = min(first_slot + SWEEP_THREADS, n) - 1;
            const int shift = row_shift(t, box);
            const int lo = __ldg(offset +
                clamp_cube(__ldg(cube_sorted + first_slot) + shift - 1, box.n_cubes));
            const int hi = __ldg(offset +
                clamp_cube(__ldg(cube_sorted + last_slot) + shift + 2, box.n_cubes));
            s_row_lo[t] = lo;
            s_row_v[t + 1] = hi > lo ? hi - lo : 0;  // length, scanned below
        }
        float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
        int my_cube = 0;
        if (live) {
            me = __ldg(pos4 + k);
            my_cube = __ldg(cube_sorted + k);
        }
        // This thread's candidate slots per row, as global slot numbers. The 18
        // loads are independent and fly while the spans are being staged.
        // Cubes that lie entirely beyond the cut-off are trimmed: a cell sits at
        // fraction f of its cube, so everything in the cube at offset -1 (+1)
        // along an axis is at least f (1 - f) cube sizes away along that axis.
        // That drops about a quarter of the candidates (corner cubes half of
        // the time, edge cubes a fifth) and never a pair the exact test accepts.
        float gap2[3][3];  // [axis][0: same, 1: -1, 2: +1], in cube_size^2
        trim_gaps(me, cube_size, box, my_cube, gap2);
        int my_lo[SWEEP_ROWS], my_hi[SWEEP_ROWS];
#pragma unroll
        for (int r = 0; r < SWEEP_ROWS; r++) {
            const int c = my_cube + row_shift(r, box);
            const float row_gap2 = gap2[1][r % 3] + gap2[2][r / 3];
            const bool row_out = !live || row_gap2 >= SWEEP_TRIM_LIMIT;
            const int first = row_gap2 + gap2[0][1] >= SWEEP_TRIM_LIMIT ? c : c - 1;
            const int last = row_gap2 + gap2[0][2] >= SWEEP_TRIM_LIMIT ? c + 1 : c + 2;
            my_lo[r] = row_out ? 0 : __ldg(offset + clamp_cube(first, box.n_cubes));
            my_hi[r] = row_out ? 0 : __ldg(offset + clamp_cube(last, box.n_cubes));
        }
        __syncthreads();
        if (t == 0) {
            int v = 0;
            s_row_v[0] = 0;
#pragma unroll
            for (int r = 0; r < SWEEP_ROWS; r++) {
                v += s_row_v[r + 1];
                s_row_v[r + 1] = v;
            }
        }
        __syncthreads();
        const int total = s_row_v[SWEEP_ROWS];

        const int my_id = __float_as_int(me.w);
        // ghost cells (domain decomposition) are candidates, never subjects
        const bool owned = live && my_id < n_owned;
        Pt Xi{0};
        if (owned) Xi = assemble_pt<Pt>(me, aux + size_t(k) * L::aux_vec4);
        Pt F{0};
        float3 sum_v{0.f, 0.f, 0.f};
        float sum_friction = 0.f;

        for (int v0 = 0; v0 < total; v0 += SWEEP_STAGE_CAP) {
            // -- stage the window [v0, v0 + CAP) of the concatenated spans
            if (v0 > 0) __syncthreads();  // everyone is done with the old window
            if (t == 0) {
                const int v1 = min(v0 + SWEEP_STAGE_CAP, total);
                mbar_expect_tx(&s_bar, uint32_t(v1 - v0) * sizeof(float4));
#pragma unroll 1
                for (int r = 0; r < SWEEP_ROWS; r++) {
                    const int a = max(s_row_v[r], v0);
                    const int b = min(s_row_v[r + 1], v1);
                    if (b > a)
                        bulk_g2s(s_pos + (a - v0),
                            pos4 + (s_row_lo[r] + (a - s_row_v[r])),
                            uint32_t(b - a) * sizeof(float4), &s_bar);
                }
            }
            // window-relative [a, b) per row, packed a | b << 16, parked in
            // shared memory so the scan below can index rows dynamically
#pragma unroll
            for (int r = 0; r < SWEEP_ROWS; r++) {
                const int base = s_row_v[r] - s_row_lo[r] - v0;
                const int a = min(max(my_lo[r] + base, 0), SWEEP_STAGE_CAP);
                const int b = min(max(my_hi[r] + base, a), SWEEP_STAGE_CAP);
                s_range[r * SWEEP_THREADS + t] = uint32_t(a) | (uint32_t(b) << 16);
            }
            mbar_wait(&s_bar, parity);
            parity ^= 1u;

            // -- phase 1 (scan) and phase 2 (interact). The lanes of a warp
            //    walk the 9 rows together (cells that are not subjects have
            //    empty ranges). A lane stops scanning when its list is full;
            //    if any lane of the warp could not finish its row, the whole
            //    warp interacts with what is listed and resumes where it
            //    stopped, so the order of the pairs never changes.
            const uint32_t list_begin = smem_u32(s_list + t);
            const uint32_t list_full =
                list_begin + SWEEP_LIST_CAP * SWEEP_THREADS * sizeof(uint16_t);
            int r = 0;
            uint32_t range = owned ? s_range[t] : 0u;
            int a = int(range & 0xffffu), b = int(range >> 16);
            while (true) {
                // shared-memory address of the next free entry of this thread's
                // list column (a plain 32-bit register: one add per hit)
                uint32_t la = list_begin;
                while (true) {
                    const float4* pp = s_pos + a;
                    // the list entry (row << 12 | position) is the loop counter
                    uint32_t entry = (uint32_t(r) << 12) | uint32_t(a);
                    const uint32_t entry_end = entry + uint32_t(b - a);
#pragma unroll 2
                    for (; entry < entry_end && la != list_full; entry++, pp++) {
                        const float4 p = *pp;
                        const float dx = me.x - p.x, dy = me.y - p.y,
                                    dz = me.z - p.z;
                        const float d2 = dx * dx + dy * dy + dz * dz;
                        if (!(d2 > reach2)) {
                            asm volatile("st.shared.u16 [%0], %1;" ::"r"(la),
                                         "h"(uint16_t(entry))
                                         : "memory");
                            la += SWEEP_THREADS * sizeof(uint16_t);
                        }
                    }
                    a = int(entry & 4095u);
                    if (__any_sync(0xffffffffu, a < b)) break;  // list(s) full
                    // (the next range is fetched before the exit test on
                    // purpose: with the test first, ptxas 12.9 replaces r in
                    // the address by the exit value and reads row 9)
                    ++r;
                    range = owned && r < SWEEP_ROWS
                                ? s_range[r * SWEEP_THREADS + t]
                                : 0u;
                    a = int(range & 0xffffu), b = int(range >> 16);
                    if (r == SWEEP_ROWS) break;
                }
                const int listed = listed_at(la, list_begin);

                for (int e = 0; e < listed; e++) {
                    const unsigned entry = s_list[e * SWEEP_THREADS + t];
                    const int row = entry >> 12, at = entry & 4095;
                    const int kj = s_row_lo[row] + (v0 + at - s_row_v[row]);
                    const float4* aux_j = aux + size_t(kj) * L::aux_vec4;
                    // issued early so the L2 round trip overlaps the functor
                    const float3 vj = velocity_of<Pt>(aux_j);
                    const float4 pj = s_pos[at];
                    const Pt Xj = assemble_pt<Pt>(pj, aux_j);
                    const Pt rij = Xi - Xj;
                    const int j_id = __float_as_int(pj.w);
                    const float dist =
                        pair_distance(rij.x, rij.y, rij.z, j_id == my_id);
                    if (dist >= cube_size) continue;

                    F += pw_int(Xi, rij, dist, my_id, j_id);
                    const float friction = pw_friction(Xi, rij, dist, my_id, j_id);
                    sum_friction += friction;
                    if (friction != 0.f) sum_v += friction * vj;
                }
                if (r == SWEEP_ROWS) break;
            }
        }

        // -- epilogue: dX (+)= F, friction term, drift partial
        float3 mine{0.f, 0.f, 0.f};
        if (owned) {
            Pt dX = F;
            if (SEEDED) {
                dX = load_pt_rw(d_dX, my_id);
                dX += F;
            }
            if (sum_friction > 0) {
                dX.x += sum_v.x / sum_friction;
                dX.y += sum_v.y / sum_friction;
                dX.z += sum_v.z / sum_friction;
            }
            store_pt(d_dX, my_id, dX);
            mine = float3{dX.x, dX.y, dX.z};
        }
        const float3 chunk_sum =
            block_sum3<SWEEP_THREADS>(mine.x, mine.y, mine.z, s_red);
        cta_sum.x += chunk_sum.x, cta_sum.y += chunk_sum.y, cta_sum.z += chunk_sum.z;
    }

    finish_drift<SWEEP_THREADS>(cta_sum, partials, n, stage, drift_mode,
        fix_point, d_dX, ctl, s_red);
}


// ---- Grid solver sweep in two kernels, for points with extra lanes --------------
// Heavy functors (polarity forces, functors that gather their own per-cell
// arrays) spend the sweep waiting: on dependent memory round trips of the user
// code and on long FMA/XU chains. What hides that is more warps in flight and
// more L1, and the fused kernel above has neither to give -- its registers hold
// the scan state next to the functor's, and its staging windows take most of
// the SM's shared memory/L1. So for these point types the two phases run as two
// kernels:
//
//   list_cubes      phase 1 alone (no Pt, no functor: ONE instantiation for all
//                   models): the staged candidate scan, whose survivors are
//                   written to a neighbour list in global memory -- row e of
//                   nb[] holds the e-th listed slot of every cell (entry-major,
//                   so the lanes of a warp read and write consecutive words);
//   interact_lists  phase 2 alone: one thread per cell walks its list in order
//                   (the reference's order), gathers the partner from the
//                   cube-ordered planes through L1, exact cut-off, functor,
//                   friction, epilogue. No shared memory to speak of, so nearly
//                   the whole 256 KB of the SM is L1 for those gathers and for
//                   the functor's own, and the register budget is the functor's.
//
// A cell with more than LIST_MAX listed candidates (crowded tissues) raises
// Step_ctl::list_overflow; interact_lists then leaves the stage to the fused
// kernel, which is always launched behind it and returns at once otherwise.
constexpr int LIST_MAX = 64;

// Walk the cells of a chunk by descending list length, so that the lanes of a
// warp get lists of about equal length. Measured (profiles/r02_sweep_tuning.md):
// no gain -- the sweep waits on the functor's memory chain, not on idle lanes,
// and the permuted cell order costs locality -- so it is off.
#ifndef YB_INTERACT_BALANCE
#define YB_INTERACT_BALANCE 0
#endif

struct List_config {  // list_cubes: the float3 budget of the fused sweep
    static constexpr int stage_cap = 1024;
    static constexpr int list_cap = 24;
    static constexpr int min_ctas = 8;
    static constexpr size_t smem =
        size_t(stage_cap) * sizeof(float4) +
        size_t(SWEEP_ROWS) * SWEEP_THREADS * sizeof(uint32_t) +
        size_t(list_cap) * SWEEP_THREADS * sizeof(uint16_t);
};

__global__ void __launch_bounds__(SWEEP_THREADS, List_config::min_ctas)
    list_cubes(const int* __restrict__ d_n, int n_max,
        const float4* __restrict__ pos4, const int* __restrict__ cube_sorted,
        const int* __restrict__ offset, float cube_size, Grid_box box,
        int* __restrict__ nb, int* __restrict__ nb_count,
        unsigned char* __restrict__ nb_order, int nb_stride, Step_ctl* ctl,
        int overflow_at)
{
    constexpr int SWEEP_STAGE_CAP = List_config::stage_cap;
    constexpr int SWEEP_LIST_CAP = List_config::list_cap;
    extern __shared__ __align__(16) unsigned char sweep_smem[];
    float4* s_pos = reinterpret_cast<float4*>(sweep_smem);
    uint32_t* s_range = reinterpret_cast<uint32_t*>(s_pos + SWEEP_STAGE_CAP);
    uint16_t* s_list =
        reinterpret_cast<uint16_t*>(s_range + SWEEP_ROWS * SWEEP_THREADS);
    __shared__ int s_row_lo[SWEEP_ROWS];
    __shared__ int s_row_v[SWEEP_ROWS + 1];
    __shared__ __align__(8) uint64_t s_bar;
#if YB_INTERACT_BALANCE
    // balancing of interact_lists: cells per list length, per warp
    __shared__ unsigned char s_hist[SWEEP_THREADS / 32][LIST_MAX + 1];
    __shared__ unsigned char s_longer[LIST_MAX + 2];
#endif

    const int t = threadIdx.x;
    const int n = live_cells(d_n, n_max);
    const int n_chunks = ceil_div(n, SWEEP_THREADS);
    const float reach2 = cube_size * cube_size * SWEEP_PREFILTER_SLACK;
    const int n_owned = ctl->external_drift ? ctl->n_owned : n_max;

    if (t == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    uint32_t parity = 0;

    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int first_slot = chunk * SWEEP_THREADS;
        const int k = first_slot + t;
        const bool live = k < n;

        if (t < SWEEP_ROWS) {
            const int last_slot = min(first_slot + SWEEP_THREADS, n) - 1;
            const int shift = row_shift(t, box);
            const int lo = __ldg(offset +
                clamp_cube(__ldg(cube_sorted + first_slot) + shift - 1, box.n_cubes));
            const int hi = __ldg(offset +
                clamp_cube(__ldg(cube_sorted + last_slot) + shift + 2, box.n_cubes));
            s_row_lo[t] = lo;
            s_row_v[t + 1] = hi > lo ? hi - lo : 0;
        }
#if YB_INTERACT_BALANCE
        for (int q = t; q < (SWEEP_THREADS / 32) * (LIST_MAX + 1); q += SWEEP_THREADS)
            (&s_hist[0][0])[q] = 0;
#endif
        float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
        int my_cube = 0;
        if (live) {
            me = __ldg(pos4 + k);
            my_cube = __ldg(cube_sorted + k);
        }
        float gap2[3][3];
        trim_gaps(me, cube_size, box, my_cube, gap2);
        int my_lo[SWEEP_ROWS], my_hi[SWEEP_ROWS];
#pragma unroll
        for (int r = 0; r < SWEEP_ROWS; r++) {
            const int c = my_cube + row_shift(r, box);
            const float row_gap2 = gap2[1][r % 3] + gap2[2][r / 3];
            const bool row_out = !live || row_gap2 >= SWEEP_TRIM_LIMIT;
            const int first = row_gap2 + gap2[0][1] >= SWEEP_TRIM_LIMIT ? c : c - 1;
            const int last = row_gap2 + gap2[0][2] >= SWEEP_TRIM_LIMIT ? c + 1 : c + 2;
            my_lo[r] = row_out ? 0 : __ldg(offset + clamp_cube(first, box.n_cubes));
            my_hi[r] = row_out ? 0 : __ldg(offset + clamp_cube(last, box.n_cubes));
        }
        __syncthreads();
        if (t == 0) {
            int v = 0;
            s_row_v[0] = 0;
#pragma unroll
            for (int r = 0; r < SWEEP_ROWS; r++) {
                v += s_row_v[r + 1];
                s_row_v[r + 1] = v;
            }
        }
        __syncthreads();
        const int total = s_row_v[SWEEP_ROWS];
        const bool owned = live && __float_as_int(me.w) < n_owned;
        int n_listed = 0;

        for (int v0 = 0; v0 < total; v0 += SWEEP_STAGE_CAP) {
            if (v0 > 0) __syncthreads();
            if (t == 0) {
                const int v1 = min(v0 + SWEEP_STAGE_CAP, total);
                mbar_expect_tx(&s_bar, uint32_t(v1 - v0) * sizeof(float4));
#pragma unroll 1
                for (int r = 0; r < SWEEP_ROWS; r++) {
                    const int a = max(s_row_v[r], v0);
                    const int b = min(s_row_v[r + 1], v1);
                    if (b > a)
                        bulk_g2s(s_pos + (a - v0),
                            pos4 + (s_row_lo[r] + (a - s_row_v[r])),
                            uint32_t(b - a) * sizeof(float4), &s_bar);
                }
            }
#pragma unroll
            for (int r = 0; r < SWEEP_ROWS; r++) {
                const int base = s_row_v[r] - s_row_lo[r] - v0;
                const int a = min(max(my_lo[r] + base, 0), SWEEP_STAGE_CAP);
                const int b = min(max(my_hi[r] + base, a), SWEEP_STAGE_CAP);
                s_range[r * SWEEP_THREADS + t] = uint32_t(a) | (uint32_t(b) << 16);
            }
            mbar_wait(&s_bar, parity);
            parity ^= 1u;

            const uint32_t list_begin = smem_u32(s_list + t);
            const uint32_t list_full =
                list_begin + SWEEP_LIST_CAP * SWEEP_THREADS * sizeof(uint16_t);
            int r = 0;
            uint32_t range = owned ? s_range[t] : 0u;
            int a = int(range & 0xffffu), b = int(range >> 16);
            while (true) {
                uint32_t la = list_begin;
                while (true) {
                    const float4* pp = s_pos + a;
                    uint32_t entry = (uint32_t(r) << 12) | uint32_t(a);
                    const uint32_t entry_end = entry + uint32_t(b - a);
#pragma unroll 2
                    for (; entry < entry_end && la != list_full; entry++, pp++) {
                        const float4 p = *pp;
                        const float dx = me.x - p.x, dy = me.y - p.y,
                                    dz = me.z - p.z;
                        const float d2 = dx * dx + dy * dy + dz * dz;
                        if (!(d2 > reach2)) {
                            asm volatile("st.shared.u16 [%0], %1;" ::"r"(la),
                                         "h"(uint16_t(entry))
                                         : "memory");
                            la += SWEEP_THREADS * sizeof(uint16_t);
                        }
                    }
                    a = int(entry & 4095u);
                    if (__any_sync(0xffffffffu, a < b)) break;  // list(s) full
                    ++r;
                    range = owned && r < SWEEP_ROWS
                                ? s_range[r * SWEEP_THREADS + t]
                                : 0u;
                    a = int(range & 0xffffu), b = int(range >> 16);
                    if (r == SWEEP_ROWS) break;
                }
                // flush: the e-th entry of every lane goes to row n_listed + e
                const int listed = listed_at(la, list_begin);
                for (int e = 0; e < listed; e++) {
                    const unsigned entry = s_list[e * SWEEP_THREADS + t];
                    const int row = entry >> 12, at = entry & 4095;
                    const int kj = s_row_lo[row] + (v0 + at - s_row_v[row]);
                    if (n_listed + e < LIST_MAX)
                        nb[size_t(n_listed + e) * nb_stride + k] = kj;
                }
                n_listed += listed;
                if (r == SWEEP_ROWS) break;
            }
        }
        if (live) nb_count[k] = n_listed;
        if (n_listed > overflow_at) ctl->list_overflow = 1;  // <= LIST_MAX

#if YB_INTERACT_BALANCE
        // Which cell of the chunk thread q of interact_lists takes: the cells
        // in order of DESCENDING list length (ties: ascending slot), so that
        // the lanes of a warp walk lists of about the same length -- a warp
        // iterates the maximum over its lanes. A stable counting sort over the
        // (at most LIST_MAX + 1) distinct lengths.
        const int key = min(n_listed, LIST_MAX);
        const int lane_id = t & 31, warp_id = t >> 5;
        const unsigned same = __match_any_sync(0xffffffffu, key);
        if (lane_id == __ffs(same) - 1) s_hist[warp_id][key] = __popc(same);
        __syncthreads();
        if (t <= LIST_MAX) {  // cells per length, over the warps
            int total = 0;
#pragma unroll
            for (int w = 0; w < SWEEP_THREADS / 32; w++) total += s_hist[w][t];
            s_longer[t] = total;
        }
        __syncthreads();
        if (t == 0) {  // cells with a longer list than l, for every l
            int running = 0;
            for (int l = LIST_MAX; l >= 0; l--) {
                const int here = s_longer[l];
                s_longer[l] = running;
                running += here;
            }
        }
        __syncthreads();
        int position = s_longer[key] + __popc(same & ((1u << lane_id) - 1u));
        for (int w = 0; w < warp_id; w++) position += s_hist[w][key];
        if (first_slot + position < nb_stride)
            nb_order[first_slot + position] = static_cast<unsigned char>(t);
#endif
        __syncthreads();  // the next chunk reuses every shared array
    }
}

#ifndef YB_INTERACT_CTAS
#define YB_INTERACT_CTAS 6
#endif
#ifndef YB_INTERACT_PREFETCH  // gather partner e + 1 while pair e is evaluated
#define YB_INTERACT_PREFETCH 1
#endif

template<typename Pt>
struct Partner {  // one listed neighbour, as gathered from the cube-ordered planes
    float4 pos;
    float4 aux[Layout<Pt>::aux_vec4];
};

template<typename Pt>
__device__ __forceinline__ Partner<Pt> gather_partner(
    const float4* __restrict__ pos4, const float4* __restrict__ aux, int kj)
{
    using L = Layout<Pt>;
    Partner<Pt> p;
    p.pos = __ldg(pos4 + kj);
#pragma unroll
    for (int q = 0; q < L::aux_vec4; q++)
        p.aux[q] = __ldg(aux + size_t(kj) * L::aux_vec4 + q);
    return p;
}

template<typename Pt>
__device__ __forceinline__ Pt partner_pt(const Partner<Pt>& p)
{
    using L = Layout<Pt>;
    Pt X;
    lane(X, 0) = p.pos.x, lane(X, 1) = p.pos.y, lane(X, 2) = p.pos.z;
    const float* a = reinterpret_cast<const float*>(p.aux);
#pragma unroll
    for (int e = 0; e < L::extras; e++) lane(X, 3 + e) = a[e];
    return X;
}

template<typename Pt>
__device__ __forceinline__ float3 partner_velocity(const Partner<Pt>& p)
{
    const float* a = reinterpret_cast<const float*>(p.aux);
    constexpr int v = Layout<Pt>::v_lane;
    return float3{a[v], a[v + 1], a[v + 2]};
}

template<typename Pt, Pt (*pw_int)(Pt, Pt, float, int, int),
    float (*pw_friction)(Pt, Pt, float, int, int), bool SEEDED>
__global__ void __launch_bounds__(SWEEP_THREADS, YB_INTERACT_CTAS) interact_lists(
    const int* __restrict__ d_n, int n_max, const float4* __restrict__ pos4,
    const float4* __restrict__ aux, const int* __restrict__ nb,
    const int* __restrict__ nb_count, const unsigned char* __restrict__ nb_order,
    int nb_stride, float cube_size, Pt* d_dX, float* __restrict__ partials,
    int stage, int drift_mode, int fix_point, Step_ctl* ctl)
{
    using L = Layout<Pt>;
    __shared__ float s_red[3][SWEEP_THREADS / 32];
    // crowded: the fused kernel behind this one takes the stage
    if (*(volatile int*)&ctl->list_overflow) return;

    const int t = threadIdx.x;
    const int n = live_cells(d_n, n_max);
    const int n_chunks = ceil_div(n, SWEEP_THREADS);
    const int n_owned = ctl->external_drift ? ctl->n_owned : n_max;
    float3 my_sum{0.f, 0.f, 0.f};  // this thread's cells, chunk after chunk

    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int first_slot = chunk * SWEEP_THREADS;
        // thread t takes the cell with the t-th longest list of the chunk
        const int k = first_slot +
                      (YB_INTERACT_BALANCE ? __ldg(nb_order + first_slot + t) : t);
        if (first_slot + t >= n || k >= n) continue;
        const float4 me = __ldg(pos4 + k);
        const int my_id = __float_as_int(me.w);
        if (my_id >= n_owned) continue;  // ghost: a partner, never a subject
        const Pt Xi = assemble_pt<Pt>(me, aux + size_t(k) * L::aux_vec4);
        const int count = min(__ldg(nb_count + k), LIST_MAX);
        Pt F{0};
        float3 sum_v{0.f, 0.f, 0.f};
        float sum_friction = 0.f;

        // Two gathers deep: while pair e is with the functor, the planes of
        // partner e + 1 and the list entry e + 2 are already on their way.
        const int* my_nb = nb + k;
        int kj_next = count > 1 ? __ldg(my_nb + nb_stride) : 0;
        Partner<Pt> next = gather_partner<Pt>(pos4, aux, count > 0 ? __ldg(my_nb) : k);
        for (int e = 0; e < count; e++) {
#if YB_INTERACT_PREFETCH
            const Partner<Pt> now = next;
            const int kj_after =
                e + 2 < count ? __ldg(my_nb + size_t(e + 2) * nb_stride) : 0;
            if (e + 1 < count) next = gather_partner<Pt>(pos4, aux, kj_next);
            kj_next = kj_after;
#else
            const Partner<Pt> now = gather_partner<Pt>(
                pos4, aux, __ldg(my_nb + size_t(e) * nb_stride));
#endif

            const Pt rij = Xi - partner_pt<Pt>(now);
            const int j_id = __float_as_int(now.pos.w);
            const float dist = pair_distance(rij.x, rij.y, rij.z, j_id == my_id);
            if (dist >= cube_size) continue;

            F += pw_int(Xi, rij, dist, my_id, j_id);
            const float friction = pw_friction(Xi, rij, dist, my_id, j_id);
            sum_friction += friction;
            if (friction != 0.f) sum_v += friction * partner_velocity<Pt>(now);
        }

        Pt dX = F;
        if (SEEDED) {
            dX = load_pt_rw(d_dX, my_id);
            dX += F;
        }
        if (sum_friction > 0) {
            dX.x += sum_v.x / sum_friction;
            dX.y += sum_v.y / sum_friction;
            dX.z += sum_v.z / sum_friction;
        }
        store_pt(d_dX, my_id, dX);
        my_sum.x += dX.x, my_sum.y += dX.y, my_sum.z += dX.z;
    }

    // one deterministic CTA-wide sum at the end (chunks are dealt in a fixed
    // order and the cell order inside a chunk is a stable sort, so every
    // thread's running sum is reproducible)
    __syncthreads();
    const float3 cta_sum =
        block_sum3<SWEEP_THREADS>(my_sum.x, my_sum.y, my_sum.z, s_red);
    finish_drift<SWEEP_THREADS>(cta_sum, partials, n, stage, drift_mode,
        fix_point, d_dX, ctl, s_red);
}


// ---- Tile solver sweep -------------------------------------------------------
// All pairs, self pair included, j ascending -- same order as compute_tile.
// Tiles of the cube-order-free AoS state are staged to shared memory by the
// whole CTA (plain loads: n is small wherever the Tile solver is used, and Pt
// records are only 4-byte aligned, which rules out bulk copies).
#ifndef YB_TILE_THREADS
#define YB_TILE_THREADS 128
#endif
#ifndef YB_TILE_UNROLL
#define YB_TILE_UNROLL 4
#endif
constexpr int TILE_THREADS = YB_TILE_THREADS;
constexpr int TILE_UNROLL = YB_TILE_UNROLL;

template<typename Pt, Pt (*pw_int)(Pt, Pt, float, int, int),
    float (*pw_friction)(Pt, Pt, float, int, int), bool SEEDED>
__global__ void __launch_bounds__(TILE_THREADS) sweep_tiles(
    const int* __restrict__ d_n, int n_max, const Pt* __restrict__ d_X,
    const float3* __restrict__ d_old_v, Pt* d_dX, float* __restrict__ partials,
    int stage, int drift_mode, int fix_point, Step_ctl* ctl)
{
    using L = Layout<Pt>;
    __shared__ float s_X[TILE_THREADS * L::lanes];
    __shared__ float s_v[TILE_THREADS * 3];
    __shared__ float s_red[3][TILE_THREADS / 32];

    const int t = threadIdx.x;
    const int n = live_cells(d_n, n_max);
    const int n_chunks = ceil_div(n, TILE_THREADS);
    float3 cta_sum{0.f, 0.f, 0.f};

    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int i = chunk * TILE_THREADS + t;
        const bool live = i < n;
        Pt Xi{0};
        if (live) Xi = load_pt(d_X, i);
        Pt F{0};
        float3 sum_v{0.f, 0.f, 0.f};
        float sum_friction = 0.f;

        for (int tile_start = 0; tile_start < n; tile_start += TILE_THREADS) {
            const int in_tile = min(TILE_THREADS, n - tile_start);
            __syncthreads();
            // coalesced copy of in_tile AoS records, lane by lane
            const float* src_X = reinterpret_cast<const float*>(d_X + tile_start);
            for (int q = t; q < in_tile * L::lanes; q += TILE_THREADS)
                s_X[q] = __ldg(src_X + q);
            const float* src_v =
                reinterpret_cast<const float*>(d_old_v + tile_start);
            for (int q = t; q < in_tile * 3; q += TILE_THREADS)
                s_v[q] = __ldg(src_v + q);
            __syncthreads();

            if (live) {
#pragma unroll TILE_UNROLL
                for (int q = 0; q < in_tile; q++) {
                    Pt Xj;
#pragma unroll
                    for (int l = 0; l < L::lanes; l++)
                        lane(Xj, l) = s_X[q * L::lanes + l];
                    const int j = tile_start + q;
                    const Pt rij = Xi - Xj;
                    const float dist = norm3df(rij.x, rij.y, rij.z);
                    F += pw_int(Xi, rij, dist, i, j);
                    const float friction = pw_friction(Xi, rij, dist, i, j);
                    sum_friction += friction;
                    if (friction != 0.f)
                        sum_v += friction *
                                 float3{s_v[3 * q], s_v[3 * q + 1], s_v[3 * q + 2]};
                }
            }
        }

        float3 mine{0.f, 0.f, 0.f};
        if (live) {
            Pt dX = F;
            if (SEEDED) {
                dX = load_pt_rw(d_dX, i);
                dX += F;
            }
            if (sum_friction > 0) {
                dX.x += sum_v.x / sum_friction;
                dX.y += sum_v.y / sum_friction;
                dX.z += sum_v.z / sum_friction;
            }
            store_pt(d_dX, i, dX);
            mine = float3{dX.x, dX.y, dX.z};
        }
        __syncthreads();
        const float3 chunk_sum =
            block_sum3<TILE_THREADS>(mine.x, mine.y, mine.z, s_red);
        cta_sum.x += chunk_sum.x, cta_sum.y += chunk_sum.y, cta_sum.z += chunk_sum.z;
    }

    finish_drift<TILE_THREADS>(cta_sum, partials, n, stage, drift_mode,
        fix_point, d_dX, ctl, s_red);
}


// ---- Tile solver sweep, pairs split across lanes (opt-in) ----------------------
// With a few hundred cells, one thread per cell leaves the machine idle and
// every thread with a chain of n dependent pair evaluations. Here a group of G
// lanes shares one cell i: lane g takes the partners j = g, g + G, ... and the
// partial sums are combined by a fixed shuffle tree (deterministic, but a
// different summation order than the reference). The functor runs in G
// threads per cell, so this is only valid for functors WITHOUT per-cell side
// effects; it is enabled per solver with `cells.split_pairs = true`.
template<typename Pt, Pt (*pw_int)(Pt, Pt, float, int, int),
    float (*pw_friction)(Pt, Pt, float, int, int), bool SEEDED, int G>
__global__ void __launch_bounds__(TILE_THREADS) sweep_tiles_split(
    const int* __restrict__ d_n, int n_max, const Pt* __restrict__ d_X,
    const float3* __restrict__ d_old_v, Pt* d_dX, float* __restrict__ partials,
    int stage, int drift_mode, int fix_point, Step_ctl* ctl)
{
    using L = Layout<Pt>;
    constexpr int CELLS = TILE_THREADS / G;  // cells per CTA and chunk
    __shared__ float s_X[TILE_THREADS * L::lanes];
    __shared__ float s_v[TILE_THREADS * 3];
    __shared__ float s_red[3][TILE_THREADS / 32];

    const int t = threadIdx.x, g = t % G;
    const int n = live_cells(d_n, n_max);
    const int n_chunks = ceil_div(n, CELLS);
    float3 cta_sum{0.f, 0.f, 0.f};

    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int i = chunk * CELLS + t / G;
        const bool live = i < n;
        Pt Xi{0};
        if (live) Xi = load_pt(d_X, i);
        Pt F{0};
        float3 sum_v{0.f, 0.f, 0.f};
        float sum_friction = 0.f;

        for (int tile_start = 0; tile_start < n; tile_start += TILE_THREADS) {
            const int in_tile = min(TILE_THREADS, n - tile_start);
            __syncthreads();
            const float* src_X = reinterpret_cast<const float*>(d_X + tile_start);
            for (int q = t; q < in_tile * L::lanes; q += TILE_THREADS)
                s_X[q] = __ldg(src_X + q);
            const float* src_v =
                reinterpret_cast<const float*>(d_old_v + tile_start);
            for (int q = t; q < in_tile * 3; q += TILE_THREADS)
                s_v[q] = __ldg(src_v + q);
            __syncthreads();

            if (live) {
                for (int q = g; q < in_tile; q += G) {
                    Pt Xj;
#pragma unroll
                    for (int l = 0; l < L::lanes; l++)
                        lane(Xj, l) = s_X[q * L::lanes + l];
                    const int j = tile_start + q;
                    const Pt rij = Xi - Xj;
                    const float dist = norm3df(rij.x, rij.y, rij.z);
                    F += pw_int(Xi, rij, dist, i, j);
                    const float friction = pw_friction(Xi, rij, dist, i, j);
                    sum_friction += friction;
                    if (friction != 0.f)
                        sum_v += friction *
                                 float3{s_v[3 * q], s_v[3 * q + 1], s_v[3 * q + 2]};
                }
            }
        }

        // combine the G partial sums of every cell (all lanes take part)
#pragma unroll
        for (int d = G / 2; d > 0; d >>= 1) {
#pragma unroll
            for (int l = 0; l < L::lanes; l++)
                lane(F, l) += __shfl_xor_sync(0xffffffffu, lane(F, l), d);
            sum_v.x += __shfl_xor_sync(0xffffffffu, sum_v.x, d);
            sum_v.y += __shfl_xor_sync(0xffffffffu, sum_v.y, d);
            sum_v.z += __shfl_xor_sync(0xffffffffu, sum_v.z, d);
            sum_friction += __shfl_xor_sync(0xffffffffu, sum_friction, d);
        }

        float3 mine{0.f, 0.f, 0.f};
        if (live && g == 0) {
            Pt dX = F;
            if (SEEDED) {
                dX = load_pt_rw(d_dX, i);
                dX += F;
            }
            if (sum_friction > 0) {
                dX.x += sum_v.x / sum_friction;
                dX.y += sum_v.y / sum_friction;
                dX.z += sum_v.z / sum_friction;
            }
            store_pt(d_dX, i, dX);
            mine = float3{dX.x, dX.y, dX.z};
        }
        __syncthreads();
        const float3 chunk_sum =
            block_sum3<TILE_THREADS>(mine.x, mine.y, mine.z, s_red);
        cta_sum.x += chunk_sum.x, cta_sum.y += chunk_sum.y, cta_sum.z += chunk_sum.z;
    }

    finish_drift<TILE_THREADS>(cta_sum, partials, n, stage, drift_mode,
        fix_point, d_dX, ctl, s_red);
}

}  // namespace yb
