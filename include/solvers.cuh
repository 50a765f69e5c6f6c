// Solvers for N-body problems -- B200-native implementation of the ya||a
// solver API (reference: /root/reference/include/solvers.cuh).
//
// The public surface is the reference's: Solution<Pt, Solver>, take_step<pw_int
// [, pw_friction]>(dt, generic_forces), copy_to_device/host, get_d_n,
// set_fixed*, Tile_solver / Grid_solver / Gabriel_solver, Grid and d_nhood.
// A model written for ya||a compiles against this header unchanged. Underneath,
// a step is a fixed sequence of hand-written sm_100a kernels with no Thrust
// call, no host synchronisation, no allocation and no device->host copy:
//
//   Grid solver, per Heun stage          (kernels: b200/grid_build.cuh,
//     bin_cells*    cube ids + bucket counts        b200/pair_sweep.cuh,
//     scan_bins     per-cube offsets                b200/heun.cuh)
//     place_ids     bucket scatter
//     reorder_cells state -> cube order (SoA float4 planes)
//     sweep_cubes   27-cube pairwise sum + friction term + drift reduction
//     predictor_step / corrector_step   (* stage 2's binning is fused into
//                                          predictor_step)
//
// The cell count is read from d_n on the device, so a step needs nothing from
// the host; when no generic force is passed, the whole step is captured once
// as a CUDA graph and replayed.
#pragma once

#include <assert.h>
#include <stdlib.h>
#include <cmath>
// The solver itself does not use Thrust. These four headers are included only
// because ya||a models call thrust::fill / thrust::reduce in their own code and
// rely on solvers.cuh to have pulled them in (reference: solvers.cuh:5-8).
#include <thrust/execution_policy.h>
#include <thrust/fill.h>
#include <thrust/reduce.h>
#include <thrust/sort.h>
#include <functional>
#include <type_traits>
#include <utility>
#include <vector>

#include "cudebug.cuh"
#include "dtypes.cuh"

#include "b200/grid_build.cuh"
#include "b200/heun.cuh"
#include "b200/layout.cuh"
#include "b200/pair_sweep.cuh"
#include "b200/slab.cuh"
#include "b200/domain.cuh"

#define YALLA_B200 1


// ---- user-supplied pieces -------------------------------------------------------
// A pairwise interaction returns the contribution of j to dXi/dt, given
// r = Xi - Xj (all members of Pt) and dist = |r.xyz|. i and j are cell ids in
// the user's arrays (reference: solvers.cuh:15-19).
template<typename Pt>
using Pairwise_interaction = Pt(Pt Xi, Pt r, float dist, int i, int j);

// A pairwise friction coefficient weights the neighbour's previous velocity in
// v_i = F_i + <v_j(t - dt)>, see http://dx.doi.org/10.1007/s10237-014-0613-5.
template<typename Pt>
using Pairwise_friction = float(Pt Xi, Pt r, float dist, int i, int j);

// Default: neighbours closer than 1 drag on each other.
template<typename Pt>
__device__ float friction_w_neighbour(Pt Xi, Pt r, float dist, int i, int j)
{
    return (i != j && dist < 1) ? 1.f : 0.f;
}

// No neighbour friction: cells only feel the (implicit) background.
template<typename Pt>
__device__ float friction_on_background(Pt Xi, Pt r, float dist, int i, int j)
{
    return 0;
}

// Generic forces are host callables invoked once per Heun stage with the stage's
// positions and the (zeroed) derivative array, BEFORE the pairwise sweep adds to
// it -- e.g. link_forces, or a reset of neighbour counters.
// This is std::function<void(int n, const Pt* d_X, Pt* d_dX)> like the
// reference's alias (solvers.cuh:44-46), plus a converting constructor for the
// older two-argument form (d_X, d_dX) that the upstream tests still use.
template<typename Pt>
class Generic_forces
    : public std::function<void(const int, const Pt*, Pt*)> {
    using Base = std::function<void(const int, const Pt*, Pt*)>;

public:
    using Base::Base;
    Generic_forces() = default;

    template<typename F,
        typename = decltype(std::declval<F&>()(
            static_cast<const Pt*>(nullptr), static_cast<Pt*>(nullptr)))>
    Generic_forces(F two_argument_form)
        : Base([two_argument_form](const int, const Pt* d_X,
                   Pt* d_dX) mutable { two_argument_form(d_X, d_dX); })
    {}
};

template<typename Pt>
void no_gen_forces(const int n, const Pt* __restrict__ d_X, Pt* d_dX)
{}


namespace yb {

// True if the callable is exactly no_gen_forces<Pt>: then nothing seeds dX and
// the step can run from a captured graph.
template<typename Pt>
bool is_no_gen_forces(const Generic_forces<Pt>& f)
{
    using Fn = void (*)(const int, const Pt*, Pt*);
    const Fn* target = f.template target<Fn>();
    return target != nullptr && *target == &no_gen_forces<Pt>;
}

inline bool graphs_enabled()
{
    static bool enabled = [] {
        const char* v = getenv("YALLA_B200_NO_GRAPH");
        return !(v && v[0] && v[0] != '0');
    }();
    return enabled;
}

// What a solver bakes into the kernel arguments of a captured step besides the
// time step and the drift mode (compared field by field).
struct Graph_key {
    float a = 0.f, b = 0.f;
    int c = 0, d = 0;
    bool operator==(const Graph_key& o) const
    {
        return a == o.a && b == o.b && c == o.c && d == o.d;
    }
};

// A captured Heun step. Kernel arguments are baked into the graph, so it is
// keyed by everything the host passes by value; `hooks` are the checks that
// capturable generic forces asked to run before every replay.
struct Step_graph {
    const void* instantiation;  // identifies take_step<pw_int, pw_friction>
    const void* forces_type;    // typeid of the generic forces, or nullptr
    float dt;
    Graph_key solver_key;
    int mode0, mode1, fix_point;
    cudaGraphExec_t exec;
    std::vector<Replay_hook> hooks;
};

// Zero the derivative of the live cells before the generic forces add to it
// (the count is read on the device: usable inside a captured step).
template<typename Pt>
__global__ void __launch_bounds__(256) zero_cells(
    const int* __restrict__ d_n, int n_max, Pt* d_dX)
{
    const int n_floats = live_cells(d_n, n_max) * Layout<Pt>::lanes;
    float* out = reinterpret_cast<float*>(d_dX);
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n_floats;
         q += gridDim.x * blockDim.x)
        out[q] = 0.f;
}

}  // namespace yb


// Solution<Pt, Solver> combines a method, Solver, with a point type, Pt. It
// owns the host copy of the state and exposes the device arrays the solver
// integrates (reference: solvers.cuh:56-106).
template<typename Pt, template<typename> class Solver>
class Solution : public Solver<Pt> {
public:
    Pt* h_X;                                      // State on the host
    Pt* const d_X = Solver<Pt>::d_X;              // State on the device
    float3* const d_old_v = Solver<Pt>::d_old_v;  // Velocities of the last step
    int* const h_n = (int*)malloc(sizeof(int));   // Number of cells
    int* const d_n = Solver<Pt>::d_n;
    const int n_max;

    template<typename... Args>
    Solution(int n_max, Args... args)
        : Solver<Pt>{n_max, args...}, n_max{n_max}
    {
        *h_n = n_max;
        // Pinned, so that copies run at full PCIe rate; plain malloc if the
        // host cannot pin that much.
        const size_t bytes = static_cast<size_t>(n_max) * sizeof(Pt);
        if (cudaMallocHost(&h_X, bytes) == cudaSuccess) {
            h_X_pinned = true;
        } else {
            cudaGetLastError();
            h_X = static_cast<Pt*>(malloc(bytes));
        }
    }
    Solution(const Solution&) = delete;
    Solution& operator=(const Solution&) = delete;
    ~Solution()
    {
        if (h_X_pinned)
            cudaFreeHost(h_X);
        else
            free(h_X);
        free(h_n);
    }

    // Both copies move all n_max elements plus n and block, as in the
    // reference. They are issued to the solver's stream and waited for, so
    // they are ordered with the steps in flight whatever stream that is.
    void copy_to_device()
    {
        assert(*h_n <= n_max);
        const cudaStream_t s = Solver<Pt>::stream;
        YB_CUDA(cudaMemcpyAsync(d_X, h_X, static_cast<size_t>(n_max) * sizeof(Pt),
            cudaMemcpyHostToDevice, s));
        YB_CUDA(cudaMemcpyAsync(
            d_n, h_n, sizeof(int), cudaMemcpyHostToDevice, s));
        YB_CUDA(cudaStreamSynchronize(s));
    }
    void copy_to_host()
    {
        const cudaStream_t s = Solver<Pt>::stream;
        YB_CUDA(cudaMemcpyAsync(h_X, d_X, static_cast<size_t>(n_max) * sizeof(Pt),
            cudaMemcpyDeviceToHost, s));
        YB_CUDA(cudaMemcpyAsync(
            h_n, d_n, sizeof(int), cudaMemcpyDeviceToHost, s));
        YB_CUDA(cudaStreamSynchronize(s));
        assert(*h_n <= n_max);
    }
    int get_d_n() { return Solver<Pt>::get_d_n(); }

    // Extension: move only the n live cells, straight between the caller's
    // buffer and the device, on the solver's stream (full PCIe rate if the
    // buffer is pinned). upload() also sets the cell count; download() waits.
    void upload(const Pt* h_src, int n)
    {
        assert(n <= n_max);
        *h_n = n;
        YB_CUDA(cudaMemcpyAsync(d_X, h_src, static_cast<size_t>(n) * sizeof(Pt),
            cudaMemcpyHostToDevice, Solver<Pt>::stream));
        YB_CUDA(cudaMemcpyAsync(
            d_n, h_n, sizeof(int), cudaMemcpyHostToDevice, Solver<Pt>::stream));
    }
    int download(Pt* h_dst, int capacity)
    {
        const int n = get_d_n();
        assert(n <= capacity);
        YB_CUDA(cudaMemcpyAsync(h_dst, d_X, static_cast<size_t>(n) * sizeof(Pt),
            cudaMemcpyDeviceToHost, Solver<Pt>::stream));
        YB_CUDA(cudaStreamSynchronize(Solver<Pt>::stream));
        *h_n = n;
        return n;
    }

    template<Pairwise_interaction<Pt> pw_int>
    void take_step(float dt, Generic_forces<Pt> gen_forces = no_gen_forces<Pt>)
    {
        Solver<Pt>::template take_step<pw_int, friction_w_neighbour<Pt>>(
            dt, gen_forces);
    }
    template<Pairwise_interaction<Pt> pw_int, Pairwise_friction<Pt> pw_friction>
    void take_step(float dt, Generic_forces<Pt> gen_forces = no_gen_forces<Pt>)
    {
        Solver<Pt>::template take_step<pw_int, pw_friction>(dt, gen_forces);
    }

private:
    bool h_X_pinned = false;
};


// 2nd order solver for the equation v = F + <v(t - dt)> for x, y, and z, where
// <v> is the mean velocity of the neighbours weighted by the friction
// coefficients. One point or the centre of mass is kept fixed. Solves
// dw/dt = F_w for the other members of Pt (reference: solvers.cuh:109-276).
// Computer says how the pairwise sums are formed.
template<typename Pt, template<typename> class Computer>
class Heun_solver : public Computer<Pt> {
public:
    template<typename... Args>
    Heun_solver(int n_max, Args... args)
        : Computer<Pt>{n_max, args...}, n_max{n_max}
    {
        const size_t cells = static_cast<size_t>(n_max > 0 ? n_max : 1);
        YB_CUDA(cudaMalloc(&d_X, cells * sizeof(Pt)));
        YB_CUDA(cudaMalloc(&d_dX, cells * sizeof(Pt)));
        YB_CUDA(cudaMalloc(&d_X1, cells * sizeof(Pt)));
        YB_CUDA(cudaMalloc(&d_dX1, cells * sizeof(Pt)));
        YB_CUDA(cudaMalloc(&d_old_v, cells * sizeof(float3)));
        YB_CUDA(cudaMemset(d_old_v, 0, cells * sizeof(float3)));
        YB_CUDA(cudaMalloc(&d_n, sizeof(int)));
        YB_CUDA(cudaMemset(d_n, 0, sizeof(int)));

        YB_CUDA(cudaMalloc(&d_ctl, sizeof(yb::Step_ctl)));
        yb::Step_ctl fresh{};
        fresh.scan_epoch = 1;
        YB_CUDA(cudaMemcpy(
            d_ctl, &fresh, sizeof(fresh), cudaMemcpyHostToDevice));
        max_sweep_ctas = yb::sm_count() * 32;
        YB_CUDA(cudaMalloc(&d_partials, 3 * max_sweep_ctas * sizeof(float)));
        YB_CUDA(cudaStreamCreateWithFlags(
            &capture_stream, cudaStreamNonBlocking));
        YB_CUDA(cudaMallocHost(&h_n_pinned, sizeof(int)));
        YB_CUDA(cudaEventCreateWithFlags(&count_ready, cudaEventDisableTiming));
    }
    Heun_solver(const Heun_solver&) = delete;
    Heun_solver& operator=(const Heun_solver&) = delete;
    ~Heun_solver()
    {
        cudaStreamSynchronize(stream);
        for (auto& g : graphs) cudaGraphExecDestroy(g.exec);
        if (slab.capacity > 0) slab.release();
        dom.release();
        cudaFreeHost(h_n_pinned);
        cudaEventDestroy(count_ready);
        cudaStreamDestroy(capture_stream);
        cudaFree(d_partials);
        cudaFree(d_ctl);
        cudaFree(d_n);
        cudaFree(d_old_v);
        cudaFree(d_dX1);
        cudaFree(d_X1);
        cudaFree(d_dX);
        cudaFree(d_X);
    }

    // Keep the centre of mass fixed (default) ...
    void set_fixed() { fix_com = true; }
    // ... or one cell ...
    void set_fixed(int point_id)
    {
        fix_com = false;
        fix_point = point_id;
    }
    // ... or one cell in x and y, and the centre of mass in z.
    void set_fixed_xy(int point_id)
    {
        fix_com = false;
        fix_com_z = true;
        fix_point = point_id;
    }

    // Extension: number of cells whose cube id fell outside the grid so far
    // (the reference asserts on the device instead). Blocks.
    int cells_out_of_grid()
    {
        yb::Step_ctl snapshot;
        YB_CUDA(cudaMemcpy(
            &snapshot, d_ctl, sizeof(snapshot), cudaMemcpyDeviceToHost));
        return snapshot.out_of_grid;
    }

    // Extension: all solver work is issued to this stream (default: the legacy
    // default stream, like every launch in a ya||a model).
    cudaStream_t stream = 0;

    // ---- Extension: building blocks for domain decomposition ------------------
    // One solver integrates the cells of one spatial domain. Cells [0, n_owned)
    // are its own; cells [n_owned, n_total) are GHOSTS, copies of neighbouring
    // domains' boundary cells that act as interaction partners only. The drift
    // (mean force) is global, so the sweep leaves its local sum in device
    // memory, the caller reduces it over all domains (NCCL) and hands the mean
    // back. yalla_b200/dd.py drives these per stage; see DESIGN.md section 6.
    void dd_set_counts(int n_owned, int n_total)
    {
        assert(n_owned <= n_total && n_total <= n_max);
        const int header[2] = {n_owned, 1};  // n_owned, external_drift
        YB_CUDA(cudaMemcpyAsync(&d_ctl->n_owned, header, sizeof(header),
            cudaMemcpyHostToDevice, stream));
        YB_CUDA(cudaMemcpyAsync(
            d_n, &n_total, sizeof(int), cudaMemcpyHostToDevice, stream));
    }
    // Positions the given stage works on (stage 0: X, stage 1: X1), old
    // velocities, and the per-stage sums {sum dX.x, .y, .z, n_owned}.
    Pt* dd_positions(int stage) { return stage == 0 ? d_X : d_X1; }
    float3* dd_velocities() { return d_old_v; }
    const float* dd_drift_sum(int stage) { return d_ctl->drift_sum[stage]; }
    // Grid build + pairwise sweep of one stage over owned + ghost cells.
    template<Pairwise_interaction<Pt> pw_int, Pairwise_friction<Pt> pw_friction>
    void dd_forces(int stage)
    {
        cudaEvent_t sweep_start = nullptr, sweep_stop = nullptr;
        if (profiling) {
            YB_CUDA(cudaEventCreate(&sweep_start));
            YB_CUDA(cudaEventCreate(&sweep_stop));
        }
        Computer<Pt>::template pwints<pw_int, pw_friction, false>(stream, d_n,
            stage == 0 ? d_X : d_X1, d_old_v, stage == 0 ? d_dX : d_dX1,
            d_partials, max_sweep_ctas, stage, yb::DRIFT_MEAN, 0, d_ctl, false,
            sweep_start);
        if (profiling) {
            YB_CUDA(cudaEventRecord(sweep_stop, stream));
            sweep_events.emplace_back(sweep_start, sweep_stop);
        }
        YB_CUDA(cudaGetLastError());
    }
    // Predictor (stage 0) or corrector (stage 1) with the global drift d_mean
    // (3 floats in device memory).
    void dd_update(int stage, float dt, const float* d_mean)
    {
        YB_CUDA(cudaMemcpyAsync(d_ctl->drift[stage], d_mean, 3 * sizeof(float),
            cudaMemcpyDeviceToDevice, stream));
        const int blocks = yb::stride_grid(n_max, 256, yb::sm_count());
        if (stage == 0)
            yb::predictor_step<Pt, false><<<blocks, 256, 0, stream>>>(d_n, n_max,
                dt, d_X, d_dX, d_X1, d_ctl, 1.f, yb::Grid_box{}, nullptr, nullptr, nullptr);
        else
            yb::corrector_step<Pt><<<blocks, 256, 0, stream>>>(
                d_n, n_max, dt, d_dX, d_dX1, d_X, d_old_v, d_ctl);
        YB_CUDA(cudaGetLastError());
    }

    // ---- Extension: slab decomposition without host round trips ---------------
    // The same building blocks, but packing/unpacking of halo cells and of
    // migrating cells happens in kernels (b200/slab.cuh) and all counts stay on
    // the device: between slab_begin and the end of the run the host never has
    // to wait for the GPU. Exchange buffers hold a 4-float header (header[0] =
    // bits of the record count) followed by `capacity` records of
    // sizeof(Pt) / 4 + 3 floats; the caller ships them to the neighbouring
    // ranks at full size (NCCL send/recv, yalla_b200/dd.py).
    void slab_begin(float z_lo, float z_hi, float halo, int capacity)
    {
        if (slab.capacity == 0) slab.allocate(n_max);
        slab.capacity = capacity;
        slab.z_lo = z_lo, slab.z_hi = z_hi, slab.halo = halo;
        slab.has_lower = std::isfinite(z_lo), slab.has_upper = std::isfinite(z_hi);
        dd_set_counts(0, 0);
    }
    void slab_set_owned(const Pt* d_X_new, const float3* d_v_new, int n_owned)
    {
        assert(n_owned <= n_max);
        dom.flags_valid = false;
        YB_CUDA(cudaMemcpyAsync(d_X, d_X_new, sizeof(Pt) * size_t(n_owned),
            cudaMemcpyDeviceToDevice, stream));
        YB_CUDA(cudaMemcpyAsync(d_old_v, d_v_new, sizeof(float3) * size_t(n_owned),
            cudaMemcpyDeviceToDevice, stream));
        dd_set_counts(n_owned, n_owned);
    }
    // what = 0 / 1: the halo of the stage's positions (X / X1); 2: the cells that
    // left the slab during the step.
    void slab_pack(int what, float* send_lo, float* send_hi)
    {
        const bool migration = what == 2;
        const Pt* P = what == 1 ? d_X1 : d_X;
        // halo: cells within `halo` of a cut; migration: cells beyond it
        const float lo_edge = migration ? slab.z_lo : slab.z_lo + slab.halo;
        const float hi_edge = migration ? slab.z_hi : slab.z_hi - slab.halo;
        // X1 and dX are free at the end of a step: scratch for the stayers
        yb::slab_select<Pt><<<slab.n_tiles, yb::SCAN_THREADS, 0, stream>>>(d_ctl,
            slab.scan_ctl, P, d_old_v, lo_edge, hi_edge, slab.has_lower,
            slab.has_upper, migration, send_lo, send_hi, slab.capacity, d_X1,
            reinterpret_cast<float3*>(d_dX), slab.n_stay, slab.status[0],
            slab.status[1], slab.n_tiles,
            // migration: leave the cells that stay in the cube order of the
            // last force evaluation (cell identity is not tracked anyway)
            migration && slab.permute ? Computer<Pt>::dd_cube_order() : nullptr,
            d_n, n_max, slab.status[2]);
        YB_CUDA(cudaGetLastError());
    }
    void slab_unpack(int what, const float* recv_lo, const float* recv_hi)
    {
        const int blocks = yb::stride_grid(n_max, 256, yb::sm_count());
        if (what == 2) {
            yb::slab_merge<Pt><<<blocks, 256, 0, stream>>>(d_ctl, slab.n_stay,
                d_X1, reinterpret_cast<const float3*>(d_dX), recv_lo, recv_hi,
                slab.has_lower, slab.has_upper, n_max, d_X, d_old_v,
                slab.new_count);
            yb::slab_commit_count<<<1, 1, 0, stream>>>(
                d_ctl, slab.new_count, d_n);
        } else {
            yb::slab_append_ghosts<Pt><<<blocks, 256, 0, stream>>>(d_ctl,
                what == 1 ? d_X1 : d_X, d_old_v, recv_lo, recv_hi,
                slab.has_lower, slab.has_upper, n_max, d_n);
        }
        YB_CUDA(cudaGetLastError());
    }
    // Predictor / corrector with drift = (sums over all slabs) / (count over all
    // slabs); d_sums4 = {sum dX.x, .y, .z, n} after the caller's all-reduce.
    void slab_update(int stage, float dt, const float* d_sums4)
    {
        yb::slab_set_drift<<<1, 1, 0, stream>>>(d_ctl, stage, d_sums4);
        const int blocks = yb::stride_grid(n_max, 256, yb::sm_count());
        if (stage == 0)
            yb::predictor_step<Pt, false><<<blocks, 256, 0, stream>>>(d_n, n_max,
                dt, d_X, d_dX, d_X1, d_ctl, 1.f, yb::Grid_box{}, nullptr, nullptr, nullptr);
        else
            yb::corrector_step<Pt><<<blocks, 256, 0, stream>>>(
                d_n, n_max, dt, d_dX, d_dX1, d_X, d_old_v, d_ctl);
        YB_CUDA(cudaGetLastError());
    }
    // Blocking: owned cells, owned + ghost cells of the last halo round, and the
    // number of problems seen
    // (cells outside the grid; exchange-buffer overflows count 2^20 each).
    void slab_counts(int* n_owned, int* n_total, int* problems)
    {
        yb::Step_ctl snapshot;
        YB_CUDA(cudaMemcpyAsync(&snapshot, d_ctl, sizeof(snapshot),
            cudaMemcpyDeviceToHost, stream));
        const int n_now = get_d_n();  // waits for the stream
        // daughters appended since the last step are adopted by the next one
        if (dom.active && dom.grows && n_now > snapshot.n_owned)
            snapshot.n_owned = n_now;
        if (n_owned) *n_owned = snapshot.n_owned;
        // ghosts only live from a halo round to the end of its stage: report
        // how many the last round brought
        if (n_total) *n_total = snapshot.n_owned + snapshot.n_ghosts;
        if (problems) *problems = snapshot.out_of_grid;
    }

    // ---- Extension: brick decomposition over peer memory ----------------------
    // One solver per brick and GPU; halo exchange, migration and the global
    // drift sum are done by kernels that store into the neighbours' memory
    // (b200/domain.cuh). dom_begin lays out this rank's exchange allocation,
    // the caller maps the neighbours' allocations (CUDA IPC) and connects them,
    // dom_step then runs whole Heun steps without the host ever waiting.
    yb::Domain_link dom;

    // A per-cell array of the model (a Property's d_prop, curand states, ...)
    // that has to travel with the cells of a decomposed tissue: it migrates
    // with its cell, is re-stored with it when a migration round leaves the
    // cells in cube order, and -- ghosts_too, for arrays the pairwise functor
    // reads of a NEIGHBOUR, like a cell type -- comes along with the ghosts.
    // Entries are bytes_per_cell (a multiple of 4) wide and indexed like d_X;
    // register before dom_begin.
    bool dom_register_array(void* d_array, int bytes_per_cell, bool ghosts_too)
    {
        return dom.register_extra(d_array, bytes_per_cell, ghosts_too);
    }

    void dom_begin(int rank, int world, const float lo[3], const float hi[3],
        float halo, const int peer_ranks27[27], const int capacity27[27])
    {
        yb::Dd_region region{};
        for (int a = 0; a < 3; a++) region.lo[a] = lo[a], region.hi[a] = hi[a];
        region.halo = halo;
        dom.begin(rank, world, region, peer_ranks27, capacity27,
            yb::Layout<Pt>::lanes + 3, n_max);
        dd_set_counts(0, 0);
    }

    // Cells the model appended behind the owned ones since the last step
    // (division) become owned cells of this brick. dom_step does this itself;
    // models that survey the neighbourhood first (below) call it before.
    void dom_adopt()
    {
        const int blocks = yb::stride_grid(n_max, 256, yb::sm_count());
        yb::dd_flag_new_cells<Pt><<<blocks, 256, 0, stream>>>(
            d_ctl, d_n, n_max, d_X, dom.inset_faces(), dom.halo_flags);
        yb::dd_adopt_owned<<<1, 1, 0, stream>>>(d_ctl, d_n, n_max);
    }

    // For model kernels that look at the neighbourhood of the cells BETWEEN
    // steps (rewiring of protrusions, ...): an extra halo round of the current
    // positions and of the registered ghosts_too arrays. Afterwards *d_n counts
    // owned cells + ghosts and the ghosts sit behind the owned cells, as inside a
    // stage; dom_end_survey() drops them again (call it before dom_step).
    void dom_survey()
    {
        assert(dom.active && dom.connected());
        const int blocks = yb::stride_grid(n_max, 256, yb::sm_count());
        dom_round(3, d_X, d_old_v);
        yb::dd_append_ghosts<Pt><<<blocks, 256, 0, stream>>>(d_ctl, d_X, d_old_v,
            dom.inboxes(3), n_max, d_n, dom.extras, dom.halo_record_floats);
        YB_CUDA(cudaGetLastError());
    }
    void dom_end_survey()
    {
        yb::dd_drop_ghosts<<<1, 1, 0, stream>>>(d_ctl, d_n);
    }
    // n_owned / n_ghosts of this brick, in device memory (for model kernels)
    const yb::Step_ctl* dom_ctl() const { return d_ctl; }

    // Generic forces of a decomposed step must be capturable in the sense of
    // capture_generic_forces: they only enqueue work on
    // yb::current_stage()->stream and take the live count (owned cells + ghosts)
    // from d_n on the device; the n they are handed is n_max. Between two steps
    // the model may append cells behind the owned ones and bump *d_n (division):
    // the next step adopts them.
    template<Pairwise_interaction<Pt> pw_int, Pairwise_friction<Pt> pw_friction>
    void dom_step(float dt, Generic_forces<Pt> gen_forces = no_gen_forces<Pt>)
    {
        if (yb::is_no_gen_forces(gen_forces))
            dom_step_impl<pw_int, pw_friction, false>(dt, gen_forces);
        else
            dom_step_impl<pw_int, pw_friction, true>(dt, gen_forces);
    }

private:
    template<Pairwise_interaction<Pt> pw_int, Pairwise_friction<Pt> pw_friction,
        bool SEEDED>
    void dom_step_impl(float dt, Generic_forces<Pt>& gen_forces)
    {
        assert(dom.active && dom.connected());
        // With lazy module loading the first launch of a kernel may wait for the
        // device to drain -- while a dd_wait on it spins for a neighbour whose
        // kernels this very thread has yet to enqueue (bricks sharing a process)
        // or merely late (one process per GPU). Load everything up front.
        static const bool loaded = [] {
            yb::load_kernel(yb::dd_tile_counts<Pt>);
            yb::load_kernel(yb::dd_tile_offsets);
            yb::load_kernel(yb::dd_pack<Pt>);
            yb::load_kernel(yb::dd_push);
            yb::load_kernel(yb::dd_wait);
            yb::load_kernel(yb::dd_append_ghosts<Pt>);
            yb::load_kernel(yb::dd_merge<Pt>);
            yb::load_kernel(yb::dd_restore_extras);
            yb::load_kernel(yb::bin_cells_part<Pt>);
            yb::load_kernel(yb::dd_flag_new_cells<Pt>);
            yb::load_kernel(yb::dd_adopt_owned);
            yb::load_kernel(yb::dd_drop_ghosts);
            yb::load_kernel(yb::zero_cells<Pt>);
            yb::load_kernel(yb::dd_allreduce_drift);
            yb::load_kernel(yb::slab_commit_count);
            yb::load_kernel(yb::predictor_step<Pt, false>);
            yb::load_kernel(yb::corrector_step<Pt>);
            Computer<Pt>::template load_kernels<pw_int, pw_friction, SEEDED>();
            return true;
        }();
        (void)loaded;
        const int blocks = yb::stride_grid(n_max, 256, yb::sm_count());
        ++step_serial;
        dom_mark(-1);
        if (dom.grows) dom_adopt();
        for (int stage = 0; stage < 2; stage++) {
            Pt* X_stage = stage == 0 ? d_X : d_X1;
            Pt* dX_stage = stage == 0 ? d_dX : d_dX1;
            const bool binned = dom_round(stage, X_stage, d_old_v);
            yb::dd_append_ghosts<Pt><<<blocks, 256, 0, stream>>>(d_ctl, X_stage,
                d_old_v, dom.inboxes(stage), n_max, d_n, dom.extras,
                dom.halo_record_floats);
            if (binned) Computer<Pt>::bin_part(stream, d_n, X_stage, d_ctl, 1);
            dom_mark(2);
            if (SEEDED) {
                yb::zero_cells<Pt><<<blocks, 256, 0, stream>>>(d_n, n_max, dX_stage);
                yb::Stage_context context{n_max, n_max, stream};
                context.d_n_cells = d_n;
                context.stage = stage;
                context.capturing = false;
                // (no solver identity: indices of ghosts differ between the
                // stages, so nothing a force builds in stage 0 holds in stage 1)
                context.solver = nullptr;
                context.step_serial = step_serial;
                context.eager_stream = stream;
                context.hooks = nullptr;
                yb::current_stage() = &context;
                gen_forces(n_max, X_stage, dX_stage);
                yb::current_stage() = nullptr;
            }
            cudaEvent_t sweep_start = nullptr, sweep_stop = nullptr;
            if (profiling) {
                YB_CUDA(cudaEventCreate(&sweep_start));
                YB_CUDA(cudaEventCreate(&sweep_stop));
            }
            Computer<Pt>::template pwints<pw_int, pw_friction, SEEDED>(stream, d_n,
                X_stage, d_old_v, dX_stage, d_partials, max_sweep_ctas, stage,
                yb::DRIFT_MEAN, 0, d_ctl, binned, sweep_start);
            if (profiling) {
                YB_CUDA(cudaEventRecord(sweep_stop, stream));
                sweep_events.emplace_back(sweep_start, sweep_stop);
            }
            dom_mark(3);
            const unsigned epoch = ++dom.drift_epoch;
            yb::dd_allreduce_drift<<<1, yb::DD_MAX_RANKS, 0, stream>>>(d_ctl,
                stage, dom.mailboxes, dom.my_mailbox(), dom.rank, dom.world,
                static_cast<int>(epoch & 1u), epoch);
            dom_mark(4);
            if (stage == 0)
                yb::predictor_step<Pt, false><<<blocks, 256, 0, stream>>>(d_n,
                    n_max, dt, d_X, d_dX, d_X1, d_ctl, 1.f, yb::Grid_box{},
                    nullptr, nullptr, nullptr, dom.inset_faces(), dom.halo_flags);
            else  // the new state goes to X1 / v_new: the migration pass below
                  // re-stores it in X / old_v in one go
                yb::corrector_step<Pt><<<blocks, 256, 0, stream>>>(d_n, n_max, dt,
                    d_dX, d_dX1, d_X, d_old_v, d_ctl, d_X1, dom.v_new);
            dom_mark(5);
        }
        // migration: cells that crossed a face change owner; the others are
        // re-stored in the cube order of the last force evaluation
        dom_round(2, d_X1, dom.v_new);
        yb::dd_merge<Pt><<<yb::stride_grid(n_max / 16 + 1, 256, yb::sm_count()), 256,
            0, stream>>>(d_ctl, dom.n_stay, dom.inboxes(2), n_max, d_X, d_old_v,
            dom.new_count, dom.inset_faces(), dom.halo_flags, dom.extras,
            dom.record_floats);
        if (dom.extras.count > 0)
            yb::dd_restore_extras<<<blocks, 256, 0, stream>>>(dom.n_stay, dom.extras);
        yb::slab_commit_count<<<1, 1, 0, stream>>>(d_ctl, dom.new_count, d_n);
        dom.flags_valid = true;
        dom_mark(2);
        YB_CUDA(cudaGetLastError());
    }

public:

    // Extension: where a decomposed step spends its time, by CUDA events on the
    // stream while profile_sweeps is on. Milliseconds since the last read for
    // {select (packing the two halo rounds), wait (for the neighbours' flags),
    // unpack, forces (grid build + sweep), drift sum, update, push (outboxes ->
    // neighbours), migration select (with the re-store in cube order)}.
    static constexpr int DOM_PHASES = 8;
    void read_dom_profile(float* ms8)
    {
        float* ms6 = ms8;
        YB_CUDA(cudaStreamSynchronize(stream));
        for (int q = 0; q < DOM_PHASES; q++) ms6[q] = 0.f;
        for (size_t k = 1; k < dom_marks.size(); k++) {
            if (dom_marks[k].second < 0) continue;  // start of a step
            float ms = 0.f;
            YB_CUDA(cudaEventElapsedTime(
                &ms, dom_marks[k - 1].first, dom_marks[k].first));
            ms6[dom_marks[k].second] += ms;
        }
        for (auto& mark : dom_marks) cudaEventDestroy(mark.first);
        dom_marks.clear();
    }

    // Fill this brick with its share of a seeded, jittered FCC ball. Returns the
    // number of cells the share has; if that exceeds n_max, only n_max are kept.
    int dom_seed_lattice_ball(
        float radius, float dist_to_nb, float jitter, unsigned long long seed)
    {
        assert(dom.active);
        const float a = dist_to_nb * 1.41421356237f;
        const int half = static_cast<int>(ceilf(radius / a)) + 1;
        const long long side = 2 * half + 1;
        const long long n_sites = 4 * side * side * side;
        YB_CUDA(cudaMemsetAsync(d_n, 0, sizeof(int), stream));
        yb::dd_seed_lattice_ball<Pt><<<yb::sm_count() * 16, 256, 0, stream>>>(
            radius, dist_to_nb, jitter, seed, dom.region, half, n_sites, n_max,
            d_X, d_old_v, d_n);
        YB_CUDA(cudaMemcpyAsync(
            h_n_pinned, d_n, sizeof(int), cudaMemcpyDeviceToHost, stream));
        YB_CUDA(cudaStreamSynchronize(stream));
        const int wanted = *h_n_pinned;
        const int n = wanted < n_max ? wanted : n_max;
        dd_set_counts(n, n);
        dom.flags_valid = false;
        return wanted;
    }

private:
    // one exchange round: what 0 / 1 = halo of X / X1, 2 = migration
    // (cells at P with velocities v; a migration round re-stores the cells that
    // stay in d_X / d_old_v)
    // Returns true if the cube ids of the owned cells were computed on the way
    // (halo rounds: while the outboxes travel, see below).
    bool dom_round(int what, const Pt* P, const float3* v)
    {
        const bool migration = what == 2;
        const unsigned epoch = ++dom.epoch[what];
        const int width = migration ? dom.record_floats : dom.halo_record_floats;
        // a brick without neighbours has no halo; its migration round still
        // re-stores the cells in cube order
        if (dom.region.n_peers == 0 && !migration) return false;
        // the local outboxes are free once the previous round's push has read them
        if (dom.push_pending) {
            YB_CUDA(cudaStreamWaitEvent(stream, dom.pushed, 0));
            dom.push_pending = false;
        }
        const float4* order =
            migration && dom.permute ? Computer<Pt>::dd_cube_order() : nullptr;
        // stage 0: flags from the last dd_merge (not before the first step);
        // stage 1: from the predictor just now
        const unsigned char* flags =
            (what == 1 || (what != 2 && dom.flags_valid)) ? dom.halo_flags : nullptr;
        const int n_lists = dom.region.n_peers + (order != nullptr ? 1 : 0);
        yb::dd_tile_counts<Pt><<<dom.n_tiles, yb::SCAN_THREADS, 0, stream>>>(d_ctl,
            P, dom.region, migration, order, d_n, n_max, flags, dom.tile_counts,
            dom.n_tiles);
        if (n_lists > 0)
            yb::dd_tile_offsets<<<n_lists, 1024, 0, stream>>>(d_ctl,
                dom.tile_counts, dom.n_tiles, dom.local_out, dom.region.n_peers,
                dom.totals);
        yb::dd_pack<Pt><<<dom.n_tiles, yb::SCAN_THREADS, 0, stream>>>(d_ctl, P, v,
            dom.region, dom.local_out, migration, d_X, d_old_v, dom.n_stay, order,
            d_n, n_max, flags, dom.tile_counts, dom.totals, dom.n_tiles,
            dom.inset_faces(), dom.halo_flags, dom.extras, width);
        dom_mark(migration ? 7 : 0);
        bool binned = false;
        if (dom.region.n_peers > 0) {
            const int push_grid = dom.region.n_peers * yb::dd_push_ctas();
            if (what < 2 && dom.overlap) {
                // Halo exchange overlapped with interior work: the push runs on
                // a side stream (the neighbours' pushes arrive on theirs) while
                // this stream starts the grid build of the stage with the cells
                // it owns -- their cube ids and the per-cube counts. The ghosts
                // are binned when they have arrived.
                YB_CUDA(cudaEventRecord(dom.packed, stream));
                YB_CUDA(cudaStreamWaitEvent(dom.push_stream, dom.packed, 0));
                yb::dd_push<<<push_grid, 256, 0, dom.push_stream>>>(dom.local_out,
                    dom.out[what], dom.region.n_peers, yb::dd_push_ctas(), width,
                    dom.push_done, epoch);
                YB_CUDA(cudaEventRecord(dom.pushed, dom.push_stream));
                dom.push_pending = true;
                Computer<Pt>::bin_part(stream, d_n, P, d_ctl, 0);
                binned = true;
            } else {
                yb::dd_push<<<push_grid, 256, 0, stream>>>(dom.local_out,
                    dom.out[what], dom.region.n_peers, yb::dd_push_ctas(), width,
                    dom.push_done, epoch);
            }
            dom_mark(6);
            yb::dd_wait<<<1, 32, 0, stream>>>(d_ctl, dom.inboxes(what), epoch);
            dom_mark(1);
        }
        return binned;
    }
    void dom_mark(int phase)
    {
        if (!profiling) return;
        cudaEvent_t event;
        YB_CUDA(cudaEventCreate(&event));
        YB_CUDA(cudaEventRecord(event, stream));
        dom_marks.emplace_back(event, phase);
    }
    std::vector<std::pair<cudaEvent_t, int>> dom_marks;

public:
    // Extension: time the pairwise sweep kernels with CUDA events on the
    // launching stream (bench.py's roofline needs the dominant kernel's own
    // duration). While enabled, steps are issued directly instead of replayed
    // from a graph. read_sweep_profile() waits for the stream and returns the
    // accumulated milliseconds and number of sweep launches since the last read.
    void profile_sweeps(bool enable)
    {
        read_sweep_profile(nullptr, nullptr);
        profiling = enable;
    }
    void read_sweep_profile(float* total_ms, int* launches)
    {
        YB_CUDA(cudaStreamSynchronize(stream));
        float sum = 0.f;
        for (auto& pair : sweep_events) {
            float ms = 0.f;
            YB_CUDA(cudaEventElapsedTime(&ms, pair.first, pair.second));
            sum += ms;
            cudaEventDestroy(pair.first);
            cudaEventDestroy(pair.second);
        }
        if (total_ms) *total_ms = sum;
        if (launches) *launches = static_cast<int>(sweep_events.size());
        sweep_events.clear();
    }

protected:
    Pt *d_X, *d_dX, *d_X1, *d_dX1;
    float3* d_old_v;
    int* d_n;
    bool fix_com = true;
    bool fix_com_z = false;
    int fix_point = 0;
    const int n_max;

    int get_d_n()
    {
        // through a pinned word: a plain 4-byte cudaMemcpy to pageable memory
        // costs several times the round trip
        YB_CUDA(cudaMemcpyAsync(
            h_n_pinned, d_n, sizeof(int), cudaMemcpyDeviceToHost, stream));
        YB_CUDA(cudaStreamSynchronize(stream));
        const int n = *h_n_pinned;
        assert(n <= n_max);
        return n;
    }

    template<Pairwise_interaction<Pt> pw_int, Pairwise_friction<Pt> pw_friction>
    void take_step(float dt, Generic_forces<Pt> gen_forces)
    {
        // Drift selection per stage, exactly as solvers.cuh:241-253, :266-272.
        const int mode0 =
            (fix_com || fix_com_z)
                ? (fix_com_z ? yb::DRIFT_POINT_XY_MEAN_Z : yb::DRIFT_MEAN)
                : yb::DRIFT_POINT;
        const int mode1 = fix_com ? yb::DRIFT_MEAN : yb::DRIFT_POINT;
        const bool seeded = !yb::is_no_gen_forces(gen_forces);
        ++step_serial;

        // occupancy queries and attribute changes must not happen mid-capture
        if (seeded)
            Computer<Pt>::template prepare<pw_int, pw_friction, true>();
        else
            Computer<Pt>::template prepare<pw_int, pw_friction, false>();

        // The caller is recording its own graph on the solver's stream (e.g. a
        // whole model iteration): the stages join it. Generic forces must then
        // be capturable, see capture_generic_forces.
        cudaStreamCaptureStatus capture = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(stream, &capture) != cudaSuccess) {
            cudaGetLastError();
            capture = cudaStreamCaptureStatusNone;
        }
        if (capture == cudaStreamCaptureStatusActive) {
            enqueue_step<pw_int, pw_friction>(stream, dt, mode0, mode1, seeded,
                n_max, gen_forces, true, nullptr);
            return;
        }

        const bool replayable = yb::graphs_enabled() && !profiling &&
                                (!seeded || capture_generic_forces);
        if (!replayable) {
            int n = 0;
            if (seeded) {
                // The callback is arbitrary host code that needs n: one read
                // of d_n per step, then the stage kernels are issued directly.
                // The read goes through a side stream while the grid build of
                // the first stage (which needs neither n on the host nor the
                // callback's output) already runs, so the device does not idle
                // during the round trip.
                YB_CUDA(cudaEventRecord(count_ready, stream));
                YB_CUDA(cudaStreamWaitEvent(capture_stream, count_ready, 0));
                YB_CUDA(cudaMemcpyAsync(h_n_pinned, d_n, sizeof(int),
                    cudaMemcpyDeviceToHost, capture_stream));
                Computer<Pt>::index_ahead(stream, d_n, d_X, d_old_v, d_ctl);
                YB_CUDA(cudaStreamSynchronize(capture_stream));
                n = *h_n_pinned;
                assert(n <= n_max);
            }
            enqueue_step<pw_int, pw_friction>(stream, dt, mode0, mode1, seeded,
                n, gen_forces, false, nullptr);
            return;
        }

        static const char instantiation_tag = 0;
        const void* forces_type = seeded ? &gen_forces.target_type() : nullptr;
        const yb::Graph_key solver_key = Computer<Pt>::graph_key();
        for (size_t k = 0; k < graphs.size(); k++) {
            auto& g = graphs[k];
            if (g.instantiation != &instantiation_tag ||
                g.forces_type != forces_type || g.dt != dt ||
                !(g.solver_key == solver_key) || g.mode0 != mode0 ||
                g.mode1 != mode1 || g.fix_point != fix_point)
                continue;
            bool valid = true;
            for (auto& hook : g.hooks) valid = hook(stream) && valid;
            if (valid) {
                YB_CUDA(cudaGraphLaunch(g.exec, stream));
                return;
            }
            cudaGraphExecDestroy(g.exec);  // a force asked for a new capture
            graphs.erase(graphs.begin() + k);
            break;
        }
        yb::Step_graph entry{&instantiation_tag, forces_type, dt, solver_key,
            mode0, mode1, fix_point, nullptr, {}};
        cudaGraph_t graph;
        // Relaxed: capturable forces may allocate scratch while being recorded.
        YB_CUDA(cudaStreamBeginCapture(
            capture_stream, cudaStreamCaptureModeRelaxed));
        enqueue_step<pw_int, pw_friction>(capture_stream, dt, mode0, mode1,
            seeded, n_max, gen_forces, true, &entry.hooks);
        YB_CUDA(cudaStreamEndCapture(capture_stream, &graph));
        YB_CUDA(cudaGraphInstantiate(&entry.exec, graph, 0));
        YB_CUDA(cudaGraphDestroy(graph));
        YB_CUDA(cudaGraphLaunch(entry.exec, stream));
        graphs.push_back(std::move(entry));
    }

public:
    // Extension: replay steps WITH generic forces from a captured CUDA graph
    // too. Set it if the callable passed to take_step is "capturable": it
    // only enqueues work on the step's stream (yb::current_stage()->stream,
    // i.e. this solver's `stream` or the one recording it), enqueues the same
    // work every call, and reads the live cell count from d_n on the device --
    // the n it is handed is then n_max, an upper bound. link_forces,
    // wall_forces and plain kernel launches / cudaMemsetAsync on that stream
    // qualify; thrust calls and anything that reads device memory back do not.
    // Graphs are keyed by the callable's type (one lambda expression = one
    // graph), not by the values it captured.
    bool capture_generic_forces = false;

    // Drop every captured step (e.g. after changing something a capturable
    // generic force baked into its launches).
    void reset_graphs()
    {
        YB_CUDA(cudaStreamSynchronize(stream));
        for (auto& g : graphs) cudaGraphExecDestroy(g.exec);
        graphs.clear();
    }

private:
    yb::Step_ctl* d_ctl;
    float* d_partials;
    int max_sweep_ctas;
    cudaStream_t capture_stream;
    std::vector<yb::Step_graph> graphs;
    int* h_n_pinned = nullptr;
    cudaEvent_t count_ready = nullptr;
    yb::Slab_scratch slab;
    bool profiling = false;
    unsigned long long step_serial = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> sweep_events;

    template<Pairwise_interaction<Pt> pw_int, Pairwise_friction<Pt> pw_friction>
    void enqueue_step(cudaStream_t s, float dt, int mode0, int mode1,
        bool seeded, int n, Generic_forces<Pt>& gen_forces, bool capturing,
        std::vector<yb::Replay_hook>* hooks)
    {
        if (seeded) {
            enqueue_stage<pw_int, pw_friction, true>(
                s, 0, dt, mode0, n, gen_forces, capturing, hooks);
            enqueue_stage<pw_int, pw_friction, true>(
                s, 1, dt, mode1, n, gen_forces, capturing, hooks);
        } else {
            enqueue_stage<pw_int, pw_friction, false>(
                s, 0, dt, mode0, n, gen_forces, capturing, hooks);
            enqueue_stage<pw_int, pw_friction, false>(
                s, 1, dt, mode1, n, gen_forces, capturing, hooks);
        }
    }

    // One Heun stage: [seed dX with generic forces] -> pairwise sweep (writes
    // dX or dX1 and the stage's drift) -> predictor or corrector update.
    template<Pairwise_interaction<Pt> pw_int, Pairwise_friction<Pt> pw_friction,
        bool SEEDED>
    void enqueue_stage(cudaStream_t s, int stage, float dt, int drift_mode,
        int n, Generic_forces<Pt>& gen_forces, bool capturing,
        std::vector<yb::Replay_hook>* hooks)
    {
        const Pt* X_stage = stage == 0 ? d_X : d_X1;
        Pt* dX_stage = stage == 0 ? d_dX : d_dX1;
        const int blocks = yb::stride_grid(n_max, 256, yb::sm_count());
        if (SEEDED) {
            if (capturing)
                yb::zero_cells<Pt><<<blocks, 256, 0, s>>>(d_n, n_max, dX_stage);
            else
                YB_CUDA(cudaMemsetAsync(
                    dX_stage, 0, static_cast<size_t>(n) * sizeof(Pt), s));
            // lets link_forces & co. see the cell count and the stream
            yb::Stage_context context{n, n_max, s};
            context.d_n_cells = d_n;
            context.stage = stage;
            context.capturing = capturing;
            context.solver = this;
            context.step_serial = step_serial;
            context.eager_stream = stream;
            context.hooks = hooks;
            yb::current_stage() = &context;
            gen_forces(n, X_stage, dX_stage);
            yb::current_stage() = nullptr;
        }
        const bool binned_by_predictor = stage == 1;
        cudaEvent_t sweep_start = nullptr, sweep_stop = nullptr;
        if (profiling) {
            YB_CUDA(cudaEventCreate(&sweep_start));
            YB_CUDA(cudaEventCreate(&sweep_stop));
        }
        Computer<Pt>::template pwints<pw_int, pw_friction, SEEDED>(s, d_n,
            X_stage, d_old_v, dX_stage, d_partials, max_sweep_ctas, stage,
            drift_mode, fix_point, d_ctl, binned_by_predictor, sweep_start);
        if (profiling) {
            YB_CUDA(cudaEventRecord(sweep_stop, s));
            sweep_events.emplace_back(sweep_start, sweep_stop);
        }
        if (stage == 0) {
            Computer<Pt>::predict(s, blocks, d_n, dt, d_X, d_dX, d_X1, d_ctl);
        } else {
            yb::corrector_step<Pt><<<blocks, 256, 0, s>>>(
                d_n, n_max, dt, d_dX, d_dX1, d_X, d_old_v, d_ctl);
        }
        YB_CUDA(cudaGetLastError());
    }
};


// ---- all pairs ---------------------------------------------------------------------
// Kept for mesh.cuh and user code that sizes launches with it.
const auto TILE_SIZE = 32;

template<typename Pt>
class Tile_computer {
public:
    Tile_computer(int n_max) : n_max{n_max} {}

    // Extension: split the partners of every cell across 8 or 32 lanes
    // (b200/pair_sweep.cuh, sweep_tiles_split). Much faster for the few hundred
    // to few thousand cells the Tile solver is used with, but the pairwise
    // functor then runs in several threads per cell: only for functors without
    // per-cell side effects. Off by default.
    bool split_pairs = false;

protected:
    yb::Graph_key graph_key() const
    {
        return yb::Graph_key{0.f, 0.f, split_pairs ? 1 : 0, 0};
    }

    const float4* dd_cube_order() const { return nullptr; }

    template<Pairwise_interaction<Pt> pw_int, Pairwise_friction<Pt> pw_friction,
        bool SEEDED>
    static void load_kernels()
    {}

    // nothing to prepare ahead of the generic forces
    void index_ahead(
        cudaStream_t, const int*, const Pt*, const float3*, yb::Step_ctl*)
    {}
    void bin_part(cudaStream_t, const int*, const Pt*, yb::Step_ctl*, int) {}

    template<Pairwise_interaction<Pt> pw_int, Pairwise_friction<Pt> pw_friction,
        bool SEEDED>
    void prepare()
    {}

    template<Pairwise_interaction<Pt> pw_int, Pairwise_friction<Pt> pw_friction,
        bool SEEDED>
    void pwints(cudaStream_t s, const int* d_n, const Pt* d_X,
        const float3* d_old_v, Pt* d_dX, float* d_partials, int max_ctas,
        int stage, int drift_mode, int fix_point, yb::Step_ctl* d_ctl,
        bool /*binned_by_predictor*/, cudaEvent_t before_sweep = nullptr)
    {
        if (before_sweep) YB_CUDA(cudaEventRecord(before_sweep, s));
        const int cells = n_max > 0 ? n_max : 1;
        if (split_pairs && cells <= 2048) {
            launch_split<pw_int, pw_friction, SEEDED, 32>(s, cells, max_ctas, d_n,
                d_X, d_old_v, d_dX, d_partials, stage, drift_mode, fix_point,
                d_ctl);
        } else if (split_pairs && cells <= 16384) {
            launch_split<pw_int, pw_friction, SEEDED, 8>(s, cells, max_ctas, d_n,
                d_X, d_old_v, d_dX, d_partials, stage, drift_mode, fix_point,
                d_ctl);
        } else {
            int blocks = yb::ceil_div(cells, yb::TILE_THREADS);
            if (blocks > max_ctas) blocks = max_ctas;
            yb::sweep_tiles<Pt, pw_int, pw_friction, SEEDED>
                <<<blocks, yb::TILE_THREADS, 0, s>>>(d_n, n_max, d_X, d_old_v,
                    d_dX, d_partials, stage, drift_mode, fix_point, d_ctl);
        }
    }

    template<Pairwise_interaction<Pt> pw_int, Pairwise_friction<Pt> pw_friction,
        bool SEEDED, int G>
    void launch_split(cudaStream_t s, int cells, int max_ctas, const int* d_n,
        const Pt* d_X, const float3* d_old_v, Pt* d_dX, float* d_partials,
        int stage, int drift_mode, int fix_point, yb::Step_ctl* d_ctl)
    {
        int blocks = yb::ceil_div(cells, yb::TILE_THREADS / G);
        if (blocks > max_ctas) blocks = max_ctas;
        yb::sweep_tiles_split<Pt, pw_int, pw_friction, SEEDED, G>
            <<<blocks, yb::TILE_THREADS, 0, s>>>(d_n, n_max, d_X, d_old_v, d_dX,
                d_partials, stage, drift_mode, fix_point, d_ctl);
    }

    void predict(cudaStream_t s, int blocks, const int* d_n, float dt,
        const Pt* d_X, const Pt* d_dX, Pt* d_X1, yb::Step_ctl* d_ctl)
    {
        yb::predictor_step<Pt, false><<<blocks, 256, 0, s>>>(d_n, n_max, dt, d_X,
            d_dX, d_X1, d_ctl, 1.f, yb::Grid_box{}, nullptr, nullptr, nullptr);
    }

private:
    const int n_max;
};

template<typename Pt>
using Tile_solver = Heun_solver<Pt, Tile_computer>;


// ---- neighbour grid ----------------------------------------------------------------
// Offsets from a cube to its 27 neighbours (x fastest, then y, then z; within
// each: -1, 0, +1 resp. 0, -, +), filled in by the grid solvers' constructors.
// User kernels that pick random neighbour cubes read it (reference :428).
__constant__ int d_nhood[27];

namespace yb {
inline void upload_nhood(int grid_size)
{
    int h_nhood[27];
    const int step[3] = {0, -1, 1};
    for (int z = 0; z < 3; z++)
        for (int y = 0; y < 3; y++)
            for (int x = 0; x < 3; x++)
                h_nhood[9 * z + 3 * y + x] = (x - 1) + step[y] * grid_size +
                                             step[z] * grid_size * grid_size;
    YB_CUDA(cudaMemcpyToSymbol(d_nhood, h_nhood, sizeof(h_nhood)));
}

// Scratch shared by the public Grid and the grid solvers: the bucket sort.
struct Bucket_sort {
    int* key = nullptr;       // cube id per cell, original order
    int* arrival = nullptr;   // arrival rank within the cube
    int* slot_id = nullptr;   // cell id per slot, arrival order within cubes
    int* count = nullptr;     // per cube, zero between builds
    int* offset = nullptr;    // exclusive prefix sum, n_cubes + 1 (+ padding)
    unsigned long long* status = nullptr;  // look-back words of the scan
    int n_tiles = 0;

    void allocate(int n_max, int n_cubes)
    {
        const size_t cells = static_cast<size_t>(n_max > 0 ? n_max : 1);
        const size_t bins = static_cast<size_t>(scan_padded(n_cubes + 1));
        n_tiles = static_cast<int>(bins / SCAN_TILE);
        YB_CUDA(cudaMalloc(&key, cells * sizeof(int)));
        YB_CUDA(cudaMalloc(&arrival, cells * sizeof(int)));
        YB_CUDA(cudaMalloc(&slot_id, cells * sizeof(int)));
        YB_CUDA(cudaMalloc(&count, bins * sizeof(int)));
        YB_CUDA(cudaMemset(count, 0, bins * sizeof(int)));
        YB_CUDA(cudaMalloc(&offset, bins * sizeof(int)));
        YB_CUDA(cudaMemset(offset, 0, bins * sizeof(int)));
        YB_CUDA(cudaMalloc(&status, n_tiles * sizeof(unsigned long long)));
        YB_CUDA(cudaMemset(status, 0, n_tiles * sizeof(unsigned long long)));
    }
    void release()
    {
        cudaFree(status);
        cudaFree(offset);
        cudaFree(count);
        cudaFree(slot_id);
        cudaFree(arrival);
        cudaFree(key);
    }
};
}  // namespace yb


// Stand-alone neighbour grid for user kernels (e.g. protrusion updates pick
// random cells from neighbouring cubes). After build():
//   d_cube_id[k]   cube id of sorted slot k (ascending)
//   d_point_id[k]  original cell id in slot k (ascending inside a cube)
//   d_cube_start[c], d_cube_end[c]  inclusive slot range of cube c, or -1, -2
// bit-identical to the reference's arrays (solvers.cuh:380-425).
class Grid {
public:
    int *d_cube_id, *d_point_id, *d_cube_start, *d_cube_end;
    Grid* d_grid;
    const int n_max, grid_size, n_cubes;
    // Extension: the stream build() works on (default: the legacy stream, like
    // the reference). Not part of the device copy's layout contract.
    cudaStream_t stream = 0;

    Grid(int n_max, int gs = 50)
        : n_max{n_max}, grid_size{gs}, n_cubes{gs * gs * gs}
    {
        const size_t cells = static_cast<size_t>(n_max > 0 ? n_max : 1);
        YB_CUDA(cudaMalloc(&d_cube_id, cells * sizeof(int)));
        YB_CUDA(cudaMalloc(&d_point_id, cells * sizeof(int)));
        YB_CUDA(cudaMalloc(&d_cube_start, n_cubes * sizeof(int)));
        YB_CUDA(cudaMalloc(&d_cube_end, n_cubes * sizeof(int)));
        sort.allocate(n_max, n_cubes);
        YB_CUDA(cudaMalloc(&d_ctl, sizeof(yb::Step_ctl)));
        yb::Step_ctl fresh{};
        fresh.scan_epoch = 1;
        YB_CUDA(cudaMemcpy(
            d_ctl, &fresh, sizeof(fresh), cudaMemcpyHostToDevice));
        YB_CUDA(cudaMalloc(&d_n_scratch, sizeof(int)));

        YB_CUDA(cudaMalloc(&d_grid, sizeof(Grid)));
        YB_CUDA(cudaMemcpy(d_grid, this, sizeof(Grid), cudaMemcpyHostToDevice));
    }
    Grid(const Grid&) = delete;
    Grid& operator=(const Grid&) = delete;
    ~Grid()
    {
        cudaFree(d_grid);
        cudaFree(d_n_scratch);
        cudaFree(d_ctl);
        sort.release();
        cudaFree(d_cube_end);
        cudaFree(d_cube_start);
        cudaFree(d_point_id);
        cudaFree(d_cube_id);
    }

    template<typename Pt>
    void build(
        const int n, const Pt* __restrict__ d_X, const float cube_size = 1)
    {
        assert(n <= n_max);
        YB_CUDA(cudaMemcpyAsync(
            d_n_scratch, &n, sizeof(int), cudaMemcpyHostToDevice, stream));
        build_live(d_n_scratch, d_X, cube_size);
    }
    // Extension: the same with the number of points read from device memory
    // (e.g. Solution::d_n) -- no host round trip, capturable into a CUDA graph.
    template<typename Pt>
    void build_live(const int* d_n_points, const Pt* __restrict__ d_X,
        const float cube_size = 1)
    {
        const cudaStream_t s = stream;
        const int sms = yb::sm_count();
        const int blocks = yb::stride_grid(n_max, 256, sms);
        yb::fill_cube_ranges<<<yb::stride_grid(n_cubes, 256, sms), 256, 0, s>>>(
            n_cubes, d_cube_start, d_cube_end);
        yb::bin_cells<Pt><<<blocks, 256, 0, s>>>(d_n_points, n_max, d_X,
            cube_size, yb::Grid_box::cubic(grid_size), sort.key, sort.arrival,
            sort.count, d_ctl);
        yb::scan_bins<<<sort.n_tiles, yb::SCAN_THREADS, 0, s>>>(
            sort.count, sort.offset, sort.n_tiles, sort.status, d_ctl);
        yb::place_ids<<<blocks, 256, 0, s>>>(d_n_points, n_max, sort.key,
            sort.arrival, sort.offset, sort.slot_id);
        yb::publish_grid<<<blocks, 256, 0, s>>>(d_n_points, n_max, sort.key,
            sort.offset, sort.slot_id, d_cube_id, d_point_id, d_cube_start,
            d_cube_end);
        YB_CUDA(cudaGetLastError());
    }
    template<typename Pt, template<typename> class Solver>
    void build(Solution<Pt, Solver>& points, const float cube_size = 1)
    {
        auto n = points.get_d_n();
        assert(n <= n_max);
        build(n, points.d_X, cube_size);
    }

private:
    yb::Bucket_sort sort;
    yb::Step_ctl* d_ctl;
    int* d_n_scratch;
};


// Pairwise sums over the 27 neighbouring cubes ONLY for cells closer than
// cube_size; scales linearly in n (reference: Grid_computer, :465-499).
template<typename Pt>
class Grid_computer {
public:
    float cube_size;  // may be changed between steps

    Grid_computer(int n_max, int grid_size = 50, float cube_size = 1)
        : cube_size{cube_size}, n_max{n_max}, grid_size{grid_size},
          n_cubes{grid_size * grid_size * grid_size},
          box(yb::Grid_box::cubic(grid_size))
    {
        yb::upload_nhood(grid_size);
        sort.allocate(n_max, n_cubes);
        const size_t cells = static_cast<size_t>(n_max > 0 ? n_max : 1);
        YB_CUDA(cudaMalloc(&pos4, cells * sizeof(float4)));
        YB_CUDA(cudaMalloc(
            &aux, cells * yb::Layout<Pt>::aux_vec4 * sizeof(float4)));
        YB_CUDA(cudaMalloc(&cube_sorted, cells * sizeof(int)));
        if (split_sweep()) allocate_lists();
        // scratch of the state-carrying build tail only
        if (carry_state())
            YB_CUDA(cudaMalloc(&staged,
                cells * (1 + yb::Layout<Pt>::aux_vec4) * sizeof(float4)));
    }
    Grid_computer(const Grid_computer&) = delete;
    Grid_computer& operator=(const Grid_computer&) = delete;
    ~Grid_computer()
    {
        cudaFree(nb_order);
        cudaFree(nb_count);
        cudaFree(nb);
        cudaFree(staged);
        cudaFree(cube_sorted);
        cudaFree(aux);
        cudaFree(pos4);
        sort.release();
    }

protected:
    // neighbour lists of list_cubes (the split sweep, the Gabriel solver)
    void allocate_lists()
    {
        if (nb != nullptr) return;
        nb_stride = (n_max > 0 ? n_max : 1) + 31 & ~31;
        YB_CUDA(cudaMalloc(&nb, size_t(nb_stride) * yb::LIST_MAX * sizeof(int)));
        YB_CUDA(cudaMalloc(&nb_count, size_t(nb_stride) * sizeof(int)));
        YB_CUDA(cudaMalloc(&nb_order, size_t(nb_stride) + yb::SWEEP_THREADS));
    }

    // everything the stage kernels get by value
    yb::Graph_key graph_key() const
    {
        return yb::Graph_key{
            cube_size, 0.f, (box.z_half * 1024 + box.y_half) * 1024 + box.x_half,
            box.n_cubes};
    }

    // Resident CTAs per SM of a persistent sweep kernel (queried once).
    template<typename Kernel>
    static int resident_ctas(Kernel kernel, int threads, size_t smem)
    {
        // Tuning knobs (environment): YALLA_B200_SWEEP_CTAS caps the CTAs per SM
        // the persistent grid is sized for; YALLA_B200_SWEEP_CARVEOUT sets the
        // shared-memory share of the L1 in percent (what is left caches the
        // gathers of the functor and of the accepted pairs).
        YB_CUDA(cudaFuncSetAttribute(kernel,
            cudaFuncAttributePreferredSharedMemoryCarveout,
            cudaSharedmemCarveoutMaxShared));
        int resident = 0;
        YB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &resident, kernel, threads, smem));
        const char* cap_env = getenv("YALLA_B200_SWEEP_CTAS");
        if (cap_env && cap_env[0] && atoi(cap_env) > 0 && atoi(cap_env) < resident)
            resident = atoi(cap_env);
        if (resident < 1) resident = 1;
        // Ask for no more shared memory than the resident CTAs need (+1 KB each
        // that the driver reserves): the rest of the 256 KB stays L1.
        const char* carveout_env = getenv("YALLA_B200_SWEEP_CARVEOUT");
        int carveout = static_cast<int>(
            (resident * (smem + 1024 + 512) * 100 + 228 * 1024 - 1) / (228 * 1024));
        if (carveout > 100) carveout = 100;
        if (carveout_env && carveout_env[0]) carveout = atoi(carveout_env);
        YB_CUDA(cudaFuncSetAttribute(kernel,
            cudaFuncAttributePreferredSharedMemoryCarveout, carveout));
        return resident;
    }

    template<Pairwise_interaction<Pt> pw_int, Pairwise_friction<Pt> pw_friction,
        bool SEEDED>
    int prepare()
    {
        static const int ctas_per_sm =
            resident_ctas(yb::sweep_cubes<Pt, pw_int, pw_friction, SEEDED>,
                yb::SWEEP_THREADS, yb::Sweep_config<yb::Layout<Pt>::lanes>::smem);
        if (split_sweep()) {
            prepare_list();
            prepare_interact<pw_int, pw_friction, SEEDED>();
        }
        return ctas_per_sm;
    }
    static int prepare_list()
    {
        static const int ctas_per_sm = resident_ctas(
            yb::list_cubes, yb::SWEEP_THREADS, yb::List_config::smem);
        return ctas_per_sm;
    }
    template<Pairwise_interaction<Pt> pw_int, Pairwise_friction<Pt> pw_friction,
        bool SEEDED>
    static int prepare_interact()
    {
        // no shared memory to speak of: all of the SM's array is L1
        static const int ctas_per_sm = [] {
            auto kernel = yb::interact_lists<Pt, pw_int, pw_friction, SEEDED>;
            YB_CUDA(cudaFuncSetAttribute(kernel,
                cudaFuncAttributePreferredSharedMemoryCarveout,
                cudaSharedmemCarveoutMaxL1));
            int resident = 0;
            YB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                &resident, kernel, yb::SWEEP_THREADS, 0));
            return resident < 1 ? 1 : resident;
        }();
        return ctas_per_sm;
    }

    template<Pairwise_interaction<Pt> pw_int, Pairwise_friction<Pt> pw_friction,
        bool SEEDED>
    static void load_kernels()
    {
        yb::load_kernel(yb::bin_cells<Pt>);
        yb::load_kernel(yb::scan_bins);
        yb::load_kernel(yb::place_ids);
        yb::load_kernel(yb::reorder_cells<Pt>);
        yb::load_kernel(yb::place_cells<Pt>);
        yb::load_kernel(yb::settle_cells<Pt>);
        yb::load_kernel(yb::list_cubes);
        yb::load_kernel(yb::interact_lists<Pt, pw_int, pw_friction, SEEDED>);
        yb::load_kernel(yb::sweep_cubes<Pt, pw_int, pw_friction, SEEDED>);
    }

    // Points with extra lanes run the sweep as two kernels (b200/pair_sweep.cuh,
    // list_cubes + interact_lists) while the cube-ordered planes and the
    // neighbour lists fit the L2: 1 M-cell tissues gain 2-7 %, but at 10 M cells
    // the partner gathers and the lists come from HBM and the fused kernel,
    // which stages positions with bulk copies, is 20 % faster
    // (profiles/r02_sweep_tuning.md). YALLA_B200_SPLIT_SWEEP=0/1 overrides.
    bool split_sweep() const
    {
        static const int forced = [] {
            const char* env = getenv("YALLA_B200_SPLIT_SWEEP");
            return env && env[0] ? atoi(env) : -1;
        }();
        if (forced >= 0) return forced != 0;
        return yb::Layout<Pt>::lanes > 4 && n_max <= 4 * 1000 * 1000;
    }

    // Cube ids -> bucket sort -> state in cube order (b200/grid_build.cuh).
    void build_index(cudaStream_t s, const int* d_n, const Pt* d_X,
        const float3* d_old_v, yb::Step_ctl* d_ctl, bool binned_by_predictor)
    {
        const int blocks = yb::stride_grid(n_max, 256, yb::sm_count());
        if (!binned_by_predictor)
            yb::bin_cells<Pt><<<blocks, 256, 0, s>>>(d_n, n_max, d_X, cube_size,
                box, sort.key, sort.arrival, sort.count, d_ctl);
        const int tiles = yb::ceil_div(box.n_cubes + 1, yb::SCAN_TILE);
        yb::scan_bins<<<tiles, yb::SCAN_THREADS, 0, s>>>(
            sort.count, sort.offset, tiles, sort.status, d_ctl);
        if (carry_state()) {
            yb::place_cells<Pt><<<blocks, 256, 0, s>>>(d_n, n_max, d_X, d_old_v,
                sort.key, sort.arrival, sort.offset, staged);
            yb::settle_cells<Pt><<<blocks, 256, 0, s>>>(d_n, n_max, staged,
                sort.offset, cube_size, box, pos4, aux, cube_sorted);
        } else {
            yb::place_ids<<<blocks, 256, 0, s>>>(
                d_n, n_max, sort.key, sort.arrival, sort.offset, sort.slot_id);
            yb::reorder_cells<Pt><<<blocks, 256, 0, s>>>(d_n, n_max, d_X,
                d_old_v, sort.key, sort.offset, sort.slot_id, pos4, aux,
                cube_sorted);
        }
    }

    // Decomposed tissues: the first pass of the build in two parts -- the owned
    // cells (part 0, while the halos are in flight) and the ghosts (part 1);
    // pwints() is then told that the cells are binned already.
    void bin_part(cudaStream_t s, const int* d_n, const Pt* d_X,
        yb::Step_ctl* d_ctl, int part)
    {
        const int blocks = yb::stride_grid(n_max, 256, yb::sm_count());
        yb::bin_cells_part<Pt><<<blocks, 256, 0, s>>>(d_n, n_max, d_X, cube_size,
            box, sort.key, sort.arrival, sort.count, d_ctl, part);
    }

    // Which of the two equivalent build tails to use (b200/grid_build.cuh):
    // scattering the state along with the ids wins once the state outgrows
    // the L2, provided a staged record is exactly one 32-byte sector (float3,
    // float4: relu_10M 5.18 -> 4.73 ms/step; with the 48-byte records of
    // Po_cell and the branching cell it loses 3 %). YALLA_B200_CARRY_STATE=0/1
    // overrides.
    bool carry_state() const
    {
        static const int forced = [] {
            const char* env = getenv("YALLA_B200_CARRY_STATE");
            return env && env[0] ? atoi(env) : -1;
        }();
        if (forced >= 0) return forced != 0;
        return yb::Layout<Pt>::aux_vec4 == 1 && n_max >= 4 * 1000 * 1000;
    }

    // pos4 of the last build: cube order with the original index in .w
    const float4* dd_cube_order() const { return pos4; }

    // The first stage's index, built before the generic forces are known (they only
    // seed dX); the next pwints() call then goes straight to the sweep.
    void index_ahead(cudaStream_t s, const int* d_n, const Pt* d_X,
        const float3* d_old_v, yb::Step_ctl* d_ctl)
    {
        build_index(s, d_n, d_X, d_old_v, d_ctl, false);
        indexed_ahead = true;
    }
    bool take_index_ahead()
    {
        const bool ahead = indexed_ahead;
        indexed_ahead = false;
        return ahead;
    }
    bool indexed_ahead = false;

    int persistent_ctas(int ctas_per_sm, int threads, int max_ctas) const
    {
        int ctas = yb::sm_count() * ctas_per_sm;
        const int chunks = yb::ceil_div(n_max > 0 ? n_max : 1, threads);
        if (ctas > chunks) ctas = chunks;
        if (ctas > max_ctas) ctas = max_ctas;
        return ctas;
    }

    template<Pairwise_interaction<Pt> pw_int, Pairwise_friction<Pt> pw_friction,
        bool SEEDED>
    void pwints(cudaStream_t s, const int* d_n, const Pt* d_X,
        const float3* d_old_v, Pt* d_dX, float* d_partials, int max_ctas,
        int stage, int drift_mode, int fix_point, yb::Step_ctl* d_ctl,
        bool binned_by_predictor, cudaEvent_t before_sweep = nullptr)
    {
        if (!take_index_ahead())
            build_index(s, d_n, d_X, d_old_v, d_ctl, binned_by_predictor);
        const int ctas = persistent_ctas(
            prepare<pw_int, pw_friction, SEEDED>(), yb::SWEEP_THREADS, max_ctas);
        if (before_sweep) YB_CUDA(cudaEventRecord(before_sweep, s));
        const bool split = split_sweep();
        if (split) {
            yb::list_cubes<<<persistent_ctas(prepare_list(), yb::SWEEP_THREADS,
                                 max_ctas),
                yb::SWEEP_THREADS, yb::List_config::smem, s>>>(d_n, n_max, pos4,
                cube_sorted, sort.offset, cube_size, box, nb, nb_count,
                nb_order, nb_stride, d_ctl, yb::LIST_MAX);
            yb::interact_lists<Pt, pw_int, pw_friction, SEEDED>
                <<<persistent_ctas(
                       prepare_interact<pw_int, pw_friction, SEEDED>(),
                       yb::SWEEP_THREADS, max_ctas),
                    yb::SWEEP_THREADS, 0, s>>>(d_n, n_max, pos4, aux, nb,
                    nb_count, nb_order, nb_stride, cube_size, d_dX, d_partials,
                    stage, drift_mode, fix_point, d_ctl);
        }
        // alone, or as the fallback for crowded tissues behind the pair above
        yb::sweep_cubes<Pt, pw_int, pw_friction, SEEDED>
            <<<ctas, yb::SWEEP_THREADS,
                yb::Sweep_config<yb::Layout<Pt>::lanes>::smem, s>>>(d_n, n_max, pos4,
                aux, cube_sorted, sort.offset, cube_size, box, d_dX, d_partials,
                stage, drift_mode, fix_point, d_ctl, split ? 1 : 0);
    }

    void predict(cudaStream_t s, int blocks, const int* d_n, float dt,
        const Pt* d_X, const Pt* d_dX, Pt* d_X1, yb::Step_ctl* d_ctl)
    {
        yb::predictor_step<Pt, true><<<blocks, 256, 0, s>>>(d_n, n_max, dt, d_X,
            d_dX, d_X1, d_ctl, cube_size, box, sort.key, sort.arrival,
            sort.count);
    }

    yb::Bucket_sort sort;
    float4* pos4;
    float4* aux;
    int* cube_sorted;
    float4* staged = nullptr;  // cube order, arrival order inside cubes (place_cells)
    int* nb = nullptr;         // neighbour lists of the split sweep, entry-major
    int* nb_count = nullptr;
    unsigned char* nb_order = nullptr;  // per chunk: cells by descending count
    int nb_stride = 0;
    const int n_max, grid_size, n_cubes;
    // the cubes the solver works on: the reference's cubic grid by default; a
    // domain of a decomposed tissue only its own box (dd_box_grid below)
    yb::Grid_box box;

public:
    // Extension (domain decomposition): restrict the grid to the box of
    // n[0] x n[1] x n[2] cubes starting at cube first[] of the global cubic
    // grid. Keeps the per-cube tables (and the scan over them) proportional
    // to the domain. The box must fit the tables allocated for grid_size^3.
    void dd_box_grid(const int first[3], const int n[3])
    {
        const long long cubes = 1LL * n[0] * n[1] * n[2];
        assert(n[0] >= 1 && n[1] >= 1 && n[2] >= 1 && cubes <= n_cubes);
        box.nx = n[0], box.ny = n[1], box.nz = n[2];
        box.x_half = grid_size / 2 - first[0];
        box.y_half = grid_size / 2 - first[1];
        box.z_half = grid_size / 2 - first[2];
        box.n_cubes = static_cast<int>(cubes);
        box.restricted = 1;
    }
    // z layers [first_layer, first_layer + n_layers) of the cubic grid
    void dd_slab_grid(int first_layer, int n_layers)
    {
        const int first[3] = {0, 0, first_layer};
        const int n[3] = {grid_size, grid_size, n_layers};
        dd_box_grid(first, n);
    }
};

template<typename Pt>
using Grid_solver = Heun_solver<Pt, Grid_computer>;

#include "b200/gabriel.cuh"
