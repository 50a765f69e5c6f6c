// Implementation of the C ABI declared in include/yalla_b200.h on top of the
// PUBLIC ya||a API only (Solution, take_step, Grid, Links, Property, ...).
//
// Built twice from this one source:
//   -I include                       -> yalla_b200/_lib/libyalla_b200.so
//   -I /root/reference/include       -> oracle/_ref/libyalla_ref.so
// The few places that use extensions of this repo's headers are guarded by
// #ifdef YALLA_B200 (defined by include/solvers.cuh).
#include <stdio.h>
#include <string.h>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "models.cuh"
#ifdef YALLA_B200
#include "b200/brick_links.cuh"
#endif

#include "yalla_b200.h"

namespace {

thread_local std::string last_error;

int fail(int code, const std::string& what)
{
    last_error = what;
    return code;
}

int check_cuda(const char* where)
{
    const cudaError_t status = cudaGetLastError();
    if (status == cudaSuccess) return YB_OK;
    return fail(YB_ECUDA, std::string(where) + ": " + cudaGetErrorString(status));
}

}  // namespace


// The interface every model implements.
struct yb_sim {
    virtual ~yb_sim() {}
    virtual int lanes() const = 0;
    virtual int n_max() const = 0;
    virtual int set_param(const std::string& name, double value)
    {
        return fail(YB_EINVAL, "unknown parameter " + name);
    }
    virtual int set_state(const float* h_X, int n, int reset_v) = 0;
    virtual int get_state(float* h_X, int capacity, int* n_out) = 0;
    virtual int get_velocities(float* h_v, int capacity) = 0;
    virtual int set_velocities(const float* h_v, int n) = 0;
    virtual int set_ints(const std::string& name, const int* values, int n)
    {
        return fail(YB_EINVAL, "model has no int property " + name);
    }
    virtual int get_ints(const std::string& name, int* values, int capacity)
    {
        return fail(YB_EINVAL, "model has no int property " + name);
    }
    virtual int seed_sphere(int, float, unsigned long long, int)
    {
        return fail(YB_ENOSYS, "seeded generators: float3 models of the product library");
    }
    virtual int set_links(const int* h_links, int n_links)
    {
        return fail(YB_EINVAL, "model has no links");
    }
    virtual int get_links(int* h_links, int capacity, int* n_out)
    {
        return fail(YB_EINVAL, "model has no links");
    }
    // One model step (asynchronous).
    virtual int step(float dt) = 0;
    // Host buffers in, n_steps steps, host buffers out; waits.
    virtual int step_host(const float* h_in, int n, float dt, int n_steps,
        float* h_out, int capacity, int* n_out)
    {
        int status = set_state(h_in, n, 0);
        if (status != YB_OK) return status;
        for (int k = 0; k < n_steps; k++) step(dt);
        return get_state(h_out, capacity, n_out);
    }
    // Extensions of the product library: a private stream per model, and a
    // host-buffer step that only enqueues (for pipelining independent batches).
    virtual int set_stream(void*)
    {
        return fail(YB_ENOSYS, "streams need the product library");
    }
    virtual int step_host_async(const float*, int, float, int, float*, int, int*)
    {
        return fail(YB_ENOSYS, "asynchronous steps need the product library");
    }
    virtual int host_drain()
    {
        return fail(YB_ENOSYS, "asynchronous steps need the product library");
    }
    // the stream the model's work is issued to
    virtual cudaStream_t work_stream() { return 0; }
    // Device address of the current cell count (for asynchronous snapshots).
    virtual const int* count_on_device() = 0;
    virtual int current_n() = 0;
    virtual int dd_load(int, const float*, const float*, int, const float*,
        const float*, int)
    {
        return fail(YB_ENOSYS, "domain decomposition: product Grid models only");
    }
    virtual int dd_forces(int, float*)
    {
        return fail(YB_ENOSYS, "domain decomposition: product Grid models only");
    }
    virtual int dd_update(int, float, const float*)
    {
        return fail(YB_ENOSYS, "domain decomposition: product Grid models only");
    }
    virtual int dd_read(int, float*, int)
    {
        return fail(YB_ENOSYS, "domain decomposition: product Grid models only");
    }
    virtual int slab_begin(float, float, float, int, int, int)
    {
        return fail(YB_ENOSYS, "domain decomposition: product Grid models only");
    }
    virtual int slab_set_owned(const float*, const float*, int)
    {
        return fail(YB_ENOSYS, "domain decomposition: product Grid models only");
    }
    virtual int slab_pack(int, float*, float*)
    {
        return fail(YB_ENOSYS, "domain decomposition: product Grid models only");
    }
    virtual int slab_unpack(int, const float*, const float*)
    {
        return fail(YB_ENOSYS, "domain decomposition: product Grid models only");
    }
    virtual int slab_update(int, float, const float*)
    {
        return fail(YB_ENOSYS, "domain decomposition: product Grid models only");
    }
    virtual int slab_counts(int*, int*, int*)
    {
        return fail(YB_ENOSYS, "domain decomposition: product Grid models only");
    }
    virtual int dom_begin(int, int, const float*, const float*, float, const int*,
        const int*, const int*, const int*)
    {
        return fail(YB_ENOSYS, "domain decomposition: product Grid models only");
    }
    virtual int dom_register_array(void*, int, int)
    {
        return fail(YB_ENOSYS, "domain decomposition: product Grid models only");
    }
    virtual int dom_exchange(void**, long long*, long long*)
    {
        return fail(YB_ENOSYS, "domain decomposition: product Grid models only");
    }
    virtual int dom_connect(int, void*, const long long*)
    {
        return fail(YB_ENOSYS, "domain decomposition: product Grid models only");
    }
    virtual int dom_connect_mailbox(int, void*)
    {
        return fail(YB_ENOSYS, "domain decomposition: product Grid models only");
    }
    virtual int dom_seed_lattice_ball(float, float, float, unsigned long long, int*)
    {
        return fail(YB_ENOSYS, "domain decomposition: product Grid models only");
    }
    virtual int dom_step(float, int)
    {
        return fail(YB_ENOSYS, "domain decomposition: product Grid models only");
    }
    virtual int dom_read_profile(float*)
    {
        return fail(YB_ENOSYS, "domain decomposition: product Grid models only");
    }
    virtual int profile_sweeps(int enable)
    {
        return fail(YB_ENOSYS, "sweep profiling needs the product library");
    }
    virtual int read_sweep_profile(float* total_ms, int* launches)
    {
        return fail(YB_ENOSYS, "sweep profiling needs the product library");
    }
};

namespace {

// State handling shared by all models: a Solution plus host-side glue.
template<typename Pt, template<typename> class Solver>
struct Sim_base : yb_sim {
    Solution<Pt, Solver> cells;
    int n_host;  // cell count as last set/read by the host

    template<typename... Args>
    Sim_base(int n_max, Args... args) : cells{n_max, args...}, n_host{0}
    {
        memset(cells.h_X, 0, sizeof(Pt) * static_cast<size_t>(n_max));
        *cells.h_n = 0;
        cells.copy_to_device();
    }
    int lanes() const override { return sizeof(Pt) / sizeof(float); }
    int n_max() const override { return cells.n_max; }

    int set_state(const float* h_X, int n, int reset_v) override
    {
        if (n < 0 || n > cells.n_max) return fail(YB_EINVAL, "n > n_max");
        memcpy(cells.h_X, h_X, sizeof(Pt) * static_cast<size_t>(n));
        *cells.h_n = n;
        cells.copy_to_device();
        if (reset_v) {
            cudaMemsetAsync(cells.d_old_v, 0,
                sizeof(float3) * static_cast<size_t>(cells.n_max), model_stream());
            cudaStreamSynchronize(model_stream());
        }
        n_host = n;
        return check_cuda("set_state");
    }
    int get_state(float* h_X, int capacity, int* n_out) override
    {
        cells.copy_to_host();
        const int n = *cells.h_n;
        if (n > capacity) return fail(YB_EINVAL, "capacity < n");
        memcpy(h_X, cells.h_X, sizeof(Pt) * static_cast<size_t>(n));
        if (n_out) *n_out = n;
        n_host = n;
        return check_cuda("get_state");
    }
    int get_velocities(float* h_v, int capacity) override
    {
        const int n = cells.get_d_n();
        if (n > capacity) return fail(YB_EINVAL, "capacity < n");
        cudaMemcpyAsync(h_v, cells.d_old_v, sizeof(float3) * static_cast<size_t>(n),
            cudaMemcpyDeviceToHost, model_stream());
        cudaStreamSynchronize(model_stream());
        return check_cuda("get_velocities");
    }
    int set_velocities(const float* h_v, int n) override
    {
        if (n < 0 || n > cells.n_max) return fail(YB_EINVAL, "n > n_max");
        cudaMemcpyAsync(cells.d_old_v, h_v, sizeof(float3) * static_cast<size_t>(n),
            cudaMemcpyHostToDevice, model_stream());
        cudaStreamSynchronize(model_stream());
        return check_cuda("set_velocities");
    }
    cudaStream_t work_stream() override { return model_stream(); }
    int current_n() override { return n_host = cells.get_d_n(); }
    const int* count_on_device() override { return cells.d_n; }
    // stream the model's own kernels are launched on (the solver's stream)
    cudaStream_t model_stream() const
    {
#ifdef YALLA_B200
        return cells.stream;
#else
        return 0;
#endif
    }
#ifdef YALLA_B200
    int set_stream(void* stream) override
    {
        cudaStreamSynchronize(cells.stream);
        cells.stream = static_cast<cudaStream_t>(stream);
        stream_changed();
        return YB_OK;
    }
    virtual void stream_changed() {}

    // ---- pipelined host-buffer steps ---------------------------------------
    // Independent batches through this one model instance: the upload of
    // batch k + 1 and the download of batch k - 1 run on two copy streams
    // while batch k integrates, through two staging slots in device memory.
    // Events order slot reuse; nothing here waits on the host except when a
    // slot's previous download has not been drained yet.
    struct Host_pipeline {
        cudaStream_t up = nullptr, down = nullptr;
        Pt* d_in[2] = {nullptr, nullptr};
        Pt* d_out[2] = {nullptr, nullptr};
        int* d_counts = nullptr;  // [0..1] n per slot in, [2..3] n per slot out
        int* h_n_in = nullptr;    // pinned, one per slot
        cudaEvent_t uploaded[2], consumed[2], computed[2], downloaded[2];
        bool used[2] = {false, false};
        int next = 0;
    } pipe;

    void open_pipeline()
    {
        if (pipe.up) return;
        const size_t bytes = sizeof(Pt) * static_cast<size_t>(cells.n_max);
        cudaStreamCreateWithFlags(&pipe.up, cudaStreamNonBlocking);
        cudaStreamCreateWithFlags(&pipe.down, cudaStreamNonBlocking);
        for (int k = 0; k < 2; k++) {
            cudaMalloc(&pipe.d_in[k], bytes);
            cudaMalloc(&pipe.d_out[k], bytes);
            cudaEventCreateWithFlags(&pipe.uploaded[k], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&pipe.consumed[k], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&pipe.computed[k], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&pipe.downloaded[k], cudaEventDisableTiming);
        }
        cudaMalloc(&pipe.d_counts, 4 * sizeof(int));
        cudaMallocHost(&pipe.h_n_in, 2 * sizeof(int));
    }
    void close_pipeline()
    {
        if (!pipe.up) return;
        cudaStreamSynchronize(pipe.up);
        cudaStreamSynchronize(pipe.down);
        for (int k = 0; k < 2; k++) {
            cudaFree(pipe.d_in[k]);
            cudaFree(pipe.d_out[k]);
            cudaEventDestroy(pipe.uploaded[k]);
            cudaEventDestroy(pipe.consumed[k]);
            cudaEventDestroy(pipe.computed[k]);
            cudaEventDestroy(pipe.downloaded[k]);
        }
        cudaFree(pipe.d_counts);
        cudaFreeHost(pipe.h_n_in);
        cudaStreamDestroy(pipe.up);
        cudaStreamDestroy(pipe.down);
        pipe.up = pipe.down = nullptr;
    }
    ~Sim_base() override
    {
        close_recording();
        close_pipeline();
    }

    // Enqueue upload, steps and the download of `out_cells` cells plus the
    // count (into pinned *h_n_out); yb_sim_host_drain waits for all of it.
    int step_host_async(const float* h_in, int n, float dt, int n_steps,
        float* h_out, int out_cells, int* h_n_out) override
    {
        if (n < 0 || n > cells.n_max || out_cells > cells.n_max)
            return fail(YB_EINVAL, "n > n_max");
        open_pipeline();
        const int k = pipe.next;
        pipe.next ^= 1;
        const cudaStream_t work = cells.stream;
        if (pipe.used[k]) {
            // the slot's input was consumed and its output drained?
            cudaStreamWaitEvent(pipe.up, pipe.consumed[k], 0);
            cudaStreamWaitEvent(work, pipe.downloaded[k], 0);
            cudaEventSynchronize(pipe.consumed[k]);  // h_n_in[k] is free again
        }
        pipe.h_n_in[k] = n;
        cudaMemcpyAsync(pipe.d_in[k], h_in, sizeof(Pt) * size_t(n),
            cudaMemcpyHostToDevice, pipe.up);
        cudaMemcpyAsync(pipe.d_counts + k, pipe.h_n_in + k, sizeof(int),
            cudaMemcpyHostToDevice, pipe.up);
        cudaEventRecord(pipe.uploaded[k], pipe.up);

        cudaStreamWaitEvent(work, pipe.uploaded[k], 0);
        cudaMemcpyAsync(cells.d_X, pipe.d_in[k], sizeof(Pt) * size_t(n),
            cudaMemcpyDeviceToDevice, work);
        cudaMemcpyAsync(cells.d_n, pipe.d_counts + k, sizeof(int),
            cudaMemcpyDeviceToDevice, work);
        cudaEventRecord(pipe.consumed[k], work);
        for (int q = 0; q < n_steps; q++) this->step(dt);
        cudaMemcpyAsync(pipe.d_out[k], cells.d_X, sizeof(Pt) * size_t(out_cells),
            cudaMemcpyDeviceToDevice, work);
        cudaMemcpyAsync(pipe.d_counts + 2 + k, cells.d_n, sizeof(int),
            cudaMemcpyDeviceToDevice, work);
        cudaEventRecord(pipe.computed[k], work);

        cudaStreamWaitEvent(pipe.down, pipe.computed[k], 0);
        cudaMemcpyAsync(h_out, pipe.d_out[k], sizeof(Pt) * size_t(out_cells),
            cudaMemcpyDeviceToHost, pipe.down);
        cudaMemcpyAsync(h_n_out, pipe.d_counts + 2 + k, sizeof(int),
            cudaMemcpyDeviceToHost, pipe.down);
        cudaEventRecord(pipe.downloaded[k], pipe.down);
        pipe.used[k] = true;
        return check_cuda("yb_sim_step_host_async");
    }
    int host_drain() override
    {
        if (pipe.up) {
            cudaStreamSynchronize(pipe.up);
            cudaStreamSynchronize(cells.stream);
            cudaStreamSynchronize(pipe.down);
        }
        return check_cuda("yb_sim_host_drain");
    }
    // n live cells straight between the caller's buffers and the device
    // (Solution::copy_to_device/host move n_max cells through h_X)
    int step_host(const float* h_in, int n, float dt, int n_steps, float* h_out,
        int capacity, int* n_out) override
    {
        if (n < 0 || n > cells.n_max) return fail(YB_EINVAL, "n > n_max");
        cells.upload(reinterpret_cast<const Pt*>(h_in), n);
        for (int k = 0; k < n_steps; k++) this->step(dt);
        const int n_now = cells.get_d_n();
        if (n_now > capacity) return fail(YB_EINVAL, "capacity < n");
        n_host = cells.download(reinterpret_cast<Pt*>(h_out), capacity);
        if (n_out) *n_out = n_host;
        return check_cuda("yb_sim_step_host");
    }
    bool profiling_sweeps = false;
    int profile_sweeps(int enable) override
    {
        cells.profile_sweeps(enable != 0);
        profiling_sweeps = enable != 0;
        return YB_OK;
    }
    int read_sweep_profile(float* total_ms, int* launches) override
    {
        cells.read_sweep_profile(total_ms, launches);
        return YB_OK;
    }

    // ---- a whole model iteration as one CUDA graph ---------------------------
    // Models whose iteration is more than take_step (division, rewiring of
    // links, ...) record it once per dt on a side stream -- the solver issues
    // its stages into the recording (capturable generic forces) -- and replay
    // it with one graph launch. The first iteration runs directly, so that
    // kernel attributes and scratch are set up outside any recording.
    virtual void enqueue_iteration(float dt) {}
    cudaGraphExec_t iteration = nullptr;
    float iteration_dt = 0.f;
    bool warmed_up = false;
    cudaStream_t recording_stream = nullptr;

    void drop_iteration()
    {
        if (iteration == nullptr) return;
        cudaStreamSynchronize(model_stream());
        cudaGraphExecDestroy(iteration);
        iteration = nullptr;
    }
    void close_recording()
    {
        drop_iteration();
        if (recording_stream) cudaStreamDestroy(recording_stream);
        recording_stream = nullptr;
    }
    int replay_iteration(float dt)
    {
        if (!yb::graphs_enabled() || profiling_sweeps || !warmed_up) {
            warmed_up = true;
            enqueue_iteration(dt);
            return 0;
        }
        if (iteration == nullptr || iteration_dt != dt) {
            drop_iteration();
            if (recording_stream == nullptr)
                cudaStreamCreateWithFlags(&recording_stream, cudaStreamNonBlocking);
            const cudaStream_t users = cells.stream;
            cudaGraph_t graph;
            cudaStreamBeginCapture(recording_stream, cudaStreamCaptureModeRelaxed);
            cells.stream = recording_stream;
            enqueue_iteration(dt);
            cells.stream = users;
            cudaStreamEndCapture(recording_stream, &graph);
            cudaGraphInstantiate(&iteration, graph, 0);
            cudaGraphDestroy(graph);
            iteration_dt = dt;
        }
        cudaGraphLaunch(iteration, model_stream());
        return 0;
    }

    // ---- domain decomposition (device pointers) ---------------------------
    int dd_n_owned = 0;
    int dd_load(int stage, const float* X_owned, const float* v_owned,
        int n_owned, const float* X_ghost, const float* v_ghost,
        int n_ghost) override
    {
        if (n_owned < 0 || n_ghost < 0 || n_owned + n_ghost > cells.n_max)
            return fail(YB_EINVAL, "n_owned + n_ghost > n_max");
        Pt* X = cells.dd_positions(stage);
        float3* v = cells.dd_velocities();
        const cudaMemcpyKind d2d = cudaMemcpyDeviceToDevice;
        if (stage == 0) {
            cudaMemcpyAsync(
                X, X_owned, sizeof(Pt) * size_t(n_owned), d2d, cells.stream);
            cudaMemcpyAsync(
                v, v_owned, sizeof(float3) * size_t(n_owned), d2d, cells.stream);
            dd_n_owned = n_owned;
        } else if (n_owned != dd_n_owned) {
            return fail(YB_EINVAL, "stage 1 must keep the owned cells of stage 0");
        }
        if (n_ghost > 0) {
            cudaMemcpyAsync(X + n_owned, X_ghost, sizeof(Pt) * size_t(n_ghost),
                d2d, cells.stream);
            cudaMemcpyAsync(v + n_owned, v_ghost,
                sizeof(float3) * size_t(n_ghost), d2d, cells.stream);
        }
        cells.dd_set_counts(n_owned, n_owned + n_ghost);
        return check_cuda("yb_dd_load");
    }
    int dd_update(int stage, float dt, const float* mean3) override
    {
        cells.dd_update(stage, dt, mean3);
        return check_cuda("yb_dd_update");
    }
    int dd_read(int which, float* out, int n) override
    {
        if (n > cells.n_max) return fail(YB_EINVAL, "n > n_max");
        const cudaMemcpyKind d2d = cudaMemcpyDeviceToDevice;
        if (which == 0 || which == 1)
            cudaMemcpyAsync(out, cells.dd_positions(which),
                sizeof(Pt) * size_t(n), d2d, cells.stream);
        else
            cudaMemcpyAsync(out, cells.dd_velocities(),
                sizeof(float3) * size_t(n), d2d, cells.stream);
        return check_cuda("yb_dd_read");
    }
    int slab_begin(float z_lo, float z_hi, float halo, int capacity,
        int first_layer, int n_layers) override
    {
        return slab_begin_impl(z_lo, z_hi, halo, capacity, first_layer, n_layers,
            std::is_same<Solver<Pt>, Grid_solver<Pt>>{});
    }
    int slab_begin_impl(float z_lo, float z_hi, float halo, int capacity,
        int first_layer, int n_layers, std::true_type)
    {
        if (capacity <= 0) return fail(YB_EINVAL, "capacity must be positive");
        if (n_layers > 0) cells.dd_slab_grid(first_layer, n_layers);
        cells.slab_begin(z_lo, z_hi, halo, capacity);
        return check_cuda("yb_slab_begin");
    }
    int slab_begin_impl(float, float, float, int, int, int, std::false_type)
    {
        return fail(YB_ENOSYS, "domain decomposition needs a Grid model");
    }
    int slab_set_owned(const float* X, const float* v, int n_owned) override
    {
        if (n_owned < 0 || n_owned > cells.n_max)
            return fail(YB_EINVAL, "n_owned > n_max");
        cells.slab_set_owned(reinterpret_cast<const Pt*>(X),
            reinterpret_cast<const float3*>(v), n_owned);
        return check_cuda("yb_slab_set_owned");
    }
    int slab_pack(int what, float* send_lo, float* send_hi) override
    {
        cells.slab_pack(what, send_lo, send_hi);
        return check_cuda("yb_slab_pack");
    }
    int slab_unpack(int what, const float* recv_lo, const float* recv_hi) override
    {
        cells.slab_unpack(what, recv_lo, recv_hi);
        return check_cuda("yb_slab_unpack");
    }
    int slab_update(int stage, float dt, const float* sums4) override
    {
        cells.slab_update(stage, dt, sums4);
        return check_cuda("yb_slab_update");
    }
    int slab_counts(int* n_owned, int* n_total, int* problems) override
    {
        cells.slab_counts(n_owned, n_total, problems);
        return check_cuda("yb_slab_counts");
    }
    // ---- brick decomposition over peer memory (b200/domain.cuh) ---------------
    int dom_begin(int rank, int world, const float* lo3, const float* hi3,
        float halo, const int* peer_ranks27, const int* capacity27,
        const int* box_first3, const int* box_n3) override
    {
        return dom_begin_impl(rank, world, lo3, hi3, halo, peer_ranks27,
            capacity27, box_first3, box_n3,
            std::is_same<Solver<Pt>, Grid_solver<Pt>>{});
    }
    int dom_begin_impl(int rank, int world, const float* lo3, const float* hi3,
        float halo, const int* peer_ranks27, const int* capacity27,
        const int* box_first3, const int* box_n3, std::true_type)
    {
        if (world < 1 || world > yb::DD_MAX_RANKS || rank < 0 || rank >= world)
            return fail(YB_EINVAL, "bad rank / world");
        for (int d = 0; d < 27; d++)
            if (d != 13 && peer_ranks27[d] >= 0 && capacity27[d] <= 0)
                return fail(YB_EINVAL, "a neighbour needs a positive capacity");
        if (box_n3 != nullptr && box_n3[0] > 0)
            cells.dd_box_grid(box_first3, box_n3);
        cells.dom_begin(rank, world, lo3, hi3, halo, peer_ranks27, capacity27);
        return check_cuda("yb_dom_begin");
    }
    int dom_begin_impl(int, int, const float*, const float*, float, const int*,
        const int*, const int*, const int*, std::false_type)
    {
        return fail(YB_ENOSYS, "domain decomposition needs a Grid model");
    }
    int dom_register_array(void* d_array, int bytes_per_cell, int ghosts_too) override
    {
        if (d_array == nullptr || bytes_per_cell <= 0 || bytes_per_cell % 4 != 0)
            return fail(YB_EINVAL, "bytes_per_cell must be a positive multiple of 4");
        if (!cells.dom_register_array(d_array, bytes_per_cell, ghosts_too != 0))
            return fail(YB_EINVAL,
                "register arrays before yb_dom_begin, at most " +
                    std::to_string(yb::DD_MAX_EXTRAS));
        return YB_OK;
    }
    int dom_exchange(void** base, long long* bytes, long long* offsets) override
    {
        const yb::Domain_link& dom = cells.dom;
        if (!dom.active) return fail(YB_EINVAL, "yb_dom_begin first");
        *base = dom.base;
        *bytes = static_cast<long long>(dom.bytes);
        constexpr int R = yb::DD_ROUNDS;
        static_assert(2 * R == YB_DOM_OFFSETS, "yalla_b200.h: YB_DOM_OFFSETS");
        for (int q = 0; q < 27 * 2 * R; q++) offsets[q] = -1;
        for (int p = 0; p < dom.region.n_peers; p++)
            for (int q = 0; q < R; q++) {
                offsets[dom.peer_direction[p] * 2 * R + q] =
                    static_cast<long long>(dom.inbox_offset[p][q]);
                offsets[dom.peer_direction[p] * 2 * R + R + q] =
                    static_cast<long long>(dom.flag_offset[p][q]);
            }
        return YB_OK;
    }
    int dom_connect(int direction, void* peer_base,
        const long long* peer_offsets) override
    {
        if (!cells.dom.active) return fail(YB_EINVAL, "yb_dom_begin first");
        for (int q = 0; q < 2 * yb::DD_ROUNDS; q++)
            if (peer_offsets[q] < 0)
                return fail(YB_EINVAL, "the peer has no inbox for this direction");
        if (!cells.dom.connect(direction, peer_base, peer_offsets))
            return fail(YB_EINVAL, "no neighbour in this direction");
        return YB_OK;
    }
    int dom_connect_mailbox(int rank, void* peer_base) override
    {
        if (!cells.dom.active || rank < 0 || rank >= cells.dom.world)
            return fail(YB_EINVAL, "bad rank");
        cells.dom.connect_mailbox(rank, peer_base);
        return YB_OK;
    }
    int dom_seed_lattice_ball(float radius, float dist_to_nb, float jitter,
        unsigned long long seed, int* n_out) override
    {
        if (!cells.dom.active) return fail(YB_EINVAL, "yb_dom_begin first");
        const int wanted =
            cells.dom_seed_lattice_ball(radius, dist_to_nb, jitter, seed);
        if (n_out) *n_out = wanted;
        if (wanted > cells.n_max)
            return fail(YB_EINVAL, "n_max is too small for this brick: it holds " +
                                       std::to_string(wanted) + " cells");
        return check_cuda("yb_dom_seed_lattice_ball");
    }
    int dom_read_profile(float* ms6) override
    {
        cells.read_dom_profile(ms6);
        return check_cuda("yb_dom_read_profile");
    }
    template<Pairwise_interaction<Pt> force, Pairwise_friction<Pt> friction>
    int dom_step_with(float dt, int n_steps)
    {
        if (!cells.dom.active || !cells.dom.connected())
            return fail(YB_EINVAL, "the domain is not connected to its neighbours");
        for (int k = 0; k < n_steps; k++)
            cells.template dom_step<force, friction>(dt);
        return check_cuda("yb_dom_step");
    }

    // the sweep needs the model's functor: models that support decomposition
    // call this from their dd_forces override
    template<Pairwise_interaction<Pt> force, Pairwise_friction<Pt> friction>
    int dd_forces_with(int stage, float* sums4)
    {
        cells.template dd_forces<force, friction>(stage);
        cudaMemcpyAsync(sums4, cells.dd_drift_sum(stage), 4 * sizeof(float),
            cudaMemcpyDeviceToDevice, cells.stream);
        return check_cuda("yb_dd_forces");
    }
#endif

    int set_fix(const std::string& name, double value)
    {
        if (name == "fix_point") {
            cells.set_fixed(static_cast<int>(value));
            return YB_OK;
        }
        if (name == "fix_point_xy") {
            cells.set_fixed_xy(static_cast<int>(value));
            return YB_OK;
        }
        if (name == "fix_com") {
            cells.set_fixed();
            return YB_OK;
        }
        return 1;  // not a fixing parameter
    }
};

int set_spring_length(double value)
{
    const float length = static_cast<float>(value);
    cudaMemcpyToSymbol(models::d_spring_length, &length, sizeof(float));
    return check_cuda("spring_length");
}

// ---- float3 models with a stateless functor --------------------------------
template<template<typename> class Solver, Pairwise_interaction<float3> force>
struct Spring_sim : Sim_base<float3, Solver> {
    using Base = Sim_base<float3, Solver>;
    template<typename... Args>
    Spring_sim(int n_max, Args... args) : Base{n_max, args...}
    {
        set_spring_length(0.5);
#ifdef YALLA_B200
        // the functors of these models are pure, so the Tile solver may share
        // one cell between several lanes (solvers.cuh, Tile_computer);
        // set_param("split_pairs", 0) switches back to one thread per cell
        if constexpr (std::is_same<Solver<float3>, Tile_solver<float3>>::value)
            this->cells.split_pairs = true;
#endif
    }
    int set_param(const std::string& name, double value) override
    {
        if (name == "spring_length") return set_spring_length(value);
#ifdef YALLA_B200
        if constexpr (std::is_same<Solver<float3>, Tile_solver<float3>>::value) {
            if (name == "split_pairs") {
                this->cells.split_pairs = value != 0;
                return YB_OK;
            }
        }
#endif
        const int fixed = Base::set_fix(name, value);
        return fixed <= 0 ? fixed : yb_sim::set_param(name, value);
    }
    int step(float dt) override
    {
        this->cells.template take_step<force>(dt);
        return 0;
    }
#ifdef YALLA_B200
    int seed_sphere(int n, float dist_to_nb, unsigned long long seed,
        int relax_steps) override
    {
        if (n <= 0 || n > this->cells.n_max) return fail(YB_EINVAL, "n > n_max");
        *this->cells.h_n = n;
        cudaMemsetAsync(this->cells.d_old_v, 0,
            sizeof(float3) * static_cast<size_t>(this->cells.n_max),
            this->cells.stream);
        if (relax_steps == 0)
            seeded_sphere(dist_to_nb, this->cells, seed);
        else
            relaxed_seeded_sphere(dist_to_nb, this->cells, seed, 0, relax_steps);
        this->n_host = n;
        return check_cuda("yb_sim_seed_sphere");
    }
    int dd_forces(int stage, float* sums4) override
    {
        // only the grid solver knows ghosts; the Tile solver is replicas-only
        if constexpr (std::is_same<Solver<float3>, Grid_solver<float3>>::value)
            return Base::template dd_forces_with<force,
                friction_w_neighbour<float3>>(stage, sums4);
        else
            return fail(YB_ENOSYS, "domain decomposition needs a Grid model");
    }
    int dom_step(float dt, int n_steps) override
    {
        if constexpr (std::is_same<Solver<float3>, Grid_solver<float3>>::value)
            return Base::template dom_step_with<force,
                friction_w_neighbour<float3>>(dt, n_steps);
        else
            return fail(YB_ENOSYS, "domain decomposition needs a Grid model");
    }
#endif
};

// ---- float3 + Links: relu_force with link_forces as generic force -----------
struct Protrusion_sim : Sim_base<float3, Grid_solver> {
    using Base = Sim_base<float3, Grid_solver>;
    Links links;
    Protrusion_sim(int n_max, int grid_size, float cube_size)
        : Base{n_max, grid_size, cube_size}, links{4 * n_max}
    {
        links.set_d_n(0);
#ifdef YALLA_B200
        // link_forces only enqueues kernels on the step's stream: the whole
        // step replays from one CUDA graph; the links only change through
        // set_links (copy_to_device), so their per-cell index is kept.
        cells.capture_generic_forces = true;
        links.cache_topology = true;
#endif
    }
    int set_param(const std::string& name, double value) override
    {
        if (name == "link_strength") {
            links.strength = static_cast<float>(value);
            return YB_OK;
        }
        const int fixed = Base::set_fix(name, value);
        return fixed <= 0 ? fixed : yb_sim::set_param(name, value);
    }
    int set_links(const int* h_links, int n_links) override
    {
        if (n_links > links.n_max) return fail(YB_EINVAL, "too many links");
        memcpy(links.h_link, h_links, sizeof(Link) * static_cast<size_t>(n_links));
        *links.h_n = n_links;
        links.copy_to_device();
        return check_cuda("set_links");
    }
    int get_links(int* h_links, int capacity, int* n_out) override
    {
        cudaStreamSynchronize(model_stream());
        links.copy_to_host();
        if (*links.h_n > capacity) return fail(YB_EINVAL, "capacity < n_links");
        memcpy(h_links, links.h_link, sizeof(Link) * size_t(*links.h_n));
        if (n_out) *n_out = *links.h_n;
        return check_cuda("get_links");
    }
    int step(float dt) override
    {
        auto pull = [this](const int n, const float3* __restrict__ d_X,
                        float3* d_dX) { link_forces(links, d_X, d_dX); };
        cells.take_step<relu_force<float3>>(dt, pull);
        return 0;
    }
};

// ---- Po_cell epithelium -------------------------------------------------------------
struct Epithelium_sim : Sim_base<Po_cell, Grid_solver> {
    using Base = Sim_base<Po_cell, Grid_solver>;
    Epithelium_sim(int n_max, int grid_size, float cube_size)
        : Base{n_max, grid_size, cube_size}
    {}
    int set_param(const std::string& name, double value) override
    {
        const int fixed = Base::set_fix(name, value);
        return fixed <= 0 ? fixed : yb_sim::set_param(name, value);
    }
    int step(float dt) override
    {
        cells.take_step<models::layer_force, friction_on_background>(dt);
        return 0;
    }
#ifdef YALLA_B200
    int dd_forces(int stage, float* sums4) override
    {
        return dd_forces_with<models::layer_force, friction_on_background<Po_cell>>(
            stage, sums4);
    }
    int dom_step(float dt, int n_steps) override
    {
        return dom_step_with<models::layer_force, friction_on_background<Po_cell>>(
            dt, n_steps);
    }
#endif
};

// Models with a cell type and neighbour counters as Property arrays.
template<typename Pt>
struct Typed_sim : Sim_base<Pt, Grid_solver> {
    using Base = Sim_base<Pt, Grid_solver>;
    Property<models::Cell_types> type;
    Property<int> n_mes_nbs, n_epi_nbs;

    Typed_sim(int n_max, int grid_size, float cube_size)
        : Base{n_max, grid_size, cube_size}, type{n_max},
          n_mes_nbs{n_max, "n_mes_nbs"}, n_epi_nbs{n_max, "n_epi_nbs"}
    {
        for (int i = 0; i < n_max; i++) {
            type.h_prop[i] = models::mesenchyme;
            n_mes_nbs.h_prop[i] = 0;
            n_epi_nbs.h_prop[i] = 0;
        }
        type.copy_to_device();
        n_mes_nbs.copy_to_device();
        n_epi_nbs.copy_to_device();
    }
    // The functors find the property arrays through __device__ pointers
    // (models.cuh), which are global to the process: one Typed_sim is bound
    // at a time. Binding another one first waits for the stream of the model
    // that was bound before, so the kernels of two such models never overlap
    // with the wrong pointers in place; a model that stays bound pays nothing.
    struct Binding {
        const void* owner = nullptr;
        cudaStream_t stream = 0;
    };
    static Binding& binding()
    {
        static Binding current;
        return current;
    }
    void bind()
    {
        Binding& current = binding();
        if (current.owner == this) return;
        if (current.owner != nullptr) cudaStreamSynchronize(current.stream);
        const cudaStream_t s = this->model_stream();
        cudaMemcpyToSymbolAsync(models::d_type, &type.d_prop, sizeof(type.d_prop),
            0, cudaMemcpyHostToDevice, s);
        cudaMemcpyToSymbolAsync(models::d_mes_nbs, &n_mes_nbs.d_prop,
            sizeof(n_mes_nbs.d_prop), 0, cudaMemcpyHostToDevice, s);
        cudaMemcpyToSymbolAsync(models::d_epi_nbs, &n_epi_nbs.d_prop,
            sizeof(n_epi_nbs.d_prop), 0, cudaMemcpyHostToDevice, s);
        current.owner = this;
        current.stream = s;
    }
    void unbind()
    {
        Binding& current = binding();
        if (current.owner != this) return;
        cudaStreamSynchronize(current.stream);
        current.owner = nullptr;
    }
    ~Typed_sim() override { unbind(); }
#ifdef YALLA_B200
    void stream_changed() override { unbind(); }
    // Decomposed: the type travels with the cells AND their ghost copies (the
    // functors read the partner's type), the counters with the cells only.
    virtual void register_model_arrays()
    {
        auto& cells = this->cells;
        cells.dom_register_array(type.d_prop, sizeof(models::Cell_types), true);
        cells.dom_register_array(n_mes_nbs.d_prop, sizeof(int), false);
        cells.dom_register_array(n_epi_nbs.d_prop, sizeof(int), false);
    }
    int dom_begin(int rank, int world, const float* lo3, const float* hi3,
        float halo, const int* peer_ranks27, const int* capacity27,
        const int* box_first3, const int* box_n3) override
    {
        static_assert(sizeof(models::Cell_types) == 4, "4-byte cell types");
        if (this->cells.dom.extras.count == 0) register_model_arrays();
        return Base::dom_begin(rank, world, lo3, hi3, halo, peer_ranks27,
            capacity27, box_first3, box_n3);
    }
#endif
    int set_ints(const std::string& name, const int* values, int n) override
    {
        if (n > this->cells.n_max) return fail(YB_EINVAL, "n > n_max");
        if (name == "type") {
            cudaStreamSynchronize(this->model_stream());
            for (int i = 0; i < n; i++)
                type.h_prop[i] = static_cast<models::Cell_types>(values[i]);
            type.copy_to_device();
            return check_cuda("set type");
        }
        return yb_sim::set_ints(name, values, n);
    }
    int get_ints(const std::string& name, int* values, int capacity) override
    {
        const int n = this->cells.get_d_n();  // waits for the model's stream
        if (n > capacity) return fail(YB_EINVAL, "capacity < n");
        if (name == "type") {
            type.copy_to_host();
            for (int i = 0; i < n; i++) values[i] = type.h_prop[i];
        } else if (name == "mes_nbs") {
            n_mes_nbs.copy_to_host();
            memcpy(values, n_mes_nbs.h_prop, sizeof(int) * size_t(n));
        } else if (name == "epi_nbs") {
            n_epi_nbs.copy_to_host();
            memcpy(values, n_epi_nbs.h_prop, sizeof(int) * size_t(n));
        } else {
            return yb_sim::get_ints(name, values, capacity);
        }
        return check_cuda("get_ints");
    }
    // reset_nbs of passive_growth.cu:107-113 / branching.cu:188-193, using the
    // n handed to the callback instead of reading d_n back.
    void reset_counters(int n)
    {
        // inside the generic-forces callback: the stream of the step, which is
        // the solver's recording stream while a step is being captured
        cudaStream_t s = this->model_stream();
#ifdef YALLA_B200
        if (const yb::Stage_context* stage = yb::current_stage()) s = stage->stream;
#endif
        cudaMemsetAsync(n_mes_nbs.d_prop, 0, sizeof(int) * size_t(n), s);
        cudaMemsetAsync(n_epi_nbs.d_prop, 0, sizeof(int) * size_t(n), s);
    }
};

// ---- Po_cell growth ------------------------------------------------------------------
struct Growth_sim : Typed_sim<Po_cell> {
    curandState* d_state = nullptr;
    int* d_n_at_launch = nullptr;
    float prolif_rate = 0.006f, mean_dist = 0.75f;
    int seed = 2;
    bool seeded = false;
#ifdef YALLA_B200
    bool reproducible = false;
    std::unique_ptr<Cell_division<Po_cell>> division;
#endif

    Growth_sim(int n_max, int grid_size, float cube_size)
        : Typed_sim<Po_cell>{n_max, grid_size, cube_size}
    {
        cudaMalloc(&d_state, sizeof(curandState) * size_t(n_max));
        cudaMalloc(&d_n_at_launch, sizeof(int));
    }
    ~Growth_sim() override
    {
        cudaFree(d_n_at_launch);
        cudaFree(d_state);
    }
    int set_param(const std::string& name, double value) override
    {
#ifdef YALLA_B200
        drop_iteration();  // parameters are baked into the recorded launches
#endif
        if (name == "prolif_rate") {
            prolif_rate = static_cast<float>(value);
        } else if (name == "mean_dist") {
            mean_dist = static_cast<float>(value);
        } else if (name == "seed") {
            seed = static_cast<int>(value);
            seeded = false;
#ifdef YALLA_B200
            division.reset();
        } else if (name == "reproducible_division") {
            // Cell_division instead of the example's proliferate kernel
            reproducible = value != 0;
#endif
        } else {
            const int fixed = Base::set_fix(name, value);
            return fixed <= 0 ? fixed : yb_sim::set_param(name, value);
        }
        return YB_OK;
    }
    // What examples/passive_growth.cu does per iteration on the device.
    void enqueue_iteration(float dt)
#ifdef YALLA_B200
        override
#endif
    {
        const int n_max = cells.n_max;
        auto reset_nbs = [this](const int n, const Po_cell* __restrict__ d_X,
                             Po_cell* d_dX) { reset_counters(n); };
        cells.take_step<models::relu_w_epithelium>(dt, reset_nbs);
        if (prolif_rate > 0) {
            // sized for the capacity; the kernel reads the live count itself
            models::snapshot_count<<<1, 1, 0, model_stream()>>>(
                cells.d_n, d_n_at_launch);
            models::proliferate<<<(n_max + 128 - 1) / 128, 128, 0,
                model_stream()>>>(prolif_rate,
                mean_dist, n_max, d_state, cells.d_X, cells.d_old_v, cells.d_n,
                d_n_at_launch);
        }
    }

#ifdef YALLA_B200
    // Decomposed (one brick of the tissue per instance): the curand states
    // travel with the cells too; every brick seeds its own streams.
    void register_model_arrays() override
    {
        Typed_sim<Po_cell>::register_model_arrays();
        static_assert(sizeof(curandState) % 4 == 0, "curandState in words");
        cells.dom_register_array(d_state, sizeof(curandState), false);
    }
    int dom_step(float dt, int n_steps) override
    {
        if (!cells.dom.active || !cells.dom.connected())
            return fail(YB_EINVAL, "the domain is not connected to its neighbours");
        const int n_max = cells.n_max;
        if (!seeded) {
            setup_rand_states<<<(n_max + 128 - 1) / 128, 128, 0, model_stream()>>>(
                n_max, seed + 7919 * cells.dom.rank, d_state);
            seeded = true;
        }
        bind();
        auto reset_nbs = [this](const int n, const Po_cell* __restrict__ d_X,
                             Po_cell* d_dX) { reset_counters(n); };
        for (int k = 0; k < n_steps; k++) {
            cells.dom_step<models::relu_w_epithelium, friction_w_neighbour<Po_cell>>(
                dt, reset_nbs);
            if (prolif_rate > 0) {
                models::snapshot_count<<<1, 1, 0, model_stream()>>>(
                    cells.d_n, d_n_at_launch);
                models::proliferate<<<(n_max + 128 - 1) / 128, 128, 0,
                    model_stream()>>>(prolif_rate, mean_dist, n_max, d_state,
                    cells.d_X, cells.d_old_v, cells.d_n, d_n_at_launch);
            }
        }
        return check_cuda("yb_dom_step");
    }
#endif

    int step(float dt) override
    {
        const int n_max = cells.n_max;
        if (!seeded) {
            setup_rand_states<<<(n_max + 128 - 1) / 128, 128, 0, model_stream()>>>(
                n_max, seed, d_state);
            seeded = true;
        }
        bind();
#ifdef YALLA_B200
        if (prolif_rate > 0 && reproducible) {
            auto reset_nbs = [this](const int n, const Po_cell* __restrict__ d_X,
                                 Po_cell* d_dX) { reset_counters(n); };
            cells.take_step<models::relu_w_epithelium>(dt, reset_nbs);
            if (!division)
                division.reset(new Cell_division<Po_cell>(n_max, seed));
            cudaMemcpyToSymbolAsync(models::d_prolif_rate, &prolif_rate,
                sizeof(float), 0, cudaMemcpyHostToDevice, model_stream());
            division->divide<models::growth_division_rate, models::growth_inherit>(
                cells, mean_dist);
            return 0;
        }
        // The whole iteration -- both Heun stages with the counter resets as
        // (capturable) generic force, then the division kernels -- replays
        // from one graph. reset_counters is handed n_max there: clearing the
        // counters of dead slots as well is harmless.
        return replay_iteration(dt);
#else
        enqueue_iteration(dt);
        return 0;
#endif
    }
};

// ---- 7-float branching cell --------------------------------------------------------
struct Branching_sim : Typed_sim<models::Cell> {
    Branching_sim(int n_max, int grid_size, float cube_size)
        : Typed_sim<models::Cell>{n_max, grid_size, cube_size}
    {
#ifdef YALLA_B200
        // the counter reset is two cudaMemsetAsync on the step's stream
        cells.capture_generic_forces = true;
#endif
    }
    int set_param(const std::string& name, double value) override
    {
        const int fixed = Base::set_fix(name, value);
        return fixed <= 0 ? fixed : yb_sim::set_param(name, value);
    }
    int step(float dt) override
    {
        this->bind();
        auto reset_nbs = [this](const int n, const models::Cell* __restrict__ d_X,
                             models::Cell* d_dX) { reset_counters(n); };
        cells.take_step<models::epi_turing_mes_noturing>(dt, reset_nbs);
        return 0;
    }
};

// ---- branching cell + division + protrusions rewired every step ----------------
// BASELINE.json configs[3]: per iteration the neighbour grid
// of the cells is rebuilt (Grid::build, cube size r_protrusion), one protrusion
// per cell is rewired at random among its neighbours (update_protrusions,
// curand), the tissue takes a Heun step with counter reset + link_forces as
// generic force, and cells divide (proliferate_branching, curand).
struct Branching_growth_sim : Typed_sim<models::Cell> {
    Links protrusions;
    Grid grid;
    curandState* d_state = nullptr;
    int* d_n_at_launch = nullptr;
    float mes_rate = 0.006f, epi_rate = 0.006f, mean_dist = 0.75f;
    int seed = 4;
    bool seeded = false;

    Branching_growth_sim(int n_max, int grid_size, float cube_size)
        : Typed_sim<models::Cell>{n_max, grid_size, cube_size},
          protrusions{n_max * models::prots_per_cell, 0.2f},
          // cubes of edge r_protrusion = 2 cover the same volume with half the
          // cubes per axis
          grid{n_max, grid_size / 2 + 2}
    {
        cudaMalloc(&d_state, sizeof(curandState) * size_t(n_max));
        cudaMalloc(&d_n_at_launch, sizeof(int));
        protrusions.set_d_n(0);
    }
    ~Branching_growth_sim() override
    {
#ifdef YALLA_B200
        close_recording();
#endif
        cudaFree(d_n_at_launch);
        cudaFree(d_state);
    }
    int set_param(const std::string& name, double value) override
    {
#ifdef YALLA_B200
        drop_iteration();  // parameters are baked into the recorded launches
#endif
        if (name == "mes_rate") {
            mes_rate = static_cast<float>(value);
        } else if (name == "epi_rate") {
            epi_rate = static_cast<float>(value);
        } else if (name == "mean_dist") {
            mean_dist = static_cast<float>(value);
        } else if (name == "link_strength") {
            protrusions.strength = static_cast<float>(value);
        } else if (name == "seed") {
            seed = static_cast<int>(value);
            seeded = false;
        } else {
            const int fixed = Base::set_fix(name, value);
            return fixed <= 0 ? fixed : yb_sim::set_param(name, value);
        }
        return YB_OK;
    }
    int get_links(int* h_links, int capacity, int* n_out) override
    {
        cudaStreamSynchronize(model_stream());
        protrusions.copy_to_host();
        const int n = *protrusions.h_n;
        if (n > capacity) return fail(YB_EINVAL, "capacity < n_links");
        memcpy(h_links, protrusions.h_link, sizeof(Link) * size_t(n));
        if (n_out) *n_out = n;
        return check_cuda("get_links");
    }

    void enqueue_iteration(float dt)
#ifdef YALLA_B200
        override
#endif
    {
        const int n_max = cells.n_max;
        const cudaStream_t s = model_stream();
        const int n_links_max = n_max * models::prots_per_cell;
        // intercalation_w_gradient.cu:236-241, with the counts left on the device
        models::set_link_count<<<1, 1, 0, s>>>(cells.d_n, protrusions.d_n);
#ifdef YALLA_B200
        grid.stream = s;
        grid.build_live(cells.d_n, cells.d_X, models::r_protrusion);
#else
        grid.build(cells, models::r_protrusion);
#endif
        models::update_protrusions<<<(n_links_max + 32 - 1) / 32, 32, 0, s>>>(
            cells.d_n, grid.d_grid, cells.d_X, protrusions.d_state,
            protrusions.d_link);
        auto intercalation = [this](const int n, const models::Cell* __restrict__ d_X,
                                 models::Cell* d_dX) {
            reset_counters(n);
            link_forces(protrusions, d_X, d_dX);
        };
        cells.take_step<models::epi_turing_mes_noturing>(dt, intercalation);
        if (mes_rate > 0 || epi_rate > 0) {
            models::snapshot_count<<<1, 1, 0, s>>>(cells.d_n, d_n_at_launch);
            models::proliferate_branching<<<(n_max + 128 - 1) / 128, 128, 0, s>>>(
                mes_rate, epi_rate, mean_dist, n_max, d_state, cells.d_X,
                cells.d_old_v, cells.d_n, d_n_at_launch);
        }
    }

#ifdef YALLA_B200
    // ---- decomposed (b200/brick_links.cuh) --------------------------------------
    // One brick of the tissue per instance. Cells carry an identity and keep
    // their protrusion as the identity of the partner; both travel with the
    // cells and their ghost copies. Per iteration: adopt the daughters, survey
    // the neighbourhood (extra halo round), rewire with the example's kernel on
    // a Link array resolved over owned + ghost cells, then the decomposed Heun
    // step whose generic force resolves the links again for each stage's ghosts.
    yb::Brick_links brick_links;
    void register_model_arrays() override
    {
        Typed_sim<models::Cell>::register_model_arrays();
        brick_links.allocate(cells.n_max, models::prots_per_cell);
        cells.dom_register_array(brick_links.identity, sizeof(int), true);
        cells.dom_register_array(brick_links.partner,
            sizeof(int) * models::prots_per_cell, true);
    }
    // a freshly loaded tissue gets fresh identities (and no links)
    int slab_set_owned(const float* X, const float* v, int n_owned) override
    {
        brick_links.issued = false;
        return Typed_sim<models::Cell>::slab_set_owned(X, v, n_owned);
    }
    int get_ints(const std::string& name, int* values, int capacity) override
    {
        if (name != "identity" && name != "partner" && name != "unresolved_links")
            return Typed_sim<models::Cell>::get_ints(name, values, capacity);
        if (brick_links.identity == nullptr)
            return fail(YB_EINVAL, "identities exist in decomposed runs only");
        const int n = cells.get_d_n();  // waits for the model's stream
        if (name == "unresolved_links") {
            if (capacity < 1) return fail(YB_EINVAL, "capacity < 1");
            cudaMemcpy(values, brick_links.diagnostics, sizeof(int),
                cudaMemcpyDeviceToHost);
            return check_cuda("get_ints");
        }
        const int per_cell = name == "partner" ? models::prots_per_cell : 1;
        if (n * per_cell > capacity) return fail(YB_EINVAL, "capacity < n");
        cudaMemcpy(values,
            name == "partner" ? brick_links.partner : brick_links.identity,
            sizeof(int) * size_t(n) * per_cell, cudaMemcpyDeviceToHost);
        return check_cuda("get_ints");
    }
    int dom_step(float dt, int n_steps) override
    {
        if (!cells.dom.active || !cells.dom.connected())
            return fail(YB_EINVAL, "the domain is not connected to its neighbours");
        const int n_max = cells.n_max;
        const cudaStream_t s = model_stream();
        const int n_links_max = n_max * models::prots_per_cell;
        if (!seeded) {
            const int offset = 7919 * cells.dom.rank;  // own streams per brick
            setup_rand_states<<<(n_max + 128 - 1) / 128, 128, 0, s>>>(
                n_max, seed + offset, d_state);
            setup_rand_states<<<(n_links_max + 128 - 1) / 128, 128, 0, s>>>(
                n_links_max, seed + 1 + offset, protrusions.d_state);
            seeded = true;
        }
        this->bind();
        brick_links.rank = cells.dom.rank;
        auto intercalation = [this](const int n, const models::Cell* __restrict__ d_X,
                                 models::Cell* d_dX) {
            reset_counters(n);
            brick_links.resolve(yb::current_stage()->stream, cells.dom_ctl(),
                cells.d_n, protrusions.d_link, protrusions.d_n);
            link_forces(protrusions, d_X, d_dX);
        };
        // identities: for all cells before the first step, for the daughters
        // right after every division (so that they can be read back with them)
        if (!brick_links.issued) brick_links.issue(s, cells.dom_ctl(), cells.d_n);
        for (int k = 0; k < n_steps; k++) {
            cells.dom_adopt();
            cells.dom_survey();
            brick_links.resolve(
                s, cells.dom_ctl(), cells.d_n, protrusions.d_link, protrusions.d_n);
            grid.stream = s;
            grid.build_live(cells.d_n, cells.d_X, models::r_protrusion);
            models::update_protrusions<<<(n_links_max + 32 - 1) / 32, 32, 0, s>>>(
                cells.d_n, grid.d_grid, cells.d_X, protrusions.d_state,
                protrusions.d_link);
            brick_links.commit(s, cells.dom_ctl(), protrusions.d_link);
            cells.dom_end_survey();
            cells.dom_step<models::epi_turing_mes_noturing,
                friction_w_neighbour<models::Cell>>(dt, intercalation);
            if (mes_rate > 0 || epi_rate > 0) {
                models::snapshot_count<<<1, 1, 0, s>>>(cells.d_n, d_n_at_launch);
                models::proliferate_branching<<<(n_max + 128 - 1) / 128, 128, 0,
                    s>>>(mes_rate, epi_rate, mean_dist, n_max, d_state, cells.d_X,
                    cells.d_old_v, cells.d_n, d_n_at_launch);
                brick_links.issue(s, cells.dom_ctl(), cells.d_n);
            }
        }
        return check_cuda("yb_dom_step");
    }
#endif

    int step(float dt) override
    {
        const int n_max = cells.n_max;
        if (!seeded) {
            // fixed seeds for both generators (the Links constructor seeds its
            // states from the wall clock)
            setup_rand_states<<<(n_max + 128 - 1) / 128, 128, 0, model_stream()>>>(
                n_max, seed, d_state);
            const int n_links_max = n_max * models::prots_per_cell;
            setup_rand_states<<<(n_links_max + 128 - 1) / 128, 128, 0,
                model_stream()>>>(n_links_max, seed + 1, protrusions.d_state);
            seeded = true;
        }
        this->bind();
#ifdef YALLA_B200
        return replay_iteration(dt);
#else
        enqueue_iteration(dt);
        return 0;
#endif
    }
};

}  // namespace


namespace {
using Lanes4 = Lanes4_cell;

template<typename Pt>
int grid_build(const float* d_X, int n, int grid_size, float cube_size,
    int* d_cube_id, int* d_point_id, int* d_cube_start, int* d_cube_end)
{
    Grid grid{n, grid_size};
    grid.build(n, reinterpret_cast<const Pt*>(d_X), cube_size);
    const size_t cells = sizeof(int) * static_cast<size_t>(n);
    const size_t cubes = sizeof(int) * static_cast<size_t>(grid.n_cubes);
    cudaMemcpy(d_cube_id, grid.d_cube_id, cells, cudaMemcpyDeviceToDevice);
    cudaMemcpy(d_point_id, grid.d_point_id, cells, cudaMemcpyDeviceToDevice);
    cudaMemcpy(d_cube_start, grid.d_cube_start, cubes, cudaMemcpyDeviceToDevice);
    cudaMemcpy(d_cube_end, grid.d_cube_end, cubes, cudaMemcpyDeviceToDevice);
    cudaDeviceSynchronize();
    return check_cuda("yb_grid_build");
}
}  // namespace

namespace {
template<typename Pt>
int run_link_forces(const float* d_X, float* d_dX, int n, const int* d_links,
    int n_links, float strength)
{
    Links links{n_links, strength};
    cudaMemcpy(links.d_link, d_links, sizeof(Link) * static_cast<size_t>(n_links),
        cudaMemcpyDeviceToDevice);
    links.set_d_n(n_links);
#ifdef YALLA_B200
    // Outside a solver step nobody tells link_forces how many cells there are;
    // provide the context the solver would.
    yb::Stage_context context{n, n, 0};
    yb::current_stage() = &context;
#endif
    link_forces(links, reinterpret_cast<const Pt*>(d_X),
        reinterpret_cast<Pt*>(d_dX));
#ifdef YALLA_B200
    yb::current_stage() = nullptr;
#endif
    cudaDeviceSynchronize();
    return check_cuda("yb_link_forces");
}
}  // namespace

namespace {
__global__ void bending_pairs(
    const Po_cell* Xi, const Po_cell* Xj, int n_pairs, Po_cell* out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pairs) return;
    const Po_cell r = Xi[k] - Xj[k];
    const float dist = norm3df(r.x, r.y, r.z);
    out[k] = bending_force(Xi[k], r, dist);
}

__global__ void polarization_pairs(
    const Po_cell* Xi, const Po_cell* Xj, int n_pairs, Po_cell* out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pairs) return;
    const Polarity p{Xj[k].theta, Xj[k].phi};
    out[k] = bidirectional_polarization_force(Xi[k], p);
}

template<typename Kernel>
int run_pairs(Kernel kernel, const float* h_Xi, const float* h_Xj, int n_pairs,
    float* h_out)
{
    if (n_pairs <= 0) return fail(YB_EINVAL, "bad size");
    const size_t bytes = sizeof(Po_cell) * static_cast<size_t>(n_pairs);
    Po_cell *d_Xi, *d_Xj, *d_out;
    cudaMalloc(&d_Xi, bytes);
    cudaMalloc(&d_Xj, bytes);
    cudaMalloc(&d_out, bytes);
    cudaMemcpy(d_Xi, h_Xi, bytes, cudaMemcpyHostToDevice);
    cudaMemcpy(d_Xj, h_Xj, bytes, cudaMemcpyHostToDevice);
    kernel<<<(n_pairs + 127) / 128, 128>>>(d_Xi, d_Xj, n_pairs, d_out);
    cudaMemcpy(h_out, d_out, bytes, cudaMemcpyDeviceToHost);
    cudaFree(d_Xi);
    cudaFree(d_Xj);
    cudaFree(d_out);
    return check_cuda("polarity pairs");
}
}  // namespace


extern "C" {

const char* yb_build_info(void)
{
#ifdef YALLA_B200
    return "yalla-b200 (B200-native headers, sm_100a)";
#else
    return "yalla-reference (unmodified reference headers, sm_100a)";
#endif
}

const char* yb_last_error(void) { return last_error.c_str(); }

int yb_sim_create(const char* model, int n_max, int grid_size, float cube_size,
    yb_sim** out)
{
    if (!model || !out || n_max <= 0) return fail(YB_EINVAL, "bad argument");
    const std::string name = model;
    if (grid_size <= 0) grid_size = 50;
    if (!(cube_size > 0)) cube_size = 1.f;
    yb_sim* sim = nullptr;
    if (name == "springs")
        sim = new Spring_sim<Tile_solver, models::spring>(n_max);
    else if (name == "spring_tile")
        sim = new Spring_sim<Tile_solver, models::clipped_spring>(n_max);
    else if (name == "spring_grid")
        sim = new Spring_sim<Grid_solver, models::clipped_spring>(
            n_max, grid_size, cube_size);
    else if (name == "relu_tile")
        sim = new Spring_sim<Tile_solver, relu_force<float3>>(n_max);
    else if (name == "relu_grid")
        sim = new Spring_sim<Grid_solver, relu_force<float3>>(
            n_max, grid_size, cube_size);
    else if (name == "relu_gabriel")
        sim = new Spring_sim<Gabriel_solver, relu_force<float3>>(
            n_max, grid_size, cube_size);
    else if (name == "protrusions")
        sim = new Protrusion_sim(n_max, grid_size, cube_size);
    else if (name == "epithelium")
        sim = new Epithelium_sim(n_max, grid_size, cube_size);
    else if (name == "growth")
        sim = new Growth_sim(n_max, grid_size, cube_size);
    else if (name == "branching")
        sim = new Branching_sim(n_max, grid_size, cube_size);
    else if (name == "branching_growth")
        sim = new Branching_growth_sim(n_max, grid_size, cube_size);
    else
        return fail(YB_EINVAL, "unknown model " + name);
    const int status = check_cuda("yb_sim_create");
    if (status != YB_OK) {
        delete sim;
        return status;
    }
    *out = sim;
    return YB_OK;
}

void yb_sim_destroy(yb_sim* sim)
{
    cudaDeviceSynchronize();
    delete sim;
}

int yb_sim_lanes(const yb_sim* sim) { return sim->lanes(); }
int yb_sim_n_max(const yb_sim* sim) { return sim->n_max(); }

int yb_sim_set_param(yb_sim* sim, const char* name, double value)
{
    return sim->set_param(name, value);
}

int yb_sim_set_state(yb_sim* sim, const float* h_X, int n, int reset_v)
{
    return sim->set_state(h_X, n, reset_v);
}

int yb_sim_get_state(yb_sim* sim, float* h_X, int capacity, int* n_out)
{
    return sim->get_state(h_X, capacity, n_out);
}

int yb_sim_get_velocities(yb_sim* sim, float* h_v, int capacity)
{
    return sim->get_velocities(h_v, capacity);
}

int yb_sim_set_velocities(yb_sim* sim, const float* h_v, int n)
{
    return sim->set_velocities(h_v, n);
}

int yb_sim_set_ints(yb_sim* sim, const char* name, const int* h_values, int n)
{
    return sim->set_ints(name, h_values, n);
}

int yb_sim_get_ints(yb_sim* sim, const char* name, int* h_values, int capacity)
{
    return sim->get_ints(name, h_values, capacity);
}

int yb_sim_seed_sphere(yb_sim* sim, int n, float dist_to_nb,
    unsigned long long seed, int relax_steps)
{
    return sim->seed_sphere(n, dist_to_nb, seed, relax_steps);
}

int yb_sim_set_links(yb_sim* sim, const int* h_links, int n_links)
{
    return sim->set_links(h_links, n_links);
}

int yb_sim_get_links(yb_sim* sim, int* h_links, int capacity, int* n_out)
{
    return sim->get_links(h_links, capacity, n_out);
}

int yb_sim_step(yb_sim* sim, float dt, int n_steps)
{
    for (int k = 0; k < n_steps; k++) sim->step(dt);
    return check_cuda("yb_sim_step");
}

int yb_sim_step_timed(yb_sim* sim, float dt, int n_steps, float* ms_out,
    long long* cell_updates_out)
{
    const cudaStream_t stream = sim->work_stream();
    cudaEvent_t start, stop;
    cudaEventCreate(&start);
    cudaEventCreate(&stop);
    // cell count at the start of every step, snapshotted asynchronously into
    // pinned memory so that counting does not serialise the steps
    int* h_counts = nullptr;
    cudaMallocHost(&h_counts, sizeof(int) * static_cast<size_t>(n_steps + 1));
    const int* d_count = sim->count_on_device();
    cudaDeviceSynchronize();
    cudaEventRecord(start, stream);
    for (int k = 0; k < n_steps; k++) {
        cudaMemcpyAsync(
            h_counts + k, d_count, sizeof(int), cudaMemcpyDeviceToHost, stream);
        sim->step(dt);
    }
    cudaEventRecord(stop, stream);
    cudaEventSynchronize(stop);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, start, stop);
    cudaEventDestroy(start);
    cudaEventDestroy(stop);
    long long updates = 0;
    for (int k = 0; k < n_steps; k++) updates += h_counts[k];
    cudaFreeHost(h_counts);
    if (ms_out) *ms_out = ms;
    if (cell_updates_out) *cell_updates_out = updates;
    return check_cuda("yb_sim_step_timed");
}

int yb_sim_step_host(yb_sim* sim, const float* h_in, int n, float dt,
    int n_steps, float* h_out, int capacity, int* n_out)
{
    return sim->step_host(h_in, n, dt, n_steps, h_out, capacity, n_out);
}

int yb_sim_set_stream(yb_sim* sim, void* stream)
{
    return sim->set_stream(stream);
}

int yb_sim_step_host_async(yb_sim* sim, const float* h_in, int n, float dt,
    int n_steps, float* h_out, int out_cells, int* h_n_out)
{
    return sim->step_host_async(h_in, n, dt, n_steps, h_out, out_cells, h_n_out);
}

int yb_sim_host_drain(yb_sim* sim) { return sim->host_drain(); }

int yb_dd_load(yb_sim* sim, int stage, const float* X_owned,
    const float* v_owned, int n_owned, const float* X_ghost,
    const float* v_ghost, int n_ghost)
{
    return sim->dd_load(stage, X_owned, v_owned, n_owned, X_ghost, v_ghost, n_ghost);
}

int yb_dd_forces(yb_sim* sim, int stage, float* sums4)
{
    return sim->dd_forces(stage, sums4);
}

int yb_dd_update(yb_sim* sim, int stage, float dt, const float* mean3)
{
    return sim->dd_update(stage, dt, mean3);
}

int yb_dd_read(yb_sim* sim, int which, float* out, int n)
{
    return sim->dd_read(which, out, n);
}

int yb_slab_begin(yb_sim* sim, float z_lo, float z_hi, float halo,
    int capacity, int first_layer, int n_layers)
{
    return sim->slab_begin(z_lo, z_hi, halo, capacity, first_layer, n_layers);
}

int yb_slab_set_owned(yb_sim* sim, const float* X, const float* v, int n_owned)
{
    return sim->slab_set_owned(X, v, n_owned);
}

int yb_slab_pack(yb_sim* sim, int what, float* send_lo, float* send_hi)
{
    return sim->slab_pack(what, send_lo, send_hi);
}

int yb_slab_unpack(yb_sim* sim, int what, const float* recv_lo,
    const float* recv_hi)
{
    return sim->slab_unpack(what, recv_lo, recv_hi);
}

int yb_slab_update(yb_sim* sim, int stage, float dt, const float* sums4)
{
    return sim->slab_update(stage, dt, sums4);
}

int yb_slab_counts(yb_sim* sim, int* n_owned, int* n_total, int* problems)
{
    return sim->slab_counts(n_owned, n_total, problems);
}

int yb_dom_begin(yb_sim* sim, int rank, int world, const float* lo3,
    const float* hi3, float halo, const int* peer_ranks27,
    const int* capacity27, const int* box_first3, const int* box_n3)
{
    return sim->dom_begin(rank, world, lo3, hi3, halo, peer_ranks27, capacity27,
        box_first3, box_n3);
}

int yb_dom_register_array(yb_sim* sim, void* d_array, int bytes_per_cell,
    int ghosts_too)
{
    return sim->dom_register_array(d_array, bytes_per_cell, ghosts_too);
}

int yb_dom_exchange(yb_sim* sim, void** d_base_out, long long* bytes_out,
    long long* offsets27xK_out)
{
    return sim->dom_exchange(d_base_out, bytes_out, offsets27xK_out);
}

int yb_dom_connect(yb_sim* sim, int direction, void* d_peer_base,
    const long long* peer_offsets)
{
    return sim->dom_connect(direction, d_peer_base, peer_offsets);
}

int yb_dom_connect_mailbox(yb_sim* sim, int rank, void* d_peer_base)
{
    return sim->dom_connect_mailbox(rank, d_peer_base);
}

int yb_dom_seed_lattice_ball(yb_sim* sim, float radius, float dist_to_nb,
    float jitter, unsigned long long seed, int* n_out)
{
    return sim->dom_seed_lattice_ball(radius, dist_to_nb, jitter, seed, n_out);
}

int yb_dom_step(yb_sim* sim, float dt, int n_steps)
{
    return sim->dom_step(dt, n_steps);
}

int yb_dom_read_profile(yb_sim* sim, float* ms8)
{
    return sim->dom_read_profile(ms8);
}

int yb_ipc_export(const void* d_base, unsigned char* handle64)
{
    cudaIpcMemHandle_t handle;
    static_assert(sizeof(handle) == 64, "CUDA IPC handles are 64 bytes");
    if (cudaIpcGetMemHandle(&handle, const_cast<void*>(d_base)) != cudaSuccess)
        return check_cuda("yb_ipc_export");
    memcpy(handle64, &handle, sizeof(handle));
    return YB_OK;
}

int yb_ipc_import(const unsigned char* handle64, void** d_base_out)
{
    cudaIpcMemHandle_t handle;
    memcpy(&handle, handle64, sizeof(handle));
    if (cudaIpcOpenMemHandle(d_base_out, handle, cudaIpcMemLazyEnablePeerAccess) !=
        cudaSuccess)
        return check_cuda("yb_ipc_import");
    return YB_OK;
}

int yb_ipc_release(void* d_base)
{
    cudaIpcCloseMemHandle(d_base);
    return check_cuda("yb_ipc_release");
}

int yb_sim_profile_sweeps(yb_sim* sim, int enable)
{
    return sim->profile_sweeps(enable);
}

int yb_sim_read_sweep_profile(yb_sim* sim, float* total_ms, int* launches)
{
    return sim->read_sweep_profile(total_ms, launches);
}

int yb_sim_n(yb_sim* sim, int* n_out)
{
    *n_out = sim->current_n();
    return check_cuda("yb_sim_n");
}

int yb_sim_sync(yb_sim* sim)
{
    cudaDeviceSynchronize();
    return check_cuda("yb_sim_sync");
}


// ---- grid build ---------------------------------------------------------------

int yb_grid_build(const float* d_X, int lanes, int n, int grid_size,
    float cube_size, int* d_cube_id, int* d_point_id, int* d_cube_start,
    int* d_cube_end)
{
    if (n <= 0 || grid_size <= 0) return fail(YB_EINVAL, "bad size");
    switch (lanes) {
        case 3:
            return grid_build<float3>(d_X, n, grid_size, cube_size, d_cube_id,
                d_point_id, d_cube_start, d_cube_end);
        case 4:
            return grid_build<Lanes4>(d_X, n, grid_size, cube_size, d_cube_id,
                d_point_id, d_cube_start, d_cube_end);
        case 5:
            return grid_build<Po_cell>(d_X, n, grid_size, cube_size, d_cube_id,
                d_point_id, d_cube_start, d_cube_end);
        case 7:
            return grid_build<models::Cell>(d_X, n, grid_size, cube_size,
                d_cube_id, d_point_id, d_cube_start, d_cube_end);
    }
    return fail(YB_EINVAL, "lanes must be 3, 4, 5 or 7");
}

int yb_nhood(int grid_size, int* h_nhood27)
{
    // The table is written by the grid solver's constructor.
    Solution<float3, Grid_solver> probe{1, grid_size};
    cudaMemcpyFromSymbol(h_nhood27, d_nhood, 27 * sizeof(int));
    return check_cuda("yb_nhood");
}


// ---- link forces --------------------------------------------------------------

int yb_link_forces(const float* d_X, float* d_dX, int lanes, int n,
    const int* d_links, int n_links, float strength)
{
    if (n <= 0 || n_links <= 0) return fail(YB_EINVAL, "bad size");
    switch (lanes) {
        case 3:
            return run_link_forces<float3>(
                d_X, d_dX, n, d_links, n_links, strength);
        case 4:
            return run_link_forces<Lanes4>(
                d_X, d_dX, n, d_links, n_links, strength);
        case 5:
            return run_link_forces<Po_cell>(
                d_X, d_dX, n, d_links, n_links, strength);
        case 7:
            return run_link_forces<models::Cell>(
                d_X, d_dX, n, d_links, n_links, strength);
    }
    return fail(YB_EINVAL, "lanes must be 3, 4, 5 or 7");
}


// ---- polarity forces -------------------------------------------------------------

int yb_bending_force(
    const float* h_Xi, const float* h_Xj, int n_pairs, float* h_out)
{
    return run_pairs(bending_pairs, h_Xi, h_Xj, n_pairs, h_out);
}

int yb_polarization_force(
    const float* h_Xi, const float* h_Xj, int n_pairs, float* h_out)
{
    return run_pairs(polarization_pairs, h_Xi, h_Xj, n_pairs, h_out);
}

}  // extern "C"
