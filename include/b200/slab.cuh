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
// Packing is a stable stream compaction (flags -> scan_bins -> scatter), so the
// order of ghosts and migrants, and with it every floating-point sum on the
// receiving side, is reproducible.
#pragma once

#include <cuda_runtime.h>

#include "grid_build.cuh"
#include "layout.cuh"

namespace yb {

constexpr int SLAB_HEADER = 4;  // floats in front of the records

// Which owned cells go to the lower / upper neighbour. Halo: lo_edge = z_lo +
// halo, hi_edge = z_hi - halo (a cell may go both ways in a thin slab).
// Migration: lo_edge = z_lo, hi_edge = z_hi, and `stay` marks the rest.
template<typename Pt>
__global__ void __launch_bounds__(256) slab_flags(const Step_ctl* ctl,
    const Pt* __restrict__ P, float lo_edge, float hi_edge, int has_lower,
    int has_upper, int* __restrict__ flag_lo, int* __restrict__ flag_hi,
    int* __restrict__ flag_stay)
{
    const int n = ctl->n_owned;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += gridDim.x * blockDim.x) {
        const float z = __ldg(reinterpret_cast<const float*>(P + i) + 2);
        const int lo = has_lower && z < lo_edge;
        const int hi = has_upper && z >= hi_edge;
        flag_lo[i] = lo;
        flag_hi[i] = hi;
        if (flag_stay) flag_stay[i] = !(lo || hi);
    }
}

template<typename Pt>
__device__ __forceinline__ void write_record(
    float* record, const Pt* P, const float3* v, int i)
{
    using L = Layout<Pt>;
    const float* x = reinterpret_cast<const float*>(P + i);
    const float* w = reinterpret_cast<const float*>(v + i);
#pragma unroll
    for (int k = 0; k < L::lanes; k++) record[k] = x[k];
    record[L::lanes + 0] = w[0];
    record[L::lanes + 1] = w[1];
    record[L::lanes + 2] = w[2];
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

// off_*: exclusive scans of the flags (n + 1 valid entries). Cells whose flag
// was set are those with off[i + 1] != off[i].
template<typename Pt>
__global__ void __launch_bounds__(256) slab_pack(Step_ctl* ctl,
    const Pt* __restrict__ P, const float3* __restrict__ v,
    const int* __restrict__ off_lo, const int* __restrict__ off_hi,
    float* __restrict__ send_lo, float* __restrict__ send_hi, int capacity)
{
    constexpr int W = Layout<Pt>::lanes + 3;
    const int n = ctl->n_owned;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += gridDim.x * blockDim.x) {
        const int a = off_lo[i], b = off_hi[i];
        if (off_lo[i + 1] != a && a < capacity)
            write_record(send_lo + SLAB_HEADER + size_t(a) * W, P, v, i);
        if (off_hi[i + 1] != b && b < capacity)
            write_record(send_hi + SLAB_HEADER + size_t(b) * W, P, v, i);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int n_lo = off_lo[n], n_hi = off_hi[n];
        if (n_lo > capacity || n_hi > capacity) atomicAdd(&ctl->out_of_grid, 1 << 20);
        send_lo[0] = __int_as_float(min(n_lo, capacity));
        send_hi[0] = __int_as_float(min(n_hi, capacity));
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
    if (blockIdx.x == 0 && threadIdx.x == 0) *d_n = n + n_lo + n_hi;
}

// Migration, step 1: stable compaction of the cells that stay into scratch.
template<typename Pt>
__global__ void __launch_bounds__(256) slab_compact_stayers(const Step_ctl* ctl,
    const Pt* __restrict__ X, const float3* __restrict__ v,
    const int* __restrict__ off_stay, Pt* __restrict__ X_tmp,
    float3* __restrict__ v_tmp)
{
    const int n = ctl->n_owned;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += gridDim.x * blockDim.x) {
        const int to = off_stay[i];
        if (off_stay[i + 1] == to) continue;
        store_pt(X_tmp, to, load_pt(X, i));
        v_tmp[to] = v[i];
    }
}

// Migration, step 2: owned cells := stayers, arrivals from below, from above.
template<typename Pt>
__global__ void __launch_bounds__(256) slab_merge(const Step_ctl* ctl,
    const int* __restrict__ off_stay, const Pt* __restrict__ X_tmp,
    const float3* __restrict__ v_tmp, const float* __restrict__ recv_lo,
    const float* __restrict__ recv_hi, int has_lower, int has_upper, int n_max,
    Pt* X, float3* v, int* new_count)
{
    constexpr int W = Layout<Pt>::lanes + 3;
    const int n_stay = off_stay[ctl->n_owned];
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

// Scratch of the compactions: three flag arrays and their scans.
struct Slab_scratch {
    int capacity = 0;     // records per exchange buffer
    int n_tiles = 0;      // scan tiles covering n_max + 1 entries
    int* flag[3] = {nullptr, nullptr, nullptr};
    int* off[3] = {nullptr, nullptr, nullptr};
    unsigned long long* status = nullptr;
    Step_ctl* scan_ctl = nullptr;  // scan bookkeeping only
    int* new_count = nullptr;
    float z_lo = 0.f, z_hi = 0.f, halo = 0.f;
    int has_lower = 0, has_upper = 0;

    void allocate(int n_max)
    {
        const size_t bins = static_cast<size_t>(scan_padded(n_max + 1));
        n_tiles = static_cast<int>(bins / SCAN_TILE);
        for (int k = 0; k < 3; k++) {
            YB_CUDA(cudaMalloc(&flag[k], bins * sizeof(int)));
            YB_CUDA(cudaMemset(flag[k], 0, bins * sizeof(int)));
            YB_CUDA(cudaMalloc(&off[k], bins * sizeof(int)));
            YB_CUDA(cudaMemset(off[k], 0, bins * sizeof(int)));
        }
        YB_CUDA(cudaMalloc(&status, n_tiles * sizeof(unsigned long long)));
        YB_CUDA(cudaMemset(status, 0, n_tiles * sizeof(unsigned long long)));
        YB_CUDA(cudaMalloc(&scan_ctl, sizeof(Step_ctl)));
        Step_ctl fresh{};
        fresh.scan_epoch = 1;
        YB_CUDA(cudaMemcpy(
            scan_ctl, &fresh, sizeof(fresh), cudaMemcpyHostToDevice));
        YB_CUDA(cudaMalloc(&new_count, sizeof(int)));
    }
    void release()
    {
        for (int k = 0; k < 3; k++) {
            cudaFree(flag[k]);
            cudaFree(off[k]);
        }
        cudaFree(status);
        cudaFree(scan_ctl);
        cudaFree(new_count);
    }
};

}  // namespace yb
