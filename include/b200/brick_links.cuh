// Internal: Links between the cells of a brick-decomposed tissue (b200/domain.cuh).
//
// ya||a's Links (links.cuh:16-92) name their two ends by cell index. In a
// decomposed tissue an index means something only on one brick and only until
// the next migration pass: cells are re-stored in cube order every step, they
// change owner, and a partner across a cut is a ghost whose index differs from
// stage to stage. So every cell carries an IDENTITY -- a number that stays with
// it for life -- and every link is kept by the cell at its `a` end as the
// identity of the cell at its `b` end. Both arrays are registered with the
// domain (Solution::dom_register_array, ghosts_too): they migrate with their
// cells, and the ghost copies bring their links along, which is what lets the
// brick that owns `b` apply the link's pull on `b` while the brick that owns `a`
// applies it on `a` -- each end exactly once, no force crosses a cut.
//
// Whenever index-based code needs the links (the model's rewiring kernel,
// link_forces), resolve() turns them into a plain Link array over the cells
// present on this brick, owned and ghost, through a hash table from identity to
// index; commit() turns the owned cells' links back into identities.
//
//   identity   (rank << 26) | serial number on the rank that first saw the cell
//   table      open addressing, linear probing; a slot is live if it carries the
//              epoch of the current build, so the table is never cleared
#pragma once

#include <cuda_runtime.h>

#include "domain.cuh"

namespace yb {

struct Identity_table {  // kernel argument
    unsigned long long* keys;  // epoch << 32 | identity
    int* values;               // index of the cell on this brick
    unsigned mask;             // slots - 1 (a power of two)
    unsigned epoch;
};

__device__ __forceinline__ unsigned identity_hash(int id)
{
    unsigned h = static_cast<unsigned>(id) * 2654435761u;
    return h ^ (h >> 15);
}

__device__ __forceinline__ int find_identity(const Identity_table& table, int id)
{
    const unsigned long long wanted =
        (static_cast<unsigned long long>(table.epoch) << 32) |
        static_cast<unsigned>(id);
    unsigned h = identity_hash(id) & table.mask;
    for (unsigned probes = 0; probes <= table.mask; probes++) {
        const unsigned long long here = table.keys[h];
        if (here == wanted) return table.values[h];
        if (static_cast<unsigned>(here >> 32) != table.epoch) return -1;
        h = (h + 1) & table.mask;
    }
    return -1;
}

// identity -> index for all cells present (owned and ghosts)
__global__ void __launch_bounds__(256) index_identities(
    const int* __restrict__ d_n, int n_max, const int* __restrict__ identity,
    Identity_table table)
{
    const int n = live_cells(d_n, n_max);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += gridDim.x * blockDim.x) {
        const int id = identity[i];
        const unsigned long long mine =
            (static_cast<unsigned long long>(table.epoch) << 32) |
            static_cast<unsigned>(id);
        unsigned h = identity_hash(id) & table.mask;
        while (true) {
            const unsigned long long here = table.keys[h];
            if (static_cast<unsigned>(here >> 32) != table.epoch) {  // free
                if (atomicCAS(&table.keys[h], here, mine) == here) {
                    table.values[h] = i;
                    break;
                }
                continue;  // somebody else took it: look again
            }
            if (here == mine) break;  // (the same cell twice: keep the first)
            h = (h + 1) & table.mask;
        }
    }
}

// New cells -- all cells before the first step, afterwards those the model has
// appended behind the owned ones (first_new = -1: from ctl->n_owned) -- get an
// identity and links that point at themselves (no link: a == b).
__global__ void __launch_bounds__(256) issue_identities(const Step_ctl* ctl,
    const int* __restrict__ d_n, int n_max, int first_new, int rank,
    const int* __restrict__ next_serial, int* __restrict__ identity,
    int* __restrict__ partner, int links_per_cell)
{
    const int first = first_new >= 0 ? first_new : ctl->n_owned;
    const int last = live_cells(d_n, n_max);
    const int serial = *next_serial;
    for (int i = first + blockIdx.x * blockDim.x + threadIdx.x; i < last;
         i += gridDim.x * blockDim.x) {
        const int id = (rank << 26) | ((serial + (i - first)) & ((1 << 26) - 1));
        identity[i] = id;
        for (int k = 0; k < links_per_cell; k++)
            partner[size_t(i) * links_per_cell + k] = id;
    }
}

__global__ void count_issued_identities(const Step_ctl* ctl, const int* d_n,
    int n_max, int first_new, int* next_serial)
{
    const int first = first_new >= 0 ? first_new : ctl->n_owned;
    const int last = live_cells(d_n, n_max);
    if (last > first) *next_serial += last - first;
}

// Link q = c * links_per_cell + k of every cell c present: {c, index of its
// partner}, or {c, c} (ya||a's "no link") where the partner is not on this
// brick. Of the ghosts' links only those that end on an owned cell matter here.
// diagnostics[0]: links of owned cells whose partner is out of reach.
template<typename Link_t>
__global__ void __launch_bounds__(256) resolve_links(const Step_ctl* ctl,
    const int* __restrict__ d_n, int n_max, const int* __restrict__ identity,
    const int* __restrict__ partner, int links_per_cell, Identity_table table,
    Link_t* __restrict__ link, int* __restrict__ d_n_links, int* diagnostics)
{
    const int n = live_cells(d_n, n_max);
    const int n_owned = ctl->n_owned;
    const int n_links = n * links_per_cell;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n_links;
         q += gridDim.x * blockDim.x) {
        const int c = q / links_per_cell;
        const int wanted = partner[q];
        int b = c;
        if (wanted != identity[c]) {
            b = find_identity(table, wanted);
            if (b < 0) {
                if (c < n_owned) atomicAdd(diagnostics, 1);
                b = c;
            } else if (c >= n_owned && b >= n_owned) {
                b = c;  // between two ghosts: somebody else's business
            }
        }
        link[q].a = c;
        link[q].b = b;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *d_n_links = n_links;
}

// The owned cells' links, as the model's kernels left them, back to identities.
template<typename Link_t>
__global__ void __launch_bounds__(256) commit_links(const Step_ctl* ctl,
    const Link_t* __restrict__ link, const int* __restrict__ identity,
    int* __restrict__ partner, int links_per_cell)
{
    const int n_links = ctl->n_owned * links_per_cell;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n_links;
         q += gridDim.x * blockDim.x)
        partner[q] = identity[link[q].b];
}

// Host side: the two per-cell arrays and the table.
struct Brick_links {
    int n_max = 0, links_per_cell = 1, rank = 0;
    int* identity = nullptr;  // [n_max]
    int* partner = nullptr;   // [n_max * links_per_cell]
    int* next_serial = nullptr;
    int* diagnostics = nullptr;
    Identity_table table{};
    bool issued = false;

    void allocate(int n_max_, int links_per_cell_)
    {
        release();
        n_max = n_max_, links_per_cell = links_per_cell_;
        const size_t cells = n_max > 0 ? n_max : 1;
        YB_CUDA(cudaMalloc(&identity, cells * sizeof(int)));
        YB_CUDA(cudaMalloc(&partner, cells * links_per_cell * sizeof(int)));
        YB_CUDA(cudaMemset(identity, 0xff, cells * sizeof(int)));
        YB_CUDA(cudaMemset(partner, 0xff, cells * links_per_cell * sizeof(int)));
        YB_CUDA(cudaMalloc(&next_serial, sizeof(int)));
        YB_CUDA(cudaMemset(next_serial, 0, sizeof(int)));
        YB_CUDA(cudaMalloc(&diagnostics, 4 * sizeof(int)));
        YB_CUDA(cudaMemset(diagnostics, 0, 4 * sizeof(int)));
        size_t slots = 1024;
        while (slots < 2 * cells) slots *= 2;
        YB_CUDA(cudaMalloc(&table.keys, slots * sizeof(unsigned long long)));
        YB_CUDA(cudaMemset(table.keys, 0, slots * sizeof(unsigned long long)));
        YB_CUDA(cudaMalloc(&table.values, slots * sizeof(int)));
        table.mask = static_cast<unsigned>(slots - 1);
        table.epoch = 0;  // slots zeroed = epoch 0 = free from the first build on
        issued = false;
    }
    void release()
    {
        if (identity == nullptr) return;
        cudaFree(table.values);
        cudaFree(table.keys);
        cudaFree(diagnostics);
        cudaFree(next_serial);
        cudaFree(partner);
        cudaFree(identity);
        identity = nullptr;
    }
    ~Brick_links() { release(); }

    // Before Solution::dom_adopt(): identities for the cells it is about to adopt
    // (the first time: for all cells).
    void issue(cudaStream_t s, const Step_ctl* ctl, const int* d_n)
    {
        const int blocks = stride_grid(n_max, 256, sm_count());
        const int first_new = issued ? -1 : 0;
        issue_identities<<<blocks, 256, 0, s>>>(ctl, d_n, n_max, first_new, rank,
            next_serial, identity, partner, links_per_cell);
        count_issued_identities<<<1, 1, 0, s>>>(
            ctl, d_n, n_max, first_new, next_serial);
        issued = true;
    }
    // identities -> Link array over the cells present (and its count)
    template<typename Link_t>
    void resolve(cudaStream_t s, const Step_ctl* ctl, const int* d_n, Link_t* link,
        int* d_n_links)
    {
        const int blocks = stride_grid(n_max, 256, sm_count());
        table.epoch++;
        index_identities<<<blocks, 256, 0, s>>>(d_n, n_max, identity, table);
        resolve_links<Link_t><<<blocks, 256, 0, s>>>(ctl, d_n, n_max, identity,
            partner, links_per_cell, table, link, d_n_links, diagnostics);
    }
    template<typename Link_t>
    void commit(cudaStream_t s, const Step_ctl* ctl, const Link_t* link)
    {
        const int blocks = stride_grid(n_max, 256, sm_count());
        commit_links<Link_t><<<blocks, 256, 0, s>>>(
            ctl, link, identity, partner, links_per_cell);
    }
};

}  // namespace yb
