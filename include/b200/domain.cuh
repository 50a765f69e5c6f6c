// Internal: brick decomposition of one tissue over the GPUs of a node, with the
// halo exchange, the migration of cells and the global drift sum done by the
// kernels themselves over peer memory (NVLink P2P) -- no NCCL call, no host
// synchronisation and no copy engine in a step.
//
// ya||a is single-GPU; this is the extension BASELINE.json's
// north_star asks for. The tissue is cut into bx * by * bz bricks along cube
// boundaries, one process per GPU and brick. Interactions are strictly shorter
// than cube_size, so a brick needs copies ("ghosts") of the neighbouring
// bricks' cells within one cube of its faces, edges and corners: up to 26
// neighbours. Every rank owns one EXCHANGE allocation that its neighbours map
// into their address space (CUDA IPC; plain pointers inside one process):
//
//   [mailbox: {sum dX, n} of every rank for the drift, two parities, + flags]
//   [flags:   one word per (neighbour, round kind)]
//   [inboxes: per (neighbour, round kind) a header and `capacity` records]
//
// Round kinds: 0 / 1 = halo of the predictor / corrector stage, 2 = migration,
// 3 = survey (a halo of the positions ahead of the step, asked for by models
// whose own kernels need the neighbours: dom_survey).
// A round of rank A: a stable stream compaction (dd_tile_counts ->
// dd_tile_offsets -> dd_pack, entry order preserved per destination) packs the
// selected records into A's local outboxes, and dd_push streams every outbox
// into the neighbour's inbox with coalesced 16-byte stores over NVLink (packing
// straight into peer memory was measured first: record-sized scattered remote
// stores run at ~35 GB/s and stall the compaction) and whose last CTA -- after
// a system-scope fence -- stores the round's epoch into the neighbours' flag
// words. B's stream meanwhile sits in dd_wait until its flags show the epoch,
// then dd_append_ghosts / dd_merge read the inboxes. An inbox is
// written again one step later; by then its owner has consumed it, because the
// writer had to receive two later rounds from that owner first.
//
// The drift (mean force, solvers.cuh:241-255) is global: dd_allreduce_drift
// stores this rank's {sum dX, n} into every rank's mailbox, waits for all
// slots of the round and adds them in rank order -- the same bits on every rank.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include "grid_build.cuh"
#include "layout.cuh"
#include "slab.cuh"

namespace yb {

constexpr int DD_MAX_PEERS = 26;
constexpr int DD_MAX_RANKS = 64;
constexpr int DD_ROUNDS = 4;  // halo X, halo X1, migration, survey (halo of X
                              // ahead of the step, for the model's own kernels)

// direction (dx, dy, dz) in {-1, 0, 1}^3 <-> index 0..26 (13 = the brick itself)
inline int dd_direction_index(int dx, int dy, int dz)
{
    return (dx + 1) + 3 * (dy + 1) + 9 * (dz + 1);
}

// What a brick owns and who surrounds it (kernel argument, by value).
struct Dd_region {
    float lo[3], hi[3];  // owns lo <= x < hi; -+INFINITY at the tissue's ends
    float halo;          // width of the strip copied to a neighbour
    int n_peers;
    signed char dir[DD_MAX_PEERS][4];  // offset of peer p in the brick grid
};

// Where the records for every peer go (peer memory; kernel argument).
struct Dd_outboxes {
    float* buffer[DD_MAX_PEERS];    // header + records, in the peer's allocation
    unsigned* flag[DD_MAX_PEERS];   // gets the round's epoch when complete
    int capacity[DD_MAX_PEERS];
};

// This rank's own inboxes of one round kind (kernel argument).
struct Dd_inboxes {
    const float* buffer[DD_MAX_PEERS];
    const unsigned* flag[DD_MAX_PEERS];
    int n_peers;
};

struct Dd_mailbox {
    float sums[2][DD_MAX_RANKS][4];  // [parity][rank] = {sum dX.xyz, n}
    unsigned flag[2][DD_MAX_RANKS];
};

struct Dd_mailboxes {  // every rank's mailbox, mapped here (kernel argument)
    Dd_mailbox* of_rank[DD_MAX_RANKS];
};

// Per-cell arrays of the model that travel with the cells (kernel argument):
// every one migrates with its cell and is re-stored with it; those the pairwise
// functor reads of a NEIGHBOUR (a cell type, ...) also come with the ghosts.
constexpr int DD_MAX_EXTRAS = 8;

struct Dd_extras {
    int count;
    unsigned* data[DD_MAX_EXTRAS];     // the model's array, n_max entries
    unsigned* scratch[DD_MAX_EXTRAS];  // same size: target of the re-store
    int words[DD_MAX_EXTRAS];          // 32-bit words per cell
    int ghosts_too[DD_MAX_EXTRAS];
};

__device__ __forceinline__ void write_extras(
    float* tail, const Dd_extras& extras, int i, bool migration)
{
    unsigned* out = reinterpret_cast<unsigned*>(tail);
    for (int k = 0; k < extras.count; k++) {
        if (!migration && !extras.ghosts_too[k]) continue;
        const unsigned* src = extras.data[k] + size_t(i) * extras.words[k];
        for (int w = 0; w < extras.words[k]; w++) *out++ = src[w];
    }
}

__device__ __forceinline__ void read_extras(
    const float* tail, const Dd_extras& extras, int at, bool migration)
{
    const unsigned* in = reinterpret_cast<const unsigned*>(tail);
    for (int k = 0; k < extras.count; k++) {
        if (!migration && !extras.ghosts_too[k]) continue;
        unsigned* dst = extras.data[k] + size_t(at) * extras.words[k];
        for (int w = 0; w < extras.words[k]; w++) dst[w] = *in++;
    }
}

__device__ __forceinline__ void store_release_sys(unsigned* p, unsigned value)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(value)
                 : "memory");
}

__device__ __forceinline__ unsigned load_acquire_sys(const unsigned* p)
{
    unsigned value;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];"
                 : "=r"(value)
                 : "l"(p)
                 : "memory");
    return value;
}

__device__ __forceinline__ unsigned long long global_nanoseconds()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Spin until *flag >= epoch. Gives up after ten seconds (a peer died): the run
// is then wrong, which `problems` reports, but the GPU does not hang.
__device__ __forceinline__ bool wait_for_epoch(const unsigned* flag, unsigned epoch)
{
    const unsigned long long start = global_nanoseconds();
    while (static_cast<int>(load_acquire_sys(flag) - epoch) < 0) {
        if (global_nanoseconds() - start > 10000000000ull) return false;
        __nanosleep(200);
    }
    return true;
}

// Which peers get a copy of (halo round) or take over (migration round) a cell
// at position x: bit p of the result. A peer in direction d gets the cell if
// the cell is within the halo of (resp. beyond) the face towards d on every
// axis where d is not 0; in a migration round the axes where d is 0 must be
// inside, so that exactly one peer takes the cell.
__device__ __forceinline__ unsigned dd_destinations(
    const float* x, const Dd_region& region, bool migration)
{
    bool below[3], above[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float inset = migration ? 0.f : region.halo;
        below[a] = x[a] < region.lo[a] + inset;
        above[a] = x[a] >= region.hi[a] - inset;
    }
    unsigned mask = 0;
    for (int p = 0; p < region.n_peers; p++) {
        bool takes = true;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const int d = region.dir[p][a];
            if (d < 0) takes = takes && below[a];
            if (d > 0) takes = takes && above[a];
            if (d == 0 && migration) takes = takes && !below[a] && !above[a];
        }
        mask |= (takes ? 1u : 0u) << p;
    }
    return mask;
}

// The same from the face flags a previous kernel left for the cell (halo rounds).
__device__ __forceinline__ unsigned dd_destinations_of_flags(
    unsigned flags, const Dd_region& region)
{
    if (flags == 0) return 0;
    unsigned mask = 0;
    for (int p = 0; p < region.n_peers; p++) {
        bool takes = true;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const int d = region.dir[p][a];
            if (d < 0) takes = takes && ((flags >> (2 * a)) & 1u);
            if (d > 0) takes = takes && ((flags >> (2 * a + 1)) & 1u);
        }
        mask |= (takes ? 1u : 0u) << p;
    }
    return mask;
}

// One round towards all peers, as three streaming kernels over tiles of
// SCAN_TILE entries (a first version used ONE kernel with a decoupled look-back
// chain per destination; with the 7 neighbours of a brick its latency chains
// made a halo round cost 0.35 ms for 12.5 M cells):
//
//   dd_tile_counts   per tile and destination list: how many of its entries go
//                    there (lists: the peers, plus the ghost entries met when
//                    a migration round walks the cube-ordered plane);
//   dd_tile_offsets  exclusive scan of those counts over the tiles, one CTA per
//                    list; totals -> outbox headers, stayer count, overflow;
//   dd_pack          every tile ranks its entries inside the tile (ballots and
//                    one shared-memory scan) and writes the records to their
//                    final places: a stable compaction, entry order preserved.
//
// Entries are the owned cells in index order, or -- in a migration round with
// `order`, the cube-ordered pos4 plane of the last force evaluation -- every
// slot of that plane: the cells that stay are then re-stored in cube order (see
// slab.cuh). Halo rounds read the one-byte face flags a previous kernel left
// per cell when they are at hand, instead of the positions.
constexpr int DD_LISTS = DD_MAX_PEERS + 1;

// Destinations by face code (bit 2a: below the face of axis a, bit 2a + 1: at or
// above the opposite one): the peers that get a copy (halo round) or the one
// that takes the cell over (migration round). Filled once per CTA, so that the
// per-cell work is one table look-up instead of a loop over the peers.
__device__ __forceinline__ void dd_fill_destinations(
    unsigned* s_dest, const Dd_region& region, bool migration)
{
    const int code = threadIdx.x;
    if (code >= 64) return;
    unsigned mask = 0;
    for (int p = 0; p < region.n_peers; p++) {
        bool takes = true;
        for (int a = 0; a < 3; a++) {
            const int d = region.dir[p][a];
            const bool below = (code >> (2 * a)) & 1, above = (code >> (2 * a + 1)) & 1;
            if (d < 0) takes = takes && below;
            if (d > 0) takes = takes && above;
            if (d == 0 && migration) takes = takes && !below && !above;
        }
        mask |= (takes ? 1u : 0u) << p;
    }
    s_dest[code] = mask;
}

__device__ __forceinline__ unsigned dd_face_code(
    const float* x, const Dd_region& region, bool migration)
{
    const float inset = migration ? 0.f : region.halo;
    unsigned code = 0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        code |= (x[a] < region.lo[a] + inset ? 1u : 0u) << (2 * a);
        code |= (x[a] >= region.hi[a] - inset ? 1u : 0u) << (2 * a + 1);
    }
    return code;
}

// bit p: the entry goes to peer p; bit 31: a ghost entry of the walked plane
template<typename Pt>
__device__ __forceinline__ void dd_classify(int tile, int t,
    const Pt* __restrict__ P, const Dd_region& region, bool migration,
    const float4* __restrict__ order, int n, int n_owned,
    const unsigned char* __restrict__ halo_flags, const unsigned* s_dest,
    unsigned* mask, int* cell)
{
    const bool permute = order != nullptr;
    const int first = tile * SCAN_TILE;
#pragma unroll
    for (int u = 0; u < SELECT_SUB; u++) {
        const int q = first + u * SCAN_THREADS + t;
        mask[u] = 0;
        cell[u] = q;
        if (q >= n) continue;
        bool ghost = false;
        if (permute) {
            cell[u] = __float_as_int(__ldg(&order[q].w));
            ghost = cell[u] >= n_owned;
        }
        if (ghost) {
            mask[u] = 1u << 31;
        } else if (halo_flags != nullptr) {
            const unsigned code = __ldg(halo_flags + q);
            mask[u] = code ? s_dest[code] : 0u;
        } else {
            const float* x = reinterpret_cast<const float*>(P + cell[u]);
            const float pos[3] = {__ldg(x), __ldg(x + 1), __ldg(x + 2)};
            const unsigned code = dd_face_code(pos, region, migration);
            mask[u] = code ? s_dest[code] : 0u;
        }
    }
}

// list index of a mask bit
__device__ __forceinline__ int dd_list_of_bit(int bit, int n_peers)
{
    return bit == 31 ? n_peers : bit;
}

template<typename Pt>
__global__ void __launch_bounds__(SCAN_THREADS) dd_tile_counts(const Step_ctl* ctl,
    const Pt* __restrict__ P, Dd_region region, int migration,
    const float4* __restrict__ order, const int* __restrict__ d_n_total,
    int n_max, const unsigned char* __restrict__ halo_flags,
    int* __restrict__ tile_counts, int n_tiles)
{
    __shared__ int s_total[DD_LISTS];
    __shared__ unsigned s_dest[64];
    const int t = threadIdx.x, lane_id = t & 31;
    const int tile = blockIdx.x;
    const int n_owned = ctl->n_owned;
    const bool permute = order != nullptr;
    const int n = permute ? live_cells(d_n_total, n_max) : n_owned;
    const int n_lists = region.n_peers + (permute ? 1 : 0);
    if (t < DD_LISTS) s_total[t] = 0;
    dd_fill_destinations(s_dest, region, migration != 0);
    __syncthreads();
    if (tile * SCAN_TILE < n) {
        unsigned mask[SELECT_SUB];
        int cell[SELECT_SUB];
        dd_classify(tile, t, P, region, migration != 0, order, n, n_owned,
            halo_flags, s_dest, mask, cell);
        unsigned any = 0;
#pragma unroll
        for (int u = 0; u < SELECT_SUB; u++) any |= mask[u];
        // lists some lane of this warp contributes to (most warps: none)
        unsigned present = __reduce_or_sync(0xffffffffu, any);
        while (present) {
            const int bit = __ffs(present) - 1;
            present &= present - 1;
            int mine = 0;
#pragma unroll
            for (int u = 0; u < SELECT_SUB; u++) mine += (mask[u] >> bit) & 1u;
            mine = __reduce_add_sync(0xffffffffu, mine);
            if (lane_id == 0)
                atomicAdd(&s_total[dd_list_of_bit(bit, region.n_peers)], mine);
        }
    }
    __syncthreads();
    if (t < n_lists) tile_counts[size_t(t) * n_tiles + tile] = s_total[t];
}

// One CTA per list: tile_counts -> exclusive offsets (in place); the totals go
// to the outbox headers (peers) and into totals[] (all lists).
__global__ void __launch_bounds__(1024) dd_tile_offsets(Step_ctl* ctl,
    int* __restrict__ tile_counts, int n_tiles, Dd_outboxes to, int n_peers,
    int* __restrict__ totals)
{
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int l = blockIdx.x, t = threadIdx.x;
    const int lane_id = t & 31, warp_id = t >> 5;
    int* counts = tile_counts + size_t(l) * n_tiles;
    if (t == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += 1024) {
        const int q = base + t;
        const int mine = q < n_tiles ? counts[q] : 0;
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane_id >= d) incl += up;
        }
        if (lane_id == 31) s_warp[warp_id] = incl;
        __syncthreads();
        if (warp_id == 0) {
            int w = s_warp[lane_id];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, w, d);
                if (lane_id >= d) w += up;
            }
            s_warp[lane_id] = w;  // inclusive over the warps
        }
        __syncthreads();
        const int before = s_carry + (warp_id > 0 ? s_warp[warp_id - 1] : 0);
        if (q < n_tiles) counts[q] = before + incl - mine;
        __syncthreads();
        if (t == 1023) s_carry = before + incl;
        __syncthreads();
    }
    if (t == 0) {
        const int total = s_carry;
        totals[l] = total;
        if (l < n_peers) {
            if (total > to.capacity[l]) atomicAdd(&ctl->out_of_grid, 1 << 20);
            to.buffer[l][0] = __int_as_float(min(total, to.capacity[l]));
        }
    }
}

template<typename Pt>
__global__ void __launch_bounds__(SCAN_THREADS) dd_pack(const Step_ctl* ctl,
    const Pt* __restrict__ P, const float3* __restrict__ v, Dd_region region,
    Dd_outboxes to, int migration, Pt* __restrict__ X_stay,
    float3* __restrict__ v_stay, int* n_stay, const float4* __restrict__ order,
    const int* __restrict__ d_n_total, int n_max,
    const unsigned char* __restrict__ halo_flags,
    const int* __restrict__ tile_offsets, const int* __restrict__ totals,
    int n_tiles, Halo_faces faces, unsigned char* __restrict__ flags_stay,
    Dd_extras extras, int record_floats)
{
    constexpr int W_BASE = Layout<Pt>::lanes + 3;
    const int W = record_floats;
    constexpr int WARPS = SCAN_THREADS / 32;
    __shared__ unsigned short s_count[DD_LISTS][SELECT_SUB][WARPS];  // -> prefixes
    __shared__ int s_tile_prefix[DD_LISTS];
    __shared__ unsigned s_dest[64];
    __shared__ float* s_buffer[DD_MAX_PEERS];
    __shared__ int s_capacity[DD_MAX_PEERS];
    __shared__ int s_busy;

    const int t = threadIdx.x;
    const int lane_id = t & 31, warp_id = t >> 5;
    const int tile = blockIdx.x;
    const int n_owned = ctl->n_owned;
    const bool permute = order != nullptr;
    const int n = permute ? live_cells(d_n_total, n_max) : n_owned;
    const int first = tile * SCAN_TILE;
    const int n_peers = region.n_peers;
    const int n_lists = n_peers + (permute ? 1 : 0);

    if (tile == 0 && t == 0 && migration) {
        int leaving = 0;
        for (int p = 0; p < n_peers; p++) leaving += totals[p];
        *n_stay = n_owned - leaving;
    }
    if (first >= n) return;
    // what lies in front of this tile in every list; a halo tile whose own
    // counts are all zero has nothing to do
    if (t == 0) s_busy = migration;
    __syncthreads();
    if (t < n_lists) {
        const int here = tile_offsets[size_t(t) * n_tiles + tile];
        s_tile_prefix[t] = here;
        const int next = tile + 1 < n_tiles
                             ? tile_offsets[size_t(t) * n_tiles + tile + 1]
                             : totals[t];
        if (next != here) s_busy = 1;
    }
    if (t < n_peers) {
        s_buffer[t] = to.buffer[t];
        s_capacity[t] = to.capacity[t];
    }
    dd_fill_destinations(s_dest, region, migration != 0);
    __syncthreads();
    if (!s_busy) return;

    unsigned mask[SELECT_SUB];
    int cell[SELECT_SUB];
    dd_classify(tile, t, P, region, migration != 0, order, n, n_owned, halo_flags,
        s_dest, mask, cell);
    // per (sub-block, warp): how many entries go to every list -- lane l keeps
    // the count of list l; only the lists present in the warp are voted on
#pragma unroll
    for (int u = 0; u < SELECT_SUB; u++) {
        unsigned present = __reduce_or_sync(0xffffffffu, mask[u]);
        int my_count = 0;
        while (present) {
            const int bit = __ffs(present) - 1;
            present &= present - 1;
            const unsigned votes = __ballot_sync(0xffffffffu, (mask[u] >> bit) & 1u);
            if (lane_id == dd_list_of_bit(bit, n_peers)) my_count = __popc(votes);
        }
        if (lane_id < n_lists) s_count[lane_id][u][warp_id] = my_count;
    }
    __syncthreads();
    // exclusive scan of the SUB x WARPS counts of every list, one warp per list
    for (int l = warp_id; l < n_lists; l += WARPS) {
        constexpr int ENTRIES = SELECT_SUB * WARPS, PER_LANE = ENTRIES / 32;
        unsigned short* counts = &s_count[l][0][0];
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
    }
    __syncthreads();

    const unsigned below = (1u << lane_id) - 1u;
#pragma unroll
    for (int u = 0; u < SELECT_SUB; u++) {
        const int q = first + u * SCAN_THREADS + t;
        const int i = cell[u];
        // entries in front of this warp's row that leave (all peers) / are ghosts
        int before = 0;
        if (migration) {
            before = lane_id < n_lists
                         ? s_tile_prefix[lane_id] + s_count[lane_id][u][warp_id]
                         : 0;
            before = __reduce_add_sync(0xffffffffu, before);
        }
        unsigned present = __reduce_or_sync(0xffffffffu, mask[u]);
        while (present) {
            const int bit = __ffs(present) - 1;
            present &= present - 1;
            const unsigned mine = (mask[u] >> bit) & 1u;
            const unsigned votes = __ballot_sync(0xffffffffu, mine);
            const int rank = __popc(votes & below);
            before += rank;  // (a migrating entry has exactly one bit)
            if (bit < 31 && mine) {
                const int at = s_tile_prefix[bit] + s_count[bit][u][warp_id] + rank;
                if (at < s_capacity[bit]) {
                    float* record = s_buffer[bit] + SLAB_HEADER + size_t(at) * W;
                    write_record(record, P, v, i);
                    if (extras.count > 0)
                        write_extras(record + W_BASE, extras, i, migration != 0);
                }
            }
        }
        if (migration && q < n && mask[u] == 0) {
            // a cell that stays: to its place among the stayers, with the face
            // flags the first halo round of the next step will read
            const int at = q - before;
            const Pt X = load_pt(P, i);
            store_pt(X_stay, at, X);
            v_stay[at] = v[i];
            flags_stay[at] = halo_flags_of(X.x, X.y, X.z, faces);
            for (int k = 0; k < extras.count; k++) {  // re-stored via scratch
                const int words = extras.words[k];
                const unsigned* src = extras.data[k] + size_t(i) * words;
                unsigned* dst = extras.scratch[k] + size_t(at) * words;
                for (int w = 0; w < words; w++) dst[w] = src[w];
            }
        }
    }
}

// Stream every local outbox (header + the records it holds) into the inbox of
// its neighbour: a few dozen CTAs per peer, 16-byte loads and stores. After a
// system-scope fence the last CTA to finish stores the epoch into the peers'
// flag words.
inline int dd_push_ctas()
{
    static const int ctas = [] {
        const char* env = getenv("YALLA_B200_PUSH_CTAS");
        const int wanted = env && env[0] ? atoi(env) : 0;
        return wanted > 0 ? wanted : 24;
    }();
    return ctas;
}

__global__ void __launch_bounds__(256) dd_push(Dd_outboxes from, Dd_outboxes to,
    int n_peers, int ctas_per_peer, int record_floats, int* push_done,
    unsigned epoch)
{
    const int DD_PUSH_CTAS = ctas_per_peer;
    const int p = blockIdx.x / DD_PUSH_CTAS, part = blockIdx.x % DD_PUSH_CTAS;
    if (p < n_peers) {
        const int count = __float_as_int(from.buffer[p][0]);
        // whole float4s, header included (buffers are padded to 256 bytes)
        const int n_vec = (SLAB_HEADER + count * record_floats + 3) / 4;
        const float4* src = reinterpret_cast<const float4*>(from.buffer[p]);
        float4* dst = reinterpret_cast<float4*>(to.buffer[p]);
        for (int q = part * blockDim.x + threadIdx.x; q < n_vec;
             q += DD_PUSH_CTAS * blockDim.x)
            dst[q] = src[q];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(push_done, 1) == int(gridDim.x) - 1) {
            *push_done = 0;
            __threadfence_system();
            for (int q = 0; q < n_peers; q++) store_release_sys(to.flag[q], epoch);
        }
    }
}

// Holds the stream until every inbox of the round shows the epoch.
__global__ void dd_wait(Step_ctl* ctl, Dd_inboxes in, unsigned epoch)
{
    const int p = threadIdx.x;
    if (p < in.n_peers && !wait_for_epoch(in.flag[p], epoch))
        atomicAdd(&ctl->out_of_grid, 1 << 24);
}

// Number of records in every inbox and where each starts in the concatenation;
// room for at most `room` records in total.
__device__ __forceinline__ int dd_inbox_layout(
    const Dd_inboxes& in, int room, int* start)
{
    int total = 0;
    for (int p = 0; p < in.n_peers; p++) {
        start[p] = total;
        int count = __float_as_int(in.buffer[p][0]);
        count = max(0, min(count, room - total));
        total += count;
    }
    start[in.n_peers] = total;
    return total;
}

// Ghosts: append the received records behind the owned cells of P (X or X1)
// and set the total cell count. Order: by peer, then as ranked by the sender.
template<typename Pt>
__global__ void __launch_bounds__(256) dd_append_ghosts(Step_ctl* ctl, Pt* P,
    float3* v, Dd_inboxes in, int n_max, int* d_n, Dd_extras extras,
    int record_floats)
{
    constexpr int W_BASE = Layout<Pt>::lanes + 3;
    const int W = record_floats;
    __shared__ int s_start[DD_MAX_PEERS + 1];
    const int n = ctl->n_owned;
    if (threadIdx.x == 0) dd_inbox_layout(in, n_max - n, s_start);
    __syncthreads();
    const int total = s_start[in.n_peers];
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < total;
         r += gridDim.x * blockDim.x) {
        int p = 0;
        while (r >= s_start[p + 1]) p++;
        const float* record = in.buffer[p] + SLAB_HEADER + size_t(r - s_start[p]) * W;
        read_record(record, P, v, n + r);
        if (extras.count > 0) read_extras(record + W_BASE, extras, n + r, false);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *d_n = n + total;
        ctl->n_ghosts = total;
    }
}

// Migration, step 2: the arrivals, peer by peer, behind the cells that stayed
// (dd_pack has put those in place already).
template<typename Pt>
__global__ void __launch_bounds__(256) dd_merge(const Step_ctl* ctl,
    const int* __restrict__ n_stay_in, Dd_inboxes in, int n_max, Pt* X, float3* v,
    int* new_count, Halo_faces faces, unsigned char* __restrict__ halo_flags,
    Dd_extras extras, int record_floats)
{
    constexpr int W_BASE = Layout<Pt>::lanes + 3;
    const int W = record_floats;
    __shared__ int s_start[DD_MAX_PEERS + 1];
    const int n_stay = *n_stay_in;
    if (threadIdx.x == 0) dd_inbox_layout(in, n_max - n_stay, s_start);
    __syncthreads();
    const int arrivals = s_start[in.n_peers];
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < arrivals;
         a += gridDim.x * blockDim.x) {
        int p = 0;
        while (a >= s_start[p + 1]) p++;
        const int r = n_stay + a;
        const float* record = in.buffer[p] + SLAB_HEADER + size_t(a - s_start[p]) * W;
        read_record(record, X, v, r);
        if (extras.count > 0) read_extras(record + W_BASE, extras, r, true);
        const float* x = reinterpret_cast<const float*>(X + r);
        halo_flags[r] = halo_flags_of(x[0], x[1], x[2], faces);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *new_count = n_stay + arrivals;
}

// The registered arrays of the cells that stayed: scratch -> the model's arrays
// (dd_pack could not re-store them in place).
__global__ void __launch_bounds__(256) dd_restore_extras(
    const int* __restrict__ n_stay_in, Dd_extras extras)
{
    const int n_stay = *n_stay_in;
    for (int k = 0; k < extras.count; k++) {
        const long long words = static_cast<long long>(n_stay) * extras.words[k];
        for (long long w = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
             w < words; w += static_cast<long long>(gridDim.x) * blockDim.x)
            extras.data[k][w] = extras.scratch[k][w];
    }
}

// Between two steps the model may have appended cells (division) and bumped
// *d_n: flag the new ones, then adopt the count as this brick's owned cells.
template<typename Pt>
__global__ void __launch_bounds__(256) dd_flag_new_cells(const Step_ctl* ctl,
    const int* __restrict__ d_n, int n_max, const Pt* __restrict__ X,
    Halo_faces faces, unsigned char* __restrict__ halo_flags)
{
    const int n = live_cells(d_n, n_max);
    for (int i = ctl->n_owned + blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += gridDim.x * blockDim.x) {
        const float* x = reinterpret_cast<const float*>(X + i);
        halo_flags[i] = halo_flags_of(x[0], x[1], x[2], faces);
    }
}

__global__ void dd_adopt_owned(Step_ctl* ctl, const int* d_n, int n_max)
{
    ctl->n_owned = live_cells(d_n, n_max);
}

// End of a survey round: the ghosts are dropped again.
__global__ void dd_drop_ghosts(Step_ctl* ctl, int* d_n)
{
    *d_n = ctl->n_owned;
    ctl->n_ghosts = 0;
}

// Global drift of a stage: publish {sum dX, n} of the owned cells to every
// rank, wait for everybody's, add in rank order, divide like operator/= would
// (dtypes.cuh:204-208). One warp per 32 ranks; launched <<<1, DD_MAX_RANKS>>>.
__global__ void dd_allreduce_drift(Step_ctl* ctl, int stage, Dd_mailboxes boxes,
    const Dd_mailbox* mine, int rank, int world, int parity, unsigned epoch)
{
    __shared__ float s_sums[DD_MAX_RANKS][4];
    const int r = threadIdx.x;
    if (r < world) {
        Dd_mailbox* theirs = boxes.of_rank[r];
        float4 sums;
        sums.x = ctl->drift_sum[stage][0], sums.y = ctl->drift_sum[stage][1];
        sums.z = ctl->drift_sum[stage][2], sums.w = ctl->drift_sum[stage][3];
        *reinterpret_cast<float4*>(theirs->sums[parity][rank]) = sums;
        __threadfence_system();
        store_release_sys(&theirs->flag[parity][rank], epoch);
        if (!wait_for_epoch(&mine->flag[parity][r], epoch))
            atomicAdd(&ctl->out_of_grid, 1 << 24);
        const volatile float* got = mine->sums[parity][r];
        s_sums[r][0] = got[0], s_sums[r][1] = got[1];
        s_sums[r][2] = got[2], s_sums[r][3] = got[3];
    }
    __syncthreads();
    if (r == 0) {
        float total[4] = {0.f, 0.f, 0.f, 0.f};
        for (int q = 0; q < world; q++)
            for (int c = 0; c < 4; c++) total[c] += s_sums[q][c];
        const float inv_n = static_cast<float>(1. / total[3]);
        ctl->drift[stage][0] = total[0] * inv_n;
        ctl->drift[stage][1] = total[1] * inv_n;
        ctl->drift[stage][2] = total[2] * inv_n;
    }
}

// ---- a seeded tissue generated in place ---------------------------------------
// The cells of a jittered FCC ball (nearest-neighbour distance d, radius R)
// that fall into this brick, written behind *d_count. Every lattice site gets
// its jitter from a hash of (site, seed), so the tissue is the same however it
// is cut; only the order of the cells in memory depends on the cut and on
// atomic timing (a decomposed run re-stores its cells in cube order anyway).
__device__ __forceinline__ unsigned long long dd_mix(unsigned long long z)
{
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

template<typename Pt>
__global__ void __launch_bounds__(256) dd_seed_lattice_ball(float radius, float d,
    float jitter, unsigned long long seed, Dd_region region, int half,
    long long n_sites, int n_max, Pt* X, float3* v, int* d_count)
{
    const float a = d * 1.41421356237f;  // conventional FCC cell edge
    const int side = 2 * half + 1;
    for (long long s = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
         s < n_sites; s += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int basis = static_cast<int>(s & 3);
        long long c = s >> 2;
        const int ix = static_cast<int>(c % side) - half;
        c /= side;
        const int iy = static_cast<int>(c % side) - half;
        const int iz = static_cast<int>(c / side) - half;
        float x = ix * a, y = iy * a, z = iz * a;
        if (basis == 1) x += 0.5f * a, y += 0.5f * a;
        if (basis == 2) x += 0.5f * a, z += 0.5f * a;
        if (basis == 3) y += 0.5f * a, z += 0.5f * a;
        if (x * x + y * y + z * z > radius * radius) continue;
        const unsigned long long h1 = dd_mix(seed ^ (static_cast<unsigned long long>(s) * 3 + 1));
        const unsigned long long h2 = dd_mix(h1);
        const float scale = 2.f * jitter * d;
        x += ((h1 & 0xffffff) * (1.f / 16777216.f) - 0.5f) * scale;
        y += (((h1 >> 24) & 0xffffff) * (1.f / 16777216.f) - 0.5f) * scale;
        z += ((h2 & 0xffffff) * (1.f / 16777216.f) - 0.5f) * scale;
        const float pos[3] = {x, y, z};
        bool mine = true;
#pragma unroll
        for (int q = 0; q < 3; q++)
            mine = mine && pos[q] >= region.lo[q] && pos[q] < region.hi[q];
        if (!mine) continue;
        const int slot = atomicAdd(d_count, 1);
        if (slot >= n_max) continue;  // reported through the count
        Pt cell{0};
        cell.x = x, cell.y = y, cell.z = z;
        store_pt(X, slot, cell);
        v[slot] = float3{0.f, 0.f, 0.f};
    }
}

// ---- host side: the exchange allocation and its layout -------------------------
struct Domain_link {
    bool active = false;
    int rank = 0, world = 1;
    Dd_region region{};
    int peer_rank[DD_MAX_PEERS] = {};
    int peer_direction[DD_MAX_PEERS] = {};  // direction index 0..26 of peer p
    int capacity[DD_MAX_PEERS] = {};
    int record_floats = 0;        // of a migration record (the widest)
    int halo_record_floats = 0;   // of a ghost record
    Dd_extras extras{};           // registered before begin()

    unsigned char* base = nullptr;  // this rank's exchange allocation
    size_t bytes = 0;
    size_t inbox_offset[DD_MAX_PEERS][DD_ROUNDS] = {};
    size_t flag_offset[DD_MAX_PEERS][DD_ROUNDS] = {};

    Dd_outboxes out[DD_ROUNDS] = {};  // the peers' inboxes, filled by connect()
    Dd_outboxes local_out{};           // where dd_select packs, per peer
    unsigned char* local_base = nullptr;
    int* push_done = nullptr;
    Dd_mailboxes mailboxes{};
    unsigned epoch[DD_ROUNDS] = {};
    unsigned drift_epoch = 0;

    // scratch of the compaction: per list and tile, counts then offsets
    int n_tiles = 0;
    int* tile_counts = nullptr;
    int* totals = nullptr;
    int* n_stay = nullptr;
    int* new_count = nullptr;
    unsigned char* halo_flags = nullptr;  // per owned cell, see Halo_faces
    float3* v_new = nullptr;  // the corrector's velocities, until re-stored
    bool flags_valid = false;             // false until a kernel has written them
    bool permute = true;
    // the model appends cells between steps (division): every step starts by
    // adopting them (set with the first registered array; may be set by hand)
    bool grows = false;
    // halo rounds: the push runs on its own stream while the brick's stream
    // starts the grid build with the cells it owns (YALLA_B200_DD_OVERLAP=0:
    // everything on one stream, as in round 1)
    bool overlap = true;
    cudaStream_t push_stream = nullptr;
    cudaEvent_t packed = nullptr, pushed = nullptr;
    bool push_pending = false;

    Halo_faces inset_faces() const
    {
        Halo_faces faces;
        for (int a = 0; a < 3; a++) {
            faces.lo[a] = region.lo[a] + region.halo;
            faces.hi[a] = region.hi[a] - region.halo;
        }
        return faces;
    }

    static size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

    // A per-cell array of the model that has to travel with the cells; call
    // before begin(). bytes_per_cell: a multiple of 4.
    bool register_extra(void* d_array, int bytes_per_cell, bool ghosts_too)
    {
        if (active || extras.count >= DD_MAX_EXTRAS || bytes_per_cell % 4 != 0)
            return false;
        const int k = extras.count++;
        extras.data[k] = static_cast<unsigned*>(d_array);
        extras.words[k] = bytes_per_cell / 4;
        extras.ghosts_too[k] = ghosts_too ? 1 : 0;
        extras.scratch[k] = nullptr;
        grows = true;
        return true;
    }

    void begin(int rank_, int world_, const Dd_region& region_,
        const int* peer_ranks27, const int* capacity27, int base_record_floats,
        int n_max)
    {
        const Dd_extras registered = extras;
        release();
        extras = registered;
        rank = rank_, world = world_, region = region_;
        // records: the cell, its old velocity, then the registered arrays; an
        // even number of floats keeps them 8-byte aligned
        int all_words = 0, ghost_words = 0;
        for (int k = 0; k < extras.count; k++) {
            all_words += extras.words[k];
            if (extras.ghosts_too[k]) ghost_words += extras.words[k];
            YB_CUDA(cudaMalloc(&extras.scratch[k],
                sizeof(unsigned) * size_t(extras.words[k]) * (n_max > 0 ? n_max : 1)));
        }
        record_floats = (base_record_floats + all_words + 1) & ~1;
        halo_record_floats = (base_record_floats + ghost_words + 1) & ~1;
        if (extras.count == 0)
            record_floats = halo_record_floats = base_record_floats;
        region.n_peers = 0;
        for (int dir = 0; dir < 27; dir++) {
            if (dir == 13 || peer_ranks27[dir] < 0) continue;
            const int p = region.n_peers++;
            peer_rank[p] = peer_ranks27[dir];
            peer_direction[p] = dir;
            capacity[p] = capacity27[dir];
            region.dir[p][0] = static_cast<signed char>(dir % 3 - 1);
            region.dir[p][1] = static_cast<signed char>((dir / 3) % 3 - 1);
            region.dir[p][2] = static_cast<signed char>(dir / 9 - 1);
            region.dir[p][3] = 0;
        }
        size_t at = align_up(sizeof(Dd_mailbox));
        for (int p = 0; p < region.n_peers; p++)
            for (int q = 0; q < DD_ROUNDS; q++) {
                flag_offset[p][q] = at;
                at += sizeof(unsigned);
            }
        at = align_up(at);
        for (int p = 0; p < region.n_peers; p++)
            for (int q = 0; q < DD_ROUNDS; q++) {
                inbox_offset[p][q] = at;
                at = align_up(at + sizeof(float) * (SLAB_HEADER +
                                        size_t(capacity[p]) * record_floats));
            }
        bytes = at;
        YB_CUDA(cudaMalloc(&base, bytes));
        YB_CUDA(cudaMemset(base, 0, bytes));
        // local outboxes, one per peer (a round is pushed before the next packs)
        size_t local_bytes = 0;
        size_t local_offset[DD_MAX_PEERS];
        for (int p = 0; p < region.n_peers; p++) {
            local_offset[p] = local_bytes;
            local_bytes = align_up(local_bytes + sizeof(float) * (SLAB_HEADER +
                                       size_t(capacity[p]) * record_floats));
        }
        YB_CUDA(cudaMalloc(&local_base, local_bytes > 0 ? local_bytes : 256));
        YB_CUDA(cudaMemset(local_base, 0, local_bytes > 0 ? local_bytes : 256));
        local_out = Dd_outboxes{};
        for (int p = 0; p < region.n_peers; p++) {
            local_out.buffer[p] = reinterpret_cast<float*>(local_base + local_offset[p]);
            local_out.capacity[p] = capacity[p];
        }
        YB_CUDA(cudaMalloc(&push_done, sizeof(int)));
        YB_CUDA(cudaMemset(push_done, 0, sizeof(int)));
        for (int r = 0; r < DD_MAX_RANKS; r++) mailboxes.of_rank[r] = nullptr;
        mailboxes.of_rank[rank] = reinterpret_cast<Dd_mailbox*>(base);

        n_tiles = ceil_div(n_max > 0 ? n_max : 1, SCAN_TILE);
        const size_t words = size_t(DD_LISTS) * n_tiles;
        YB_CUDA(cudaMalloc(&tile_counts, words * sizeof(int)));
        YB_CUDA(cudaMemset(tile_counts, 0, words * sizeof(int)));
        YB_CUDA(cudaMalloc(&totals, DD_LISTS * sizeof(int)));
        YB_CUDA(cudaMemset(totals, 0, DD_LISTS * sizeof(int)));
        YB_CUDA(cudaMalloc(&n_stay, sizeof(int)));
        YB_CUDA(cudaMalloc(&new_count, sizeof(int)));
        YB_CUDA(cudaMalloc(&halo_flags, n_max > 0 ? n_max : 1));
        YB_CUDA(cudaMalloc(&v_new, sizeof(float3) * size_t(n_max > 0 ? n_max : 1)));
        flags_valid = false;
        YB_CUDA(cudaStreamCreateWithFlags(&push_stream, cudaStreamNonBlocking));
        YB_CUDA(cudaEventCreateWithFlags(&packed, cudaEventDisableTiming));
        YB_CUDA(cudaEventCreateWithFlags(&pushed, cudaEventDisableTiming));
        push_pending = false;
        const char* overlap_env = getenv("YALLA_B200_DD_OVERLAP");
        overlap = !(overlap_env && overlap_env[0] == '0');
        const char* env = getenv("YALLA_B200_SLAB_PERMUTE");
        permute = !(env && env[0] == '0');
        for (int q = 0; q < DD_ROUNDS; q++) epoch[q] = 0;
        drift_epoch = 0;
        active = true;
    }

    int peer_of_direction(int dir) const
    {
        for (int p = 0; p < region.n_peers; p++)
            if (peer_direction[p] == dir) return p;
        return -1;
    }

    // The peer in direction `dir` keeps my records in ITS inboxes for the
    // opposite direction: peer_base is its exchange allocation as mapped here,
    // offsets = its {inbox_offset[q]..., flag_offset[q]...} (2 * DD_ROUNDS
    // entries) for direction 26 - dir.
    bool connect(int dir, void* peer_base, const long long* offsets)
    {
        const int p = peer_of_direction(dir);
        if (p < 0) return false;
        unsigned char* theirs = static_cast<unsigned char*>(peer_base);
        for (int q = 0; q < DD_ROUNDS; q++) {
            out[q].buffer[p] = reinterpret_cast<float*>(theirs + offsets[q]);
            out[q].flag[p] =
                reinterpret_cast<unsigned*>(theirs + offsets[DD_ROUNDS + q]);
            out[q].capacity[p] = capacity[p];
        }
        return true;
    }
    void connect_mailbox(int r, void* peer_base)
    {
        mailboxes.of_rank[r] = static_cast<Dd_mailbox*>(peer_base);
    }
    Dd_inboxes inboxes(int q) const
    {
        Dd_inboxes in{};
        in.n_peers = region.n_peers;
        for (int p = 0; p < region.n_peers; p++) {
            in.buffer[p] = reinterpret_cast<const float*>(base + inbox_offset[p][q]);
            in.flag[p] = reinterpret_cast<const unsigned*>(base + flag_offset[p][q]);
        }
        return in;
    }
    const Dd_mailbox* my_mailbox() const
    {
        return reinterpret_cast<const Dd_mailbox*>(base);
    }
    bool connected() const
    {
        for (int p = 0; p < region.n_peers; p++)
            if (out[0].buffer[p] == nullptr) return false;
        for (int r = 0; r < world; r++)
            if (mailboxes.of_rank[r] == nullptr) return false;
        return true;
    }

    void release()
    {
        if (!active) return;
        cudaStreamSynchronize(push_stream);
        cudaStreamDestroy(push_stream);
        cudaEventDestroy(packed);
        cudaEventDestroy(pushed);
        push_stream = nullptr;
        for (int k = 0; k < extras.count; k++) cudaFree(extras.scratch[k]);
        extras = Dd_extras{};
        cudaFree(v_new);
        cudaFree(halo_flags);
        cudaFree(new_count);
        cudaFree(n_stay);
        cudaFree(totals);
        cudaFree(tile_counts);
        cudaFree(push_done);
        cudaFree(local_base);
        cudaFree(base);
        base = nullptr;
        active = false;
        for (int q = 0; q < DD_ROUNDS; q++) out[q] = Dd_outboxes{};
    }
};

}  // namespace yb
