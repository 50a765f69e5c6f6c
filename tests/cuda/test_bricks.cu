// GPU test program: a decomposed tissue driven from CUDA C++ through the header
// API alone (no Python, no C ABI, no communication library). Two bricks of one
// float3 tissue live in this process, each a Solution on its own stream; their
// exchange allocations are connected with plain device pointers (across
// processes the same calls take addresses mapped with CUDA IPC). Every brick
// registers an identity array that travels with its cells. The decomposed run
// must reproduce the single-domain run cell by cell (matched by identity).
// Prints "ok <name>" per check, exits non-zero on the first failure
// (run by tests/test_extensions_gpu.py).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../include/dtypes.cuh"
#include "../../include/inits.cuh"
#include "../../include/solvers.cuh"

#define CHECK(cond, name)                                   \
    do {                                                    \
        if (!(cond)) {                                      \
            printf("FAILED %s (%s:%d)\n", name, __FILE__, __LINE__); \
            exit(1);                                        \
        }                                                   \
        printf("ok %s\n", name);                            \
    } while (0)

using Brick = Solution<float3, Grid_solver>;

// the neighbour of `a` in direction `dir` is `b`: a's records for b go to the
// inboxes b keeps for the opposite direction
static void connect(Brick& a, int dir, Brick& b)
{
    const int p = b.dom.peer_of_direction(26 - dir);
    long long offsets[2 * yb::DD_ROUNDS];
    for (int q = 0; q < yb::DD_ROUNDS; q++) {
        offsets[q] = static_cast<long long>(b.dom.inbox_offset[p][q]);
        offsets[yb::DD_ROUNDS + q] = static_cast<long long>(b.dom.flag_offset[p][q]);
    }
    if (p < 0 || !a.dom.connect(dir, b.dom.base, offsets)) {
        printf("FAILED connect\n");
        exit(1);
    }
}

int main()
{
    // a jittered cubic lattice ball, squeezed so that it expands across the cut
    const float spacing = 0.66f, cut = 2.f;
    const int half = 15, grid_size = 40;
    std::vector<float3> X;
    unsigned long long state = 12345;
    auto uniform = [&state]() {
        state = state * 6364136223846793005ull + 1442695040888963407ull;
        return static_cast<float>((state >> 40) & 0xffffff) / 16777216.f;
    };
    for (int ix = -half; ix <= half; ix++)
        for (int iy = -half; iy <= half; iy++)
            for (int iz = -half; iz <= half; iz++) {
                const float x = ix * spacing + (uniform() - 0.5f) * 0.04f;
                const float y = iy * spacing + (uniform() - 0.5f) * 0.04f;
                const float z = iz * spacing + (uniform() - 0.5f) * 0.04f;
                if (x * x + y * y + z * z < 9.f * 9.f) X.push_back(float3{x, y, z});
            }
    const int n = static_cast<int>(X.size());
    const int steps = 8;
    const float dt = 0.1f;

    // ---- one domain ---------------------------------------------------------
    std::vector<float3> want(n);
    {
        Brick cells{n, grid_size, 1.f};
        for (int i = 0; i < n; i++) cells.h_X[i] = X[i];
        *cells.h_n = n;
        cells.copy_to_device();
        for (int s = 0; s < steps; s++) cells.take_step<relu_force>(dt);
        cells.copy_to_host();
        for (int i = 0; i < n; i++) want[i] = cells.h_X[i];
    }

    // ---- two bricks, cut at z = 2 ---------------------------------------------
    Brick lower{n, grid_size, 1.f}, upper{n, grid_size, 1.f};
    Brick* bricks[2] = {&lower, &upper};
    cudaStream_t streams[2];
    int* d_identity[2];
    const int up = yb::dd_direction_index(0, 0, 1), down = yb::dd_direction_index(0, 0, -1);
    for (int rank = 0; rank < 2; rank++) {
        cudaStreamCreateWithFlags(&streams[rank], cudaStreamNonBlocking);
        bricks[rank]->stream = streams[rank];
        cudaMalloc(&d_identity[rank], sizeof(int) * n);
        cudaMemset(d_identity[rank], 0xff, sizeof(int) * n);
        CHECK(bricks[rank]->dom_register_array(d_identity[rank], sizeof(int), true),
            "register an identity array");
        int peers[27], capacity[27];
        for (int d = 0; d < 27; d++) peers[d] = -1, capacity[d] = 0;
        peers[rank == 0 ? up : down] = 1 - rank;
        capacity[rank == 0 ? up : down] = n;
        const float lo[3] = {-INFINITY, -INFINITY, rank == 0 ? -INFINITY : cut};
        const float hi[3] = {INFINITY, INFINITY, rank == 0 ? cut : INFINITY};
        bricks[rank]->dom_begin(rank, 2, lo, hi, 1.5f, peers, capacity);
    }
    CHECK(!lower.dom_register_array(d_identity[0], sizeof(int), true),
        "no registration after dom_begin");
    connect(lower, up, upper);
    connect(upper, down, lower);
    for (int rank = 0; rank < 2; rank++)
        for (int other = 0; other < 2; other++)
            bricks[rank]->dom.connect_mailbox(other, bricks[other]->dom.base);
    CHECK(lower.dom.connected() && upper.dom.connected(), "bricks connected");

    int before[2];
    for (int rank = 0; rank < 2; rank++) {
        std::vector<float3> mine;
        std::vector<int> ids;
        for (int i = 0; i < n; i++)
            if ((X[i].z >= cut) == (rank == 1)) {
                mine.push_back(X[i]);
                ids.push_back(i);
            }
        before[rank] = static_cast<int>(mine.size());
        float3 *d_X, *d_v;
        cudaMalloc(&d_X, sizeof(float3) * mine.size());
        cudaMalloc(&d_v, sizeof(float3) * mine.size());
        cudaMemcpy(d_X, mine.data(), sizeof(float3) * mine.size(), cudaMemcpyHostToDevice);
        cudaMemset(d_v, 0, sizeof(float3) * mine.size());
        cudaMemcpy(d_identity[rank], ids.data(), sizeof(int) * ids.size(),
            cudaMemcpyHostToDevice);
        bricks[rank]->slab_set_owned(d_X, d_v, before[rank]);
        cudaStreamSynchronize(streams[rank]);
        cudaFree(d_X);
        cudaFree(d_v);
    }
    CHECK(before[0] > 0 && before[1] > 0 && before[0] + before[1] == n,
        "both bricks own cells");

    // a step is kernels only: the host just keeps both streams fed
    for (int s = 0; s < steps; s++)
        for (int rank = 0; rank < 2; rank++)
            bricks[rank]->dom_step<relu_force, friction_w_neighbour>(dt);
    cudaDeviceSynchronize();
    CHECK(cudaGetLastError() == cudaSuccess, "decomposed steps ran");

    int after[2], ghosts = 0;
    float worst = 0.f;
    std::vector<char> seen(n, 0);
    for (int rank = 0; rank < 2; rank++) {
        int n_total = 0, problems = 0;
        bricks[rank]->slab_counts(&after[rank], &n_total, &problems);
        CHECK(problems == 0, "no overflow, nobody gave up waiting");
        ghosts += n_total - after[rank];
        std::vector<float3> got(after[rank]);
        std::vector<int> ids(after[rank]);
        cudaMemcpy(got.data(), bricks[rank]->d_X, sizeof(float3) * after[rank],
            cudaMemcpyDeviceToHost);
        cudaMemcpy(ids.data(), d_identity[rank], sizeof(int) * after[rank],
            cudaMemcpyDeviceToHost);
        for (int k = 0; k < after[rank]; k++) {
            const int id = ids[k];
            if (id < 0 || id >= n || seen[id]) {
                printf("FAILED identity %d of brick %d\n", id, rank);
                exit(1);
            }
            seen[id] = 1;
            worst = fmaxf(worst, fabsf(got[k].x - want[id].x));
            worst = fmaxf(worst, fabsf(got[k].y - want[id].y));
            worst = fmaxf(worst, fabsf(got[k].z - want[id].z));
        }
    }
    CHECK(after[0] + after[1] == n, "no cell lost or duplicated");
    CHECK(after[0] != before[0], "cells migrated across the cut");
    CHECK(ghosts > 0, "ghosts were exchanged");
    printf("max deviation from one domain: %.3e (%d cells, %d migrated)\n", worst, n,
        abs(after[0] - before[0]));
    CHECK(worst < 2e-5f * steps * 10.f, "decomposed run equals the single-domain run");

    for (int rank = 0; rank < 2; rank++) cudaFree(d_identity[rank]);
    printf("all brick checks passed\n");
    return 0;
}
