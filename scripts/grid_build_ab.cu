// A/B of the neighbour-grid build: this repo's bucket sort (include/solvers.cuh,
// Grid::build: bin_cells, scan_bins, place_ids, publish_grid) against the
// reference's compute_cube_id + thrust::sort_by_key (CUB onesweep radix sort
// over 32 key bits) + fills + boundary kernel (reference solvers.cuh:380-425).
// One source, compiled against either header set (oracle/build_checkers.py):
//     tests/_bin/grid_ab_product    -I include
//     tests/_bin/grid_ab_reference  -I /root/reference/include
// Prints one JSON line per size: microseconds per Grid::build, CUDA events
// around 20 builds after 3 warm-up builds, plus a checksum of the four arrays
// (the two binaries must print the same one: the result is bit-identical).
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "dtypes.cuh"
#include "solvers.cuh"

static unsigned long long checksum(const int* d, size_t n)
{
    std::vector<int> h(n);
    cudaMemcpy(h.data(), d, n * sizeof(int), cudaMemcpyDeviceToHost);
    unsigned long long sum = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) sum = (sum ^ (unsigned)h[i]) * 1099511628211ull;
    return sum;
}

int main(int argc, char** argv)
{
#ifdef YALLA_B200
    const char* impl = "product";
#else
    const char* impl = "reference";
#endif
    const int sizes[][2] = {{1000000, 112}, {10000000, 240}};
    for (auto& size : sizes) {
        const int n = size[0], gs = size[1];
        // uniform ball at the density of a relaxed tissue (d = 0.8)
        const float radius = cbrtf(n / 0.64f) * 0.8f / 2;
        std::vector<float3> h_X(n);
        srand(7);
        for (int i = 0; i < n;) {
            const float x = (rand() / (RAND_MAX + 1.f) * 2 - 1) * radius;
            const float y = (rand() / (RAND_MAX + 1.f) * 2 - 1) * radius;
            const float z = (rand() / (RAND_MAX + 1.f) * 2 - 1) * radius;
            if (x * x + y * y + z * z > radius * radius) continue;
            h_X[i++] = float3{x, y, z};
        }
        float3* d_X;
        cudaMalloc(&d_X, n * sizeof(float3));
        cudaMemcpy(d_X, h_X.data(), n * sizeof(float3), cudaMemcpyHostToDevice);
        Grid grid{n, gs};
        for (int k = 0; k < 3; k++) grid.build(n, d_X, 1.f);
        cudaDeviceSynchronize();
        cudaEvent_t start, stop;
        cudaEventCreate(&start);
        cudaEventCreate(&stop);
        const int repeats = 20;
        cudaEventRecord(start);
        for (int k = 0; k < repeats; k++) grid.build(n, d_X, 1.f);
        cudaEventRecord(stop);
        cudaEventSynchronize(stop);
        float ms = 0;
        cudaEventElapsedTime(&ms, start, stop);
        const unsigned long long sum =
            checksum(grid.d_cube_id, n) ^ checksum(grid.d_point_id, n) * 3 ^
            checksum(grid.d_cube_start, grid.n_cubes) * 5 ^
            checksum(grid.d_cube_end, grid.n_cubes) * 7;
        printf("{\"impl\": \"%s\", \"cells\": %d, \"grid_size\": %d, "
               "\"us_per_build\": %.1f, \"checksum\": \"%016llx\"}\n",
            impl, n, gs, ms * 1000 / repeats, sum);
        cudaFree(d_X);
    }
    return 0;
}
