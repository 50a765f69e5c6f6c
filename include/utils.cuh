// Small helpers shared by several headers (reference: include/utils.cuh:10-33).
#pragma once

#include <curand_kernel.h>
#include <sstream>  // for user code that relies on it coming with utils.cuh
#include <string>
#include <vector>


// One XORWOW generator per index: sequence i of the given seed, offset 0
// (reference: utils.cuh:29-33). Keeping (seed, sequence=i, offset=0) is what
// makes noisy runs comparable between the two builds for equal seeds.
__global__ void setup_rand_states(int n_states, int seed, curandState* d_state)
{
    // grid-stride, so any launch shape initialises all n_states generators
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_states; i += stride)
        curand_init(seed, i, 0, d_state + i);
}


// Euclidean inner product of the spatial part of two points.
template<typename Pt_a, typename Pt_b>
__device__ __host__ float dot_product(Pt_a a, Pt_b b)
{
    return a.x * b.x + a.y * b.y + a.z * b.z;
}


// Tokenise a line at single blanks. Consecutive blanks yield empty tokens,
// exactly like std::getline(…, ' ') does in the reference (utils.cuh:10-20);
// the VTK reader depends on that for lines with a leading keyword.
inline std::vector<std::string> split(const std::string& s)
{
    std::vector<std::string> tokens;
    for (size_t from = 0; from < s.size();) {
        const size_t blank = s.find(' ', from);
        if (blank == std::string::npos) {
            tokens.push_back(s.substr(from));
            break;
        }
        tokens.push_back(s.substr(from, blank - from));
        from = blank + 1;
    }
    return tokens;
}
