// Per-cell data that is not integrated (cell type, counters, RNG states, …),
// kept as a host array and a device array of n_max entries, indexed by the
// ORIGINAL cell id (reference: include/property.cuh:7-34). The solver never
// permutes these arrays: pairwise functors receive original ids i, j.
#pragma once

#include <stdlib.h>
#include <string>

#include "cudebug.cuh"


template<typename Prop = int>
struct Property {
    Prop* h_prop;
    Prop* d_prop;
    std::string name;
    const int n_max;

    Property(int n_max, std::string name = "cell_type")
        : name{name}, n_max{n_max}
    {
        const size_t bytes = static_cast<size_t>(n_max) * sizeof(Prop);
        h_prop = static_cast<Prop*>(malloc(bytes));
        YB_CUDA(cudaMalloc(&d_prop, bytes));
    }
    Property(const Property&) = delete;
    Property& operator=(const Property&) = delete;
    ~Property()
    {
        cudaFree(d_prop);
        free(h_prop);
    }

    // Both directions move all n_max entries and block, as in the reference.
    void copy_to_device()
    {
        YB_CUDA(cudaMemcpy(d_prop, h_prop,
            static_cast<size_t>(n_max) * sizeof(Prop), cudaMemcpyHostToDevice));
    }
    void copy_to_host()
    {
        YB_CUDA(cudaMemcpy(h_prop, d_prop,
            static_cast<size_t>(n_max) * sizeof(Prop), cudaMemcpyDeviceToHost));
    }
};
