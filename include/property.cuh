// Per-cell data that is not integrated (cell type, counters, RNG states, …),
// kept as a host array and a device array of n_max entries, indexed by the
// ORIGINAL cell id (reference: include/property.cuh:7-34). The solver never
// permutes these arrays: pairwise functors receive original ids i, j.
#pragma once

#include <stdlib.h>
#include <string>

#include "cudebug.cuh"


namespace yb {

// A host buffer and a device buffer of the same size, moved as a whole. All
// Property<T> instantiations share this one piece of code.
class Mirrored_bytes {
protected:
    explicit Mirrored_bytes(size_t n_bytes) : n_bytes{n_bytes}
    {
        host_bytes = malloc(n_bytes);
        YB_CUDA(cudaMalloc(&device_bytes, n_bytes));
    }
    ~Mirrored_bytes()
    {
        cudaFree(device_bytes);
        free(host_bytes);
    }
    Mirrored_bytes(const Mirrored_bytes&) = delete;
    Mirrored_bytes& operator=(const Mirrored_bytes&) = delete;

    void push() const  // host -> device, blocking
    {
        YB_CUDA(cudaMemcpy(
            device_bytes, host_bytes, n_bytes, cudaMemcpyHostToDevice));
    }
    void pull() const  // device -> host, blocking
    {
        YB_CUDA(cudaMemcpy(
            host_bytes, device_bytes, n_bytes, cudaMemcpyDeviceToHost));
    }

    void* host_bytes = nullptr;
    void* device_bytes = nullptr;
    const size_t n_bytes;
};

}  // namespace yb


template<typename Prop = int>
struct Property : private yb::Mirrored_bytes {
    Prop* const h_prop = static_cast<Prop*>(host_bytes);
    Prop* const d_prop = static_cast<Prop*>(device_bytes);
    std::string name;
    const int n_max;

    Property(int n_max, std::string name = "cell_type")
        : yb::Mirrored_bytes{static_cast<size_t>(n_max) * sizeof(Prop)},
          name{name}, n_max{n_max}
    {}

    // Both directions move all n_max entries and block, as in the reference.
    void copy_to_device() { push(); }
    void copy_to_host() { pull(); }
};
