// Error-reporting helpers for model programs and for the library itself.
//
// Mirrors the two facilities of the reference's include/cudebug.cuh:8-35:
//   D_ASSERT(predicate)  device-side assertion
//   CHECK_CUDA           host-side "did anything go wrong so far?" probe that
//                        prints and exits with status -1
// plus YB_CUDA(call), which the B200 runtime uses internally so that a failed
// allocation or launch can never be silently ignored.
#pragma once

#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <cuda_runtime.h>


#if defined(__APPLE__)
// No device assert on macOS; fall back to a message.
#define D_ASSERT(predicate)                                                  \
    do {                                                                     \
        if (!(predicate))                                                    \
            printf("(%s:%d) Device assertion failed!\n", __FILE__, __LINE__); \
    } while (0)
#else
#define D_ASSERT(predicate) assert(predicate)
#endif


// Reports the sticky launch error first, then drains the device to surface
// asynchronous faults of kernels that are still in flight.
inline void cudaErrorCheck(const char* file, int line)
{
    const cudaError_t launch_status = cudaGetLastError();
    const cudaError_t run_status = cudaDeviceSynchronize();
    if (launch_status != cudaSuccess) {
        printf("Sync CUDA error: %s, %s(%d).\n",
            cudaGetErrorString(launch_status), file, line);
        exit(-1);
    }
    if (run_status != cudaSuccess) {
        printf("Async CUDA error: %s, %s(%d).\n",
            cudaGetErrorString(run_status), file, line);
        exit(-1);
    }
}

#define CHECK_CUDA cudaErrorCheck(__FILE__, __LINE__)


// Library-internal: every runtime call made by the solver goes through this.
inline void yb_cuda_fail(cudaError_t status, const char* what, const char* file,
    int line)
{
    fprintf(stderr, "yalla-b200: %s failed: %s (%s:%d)\n", what,
        cudaGetErrorString(status), file, line);
    abort();
}

#define YB_CUDA(call)                                                  \
    do {                                                               \
        const cudaError_t yb_status_ = (call);                         \
        if (yb_status_ != cudaSuccess)                                 \
            yb_cuda_fail(yb_status_, #call, __FILE__, __LINE__);       \
    } while (0)
