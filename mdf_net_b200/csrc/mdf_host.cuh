// mdf_host.cuh -- host-side plumbing shared by the C-ABI entry points: status codes, pointer
// validation, a device guard (entry points run on the device that owns the output pointer and are
// re-entrant: the reference's nn.DataParallel drives one Python thread per GPU, train.py:25).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/mdf_b200.h"

namespace mdf {

extern thread_local int g_last_cuda_error;

inline int cuda_fail(cudaError_t e)
{
    g_last_cuda_error = (int)e;
    return MDF_ERR_CUDA;
}

#define MDF_CUDA_TRY(expr)                                   \
    do {                                                     \
        cudaError_t _e = (expr);                             \
        if (_e != cudaSuccess) return ::mdf::cuda_fail(_e);  \
    } while (0)

// Returns the owning device of a device (or managed) pointer, or a negative status.
inline int device_of(const void* p)
{
    if (p == nullptr) return MDF_ERR_NULL_POINTER;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return MDF_ERR_NOT_DEVICE;
    }
    if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) return MDF_ERR_NOT_DEVICE;
    return a.device;
}

// All pointers must live on `dev`.
inline int check_on_device(int dev, const void* const* ptrs, int n)
{
    for (int i = 0; i < n; ++i) {
        int d = device_of(ptrs[i]);
        if (d < 0) return d;
        if (d != dev) return MDF_ERR_NOT_DEVICE;
    }
    return MDF_OK;
}

class DeviceGuard {
public:
    explicit DeviceGuard(int dev) : prev_(-1), changed_(false)
    {
        if (cudaGetDevice(&prev_) == cudaSuccess && prev_ != dev) changed_ = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard()
    {
        if (changed_) cudaSetDevice(prev_);
    }
private:
    int prev_;
    bool changed_;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline int launch_status()
{
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? MDF_OK : cuda_fail(e);
}

}  // namespace mdf
