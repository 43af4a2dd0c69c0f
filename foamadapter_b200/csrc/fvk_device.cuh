// Device-side helpers shared by the kernel translation units.
#pragma once
#include "fvk_internal.hpp"

#include <cuda_runtime.h>

#define FVK_CUDA(call)                                                                             \
    do                                                                                             \
    {                                                                                              \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fvk_fail(                                                                       \
                (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? FVK_ENODEVICE     \
                                                                               : FVK_ECUDA,        \
                "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));             \
    } while (0)

#define FVK_LAUNCH_CHECK() FVK_CUDA(cudaGetLastError())

static inline cudaStream_t fvk_cu(fvk_stream s) { return reinterpret_cast<cudaStream_t>(s); }

// number of SMs of the current device (cached); B200 = 148
int fvk_sm_count();
// current experiment variant (fvk_set_variant)
int fvk_variant();
bool fvk_brick_config(int cfg[3]); // true: {cells per thread, threads per block, resident blocks} override is set

struct Vec3d
{
    double x, y, z;
};

__device__ __forceinline__ Vec3d ld3(const double* __restrict__ p, int64_t i)
{
    return Vec3d {p[3 * i], p[3 * i + 1], p[3 * i + 2]};
}
__device__ __forceinline__ void st3(double* __restrict__ p, int64_t i, Vec3d v)
{
    p[3 * i] = v.x;
    p[3 * i + 1] = v.y;
    p[3 * i + 2] = v.z;
}

// streaming (read-once) loads: bypass L1 allocation so gathered data keeps the L1
__device__ __forceinline__ double ld_stream(const double* p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_stream(const int* p)
{
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
