// Shared helpers for libnuhtc_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/nuhtc_b200.h"

#define NUHTC_API extern "C" __attribute__((visibility("default")))

void nuhtc_set_error(const char *fmt, ...);

#define NUHTC_CHECK_ARG(cond, ...)        \
    do {                                  \
        if (!(cond)) {                    \
            nuhtc_set_error(__VA_ARGS__); \
            return NUHTC_EINVAL;          \
        }                                 \
    } while (0)

#define NUHTC_CUDA(call)                                                                         \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            nuhtc_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return NUHTC_ECUDA;                                                                  \
        }                                                                                        \
    } while (0)

#define NUHTC_LAUNCH_CHECK() NUHTC_CUDA(cudaGetLastError())

// Function attributes (opt-in shared memory), occupancy-derived grid sizes and the SM count belong to a DEVICE, not to
// the process: every cache of them is an array indexed by the current device ordinal, so one process may drive several GPUs.
constexpr int kNuhtcMaxDevices = 64;
static inline int nuhtc_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev < 0 || dev >= kNuhtcMaxDevices) ? 0 : dev;
}
static inline int nuhtc_sm_count() {
    static int n[kNuhtcMaxDevices] = {0};
    const int dev = nuhtc_device();
    if (n[dev] == 0) {
        cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
        if (n[dev] <= 0) n[dev] = 148;
    }
    return n[dev];
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// streaming 128-bit store: the RoIAlign / paste outputs are written once and not re-read by us
__device__ __forceinline__ void st_stream_f4(float *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_stream_u4(void *p, uint4 v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ float4 ldg_f4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
