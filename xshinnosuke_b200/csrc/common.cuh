// Shared helpers for libxsb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define XSB_OK 0
#define XSB_ERR_CUDA 1
#define XSB_ERR_ARG 2
#define XSB_ERR_UNSUPPORTED 3

// implemented in runtime.cu
extern "C" const char* xsb_last_error(void);
void xsb_set_error(const char* fmt, ...);
int xsb_sm_count_cached();

#define XSB_CHECK_ARG(cond, ...)                 \
    do {                                         \
        if (!(cond)) {                           \
            xsb_set_error(__VA_ARGS__);          \
            return XSB_ERR_ARG;                  \
        }                                        \
    } while (0)

#define XSB_CUDA(call)                                                                     \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            xsb_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return XSB_ERR_CUDA;                                                           \
        }                                                                                  \
    } while (0)

#define XSB_LAUNCH_CHECK()                                                                 \
    do {                                                                                   \
        cudaError_t e__ = cudaPeekAtLastError();                                           \
        if (e__ != cudaSuccess) {                                                          \
            xsb_set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            (void)cudaGetLastError();                                                      \
            return XSB_ERR_CUDA;                                                           \
        }                                                                                  \
    } while (0)

static inline cudaStream_t xsb_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming 128-bit accesses: activations are touched once per kernel, keep them out of L1
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
// same, for buffers the kernel may also write (in-place ops): no .nc
__device__ __forceinline__ float4 ldg_stream_rw(const float4* p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
