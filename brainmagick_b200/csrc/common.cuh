// Shared device/host helpers for the brainmagick_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

namespace bm {

// ---- error plumbing (C-ABI returns an int, message kept for bm_last_error()) -------------------
extern thread_local char g_last_error[512];
extern int g_debug_flags;               // bm_set_debug_flags(): experiment switches (bit 0: wgrad keeps X raw as hi)
extern long long* g_debug_buf;           // bm_set_debug_buffer(): device scratch some kernels fill with cycle counters
extern unsigned long long g_launches;   // kernels launched by this library since load (bm_launch_count())
inline int set_error(int code, const char* fmt, const char* a = "", const char* b = "") {
    snprintf(g_last_error, sizeof(g_last_error), fmt, a, b);
    return code;
}
#define BM_CHECK_ARG(cond)                                                                          \
    do {                                                                                            \
        if (!(cond)) return ::bm::set_error(2, "%s: invalid argument: %s", __func__, #cond);        \
    } while (0)
#define BM_CHECK_LAUNCH()                                                                           \
    do {                                                                                            \
        cudaError_t e_ = cudaGetLastError();                                                        \
        ++::bm::g_launches;                                                                         \
        if (e_ != cudaSuccess) return ::bm::set_error(3, "%s: launch failed: %s", __func__, cudaGetErrorString(e_)); \
    } while (0)
#define BM_CUDA(call)                                                                               \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) return ::bm::set_error(3, "%s: %s", #call, cudaGetErrorString(e_));  \
    } while (0)

// ---- math ---------------------------------------------------------------------------------------
// exact (erf) GELU, as nn.GELU() default (bm/models/simpleconv.py:85-86, bm/models/common.py:120)
__device__ __forceinline__ float gelu_f(float z) { return 0.5f * z * (1.0f + erff(z * 0.70710678118654752440f)); }
// d/dz GELU(z) = Phi(z) + z * phi(z)
__device__ __forceinline__ float gelu_grad_f(float z) {
    float cdf = 0.5f * (1.0f + erff(z * 0.70710678118654752440f));
    float pdf = 0.39894228040143267794f * expf(-0.5f * z * z);
    return cdf + z * pdf;
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

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

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace bm
