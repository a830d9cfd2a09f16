// Shared device helpers for the dcase_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define DCASE_OK 0
#define DCASE_ERR_ARG -1
#define DCASE_ERR_CUDA -2
#define DCASE_ERR_STATE -3

void dcase_set_error(const char* fmt, ...);

#define DCASE_CUDA_CHECK(expr)                                                              \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            dcase_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                   \
                            cudaGetErrorString(e__));                                       \
            return DCASE_ERR_CUDA;                                                          \
        }                                                                                   \
    } while (0)

// every kernel launch goes through this: counts launches (bench.py's gpu_launches) and checks the launch
extern unsigned long long g_dcase_launches;
#define DCASE_LAUNCH_CHECK()                 \
    do {                                     \
        ++g_dcase_launches;                  \
        DCASE_CUDA_CHECK(cudaGetLastError()); \
    } while (0)

// Optional per-kernel CUDA-event timing on the launching stream (dcase_profile_begin / _end).
struct DcaseProfScope {
    int slot;
    cudaStream_t stream;
    DcaseProfScope(const char* name, cudaStream_t s);
    ~DcaseProfScope();
};
#define DCASE_PROF(name, stream) DcaseProfScope prof_scope__(name, stream)

#define DCASE_REQUIRE(cond, msg)                                                            \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            dcase_set_error("%s:%d: requirement failed: %s (%s)", __FILE__, __LINE__, #cond, msg); \
            return DCASE_ERR_ARG;                                                           \
        }                                                                                   \
    } while (0)

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (RNG contract: include/dcase_b200.h; numpy restatement: oracle/philox.py)
// ------------------------------------------------------------------------------------------
#define DCASE_STREAM_HEAD 3u
#define DCASE_STREAM_NOISE 4u

__host__ __device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2,
                                                      uint32_t& c3, uint32_t k0, uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
}

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint64_t row, uint32_t stream, uint32_t step,
                                                        uint64_t seed) {
    uint32_t c0 = (uint32_t)row, c1 = (uint32_t)(row >> 32), c2 = stream, c3 = step;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// Per-step scalars that live in device memory so a captured CUDA graph replays with fresh values.
struct DcaseStepScalars {
    uint64_t seed;        // Philox key
    uint32_t step;        // Philox counter word 3 (global step)
    float cons_weight;    // consistency weight (main.py:127)
    float ema_alpha;      // main.py:47
    float lr;
    float bias_corr1;     // 1 - beta1^t
    float bias_corr2;     // 1 - beta2^t
    float grad_scale;     // 1 / world_size under DP
    float pad_;
};

__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
