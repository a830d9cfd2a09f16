// Self-test of the tcgen05 primitives in tc.cuh (descriptor encodings, SW128 operand layout, TMEM read-back).
// Exposed through the C ABI so the GPU test-suite pins them against a plain matmul (tests/test_gpu_tcgen05.py).
#include "../../include/dcase_b200.h"
#include "common.cuh"
#include "ctx.h"
#include "tc.cuh"

namespace {

// mode 0: D[128][64] = A[128][64] * B[64][64]^T          (A, B K-major)
// mode 1: raw TMEM dump [128 lanes][64 cols] of  D[m][n] = sum_p A[p][m] * B[p][n],  p < 128  (A, B MN-major, M = 64)
__global__ void __launch_bounds__(128)
umma_selftest_kernel(int mode, const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // SW128 atoms: 1024-B aligned
    unsigned char* a_s = smem;                 // 2 blocks x 128 rows x 128 B = 32 KB
    unsigned char* b_s = smem + 32768;         // up to 2 blocks x 128 rows
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int b_rows = mode == 0 ? 64 : 128;
    // fill operands (thread = row)
    for (int c = 0; c < 16; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(A + tid * 64 + 4 * c);
        *reinterpret_cast<float4*>(a_s + (c >> 3) * (128 * 128) +
                                   (mode == 0 ? tc::sw128_chunk(tid, c & 7) : tc::sw128b32_chunk(tid, c & 7))) = v;
    }
    if (tid < b_rows) {
        for (int c = 0; c < 16; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(B + tid * 64 + 4 * c);
            *reinterpret_cast<float4*>(b_s + (c >> 3) * (b_rows * 128) +
                                       (mode == 0 ? tc::sw128_chunk(tid, c & 7) : tc::sw128b32_chunk(tid, c & 7))) = v;
        }
    }
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 64);
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        if (mode == 0) tc::umma_128x64x64_kmajor(tmem, tc::smem_u32(a_s), tc::smem_u32(b_s), false);
        else tc::umma_64x64_mnmajor(tmem, tc::smem_u32(a_s), tc::smem_u32(b_s), 128, false);
        tc::umma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    float v[64];
    tc::tmem_ld_row64(tmem, warp, 0, v);
    for (int c = 0; c < 64; ++c) D[tid * 64 + c] = v[c];
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 64);
}

}  // namespace

extern "C" int dcase_selftest_umma(dcase_ctx* ctx, int mode, const float* A, const float* B, float* D, void* stream) {
    DCASE_REQUIRE(ctx && A && B && D, "null argument");
    DCASE_REQUIRE(mode == 0 || mode == 1, "mode must be 0 or 1");
    static bool attr_set = false;
    if (!attr_set) {
        DCASE_CUDA_CHECK(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560));
        attr_set = true;
    }
    umma_selftest_kernel<<<1, 128, 66560, (cudaStream_t)stream>>>(mode, A, B, D);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}
