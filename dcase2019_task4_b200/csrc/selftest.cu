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

namespace {
// Experiment / regression test for shifted K-major operands: A rows start `shift` rows into a 256-row SW128
// buffer (written with the swizzle of the absolute row index), 8-row groups `pitch` rows apart (SBO = pitch*128).
// base_mode 0: descriptor base_offset = 0; 1: base_offset = (start_address >> 7) & 7.
__global__ void __launch_bounds__(128)
umma_shift_selftest_kernel(int shift, int pitch, int base_mode, const float* __restrict__ A /*[256][64]*/,
                           const float* __restrict__ B /*[64][64]*/, float* __restrict__ D /*[128][64]*/) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* a_s = smem;                 // 2 blocks x 256 rows x 128 B = 64 KB
    unsigned char* b_s = smem + 65536;         // 2 blocks x 64 rows
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int r = tid; r < 256; r += 128)
        for (int c = 0; c < 16; ++c)
            *reinterpret_cast<float4*>(a_s + (c >> 3) * 32768 + tc::sw128_chunk(r, c & 7)) =
                *reinterpret_cast<const float4*>(A + r * 64 + 4 * c);
    if (tid < 64)
        for (int c = 0; c < 16; ++c)
            *reinterpret_cast<float4*>(b_s + (c >> 3) * 8192 + tc::sw128_chunk(tid, c & 7)) =
                *reinterpret_cast<const float4*>(B + tid * 64 + 4 * c);
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 64);
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        constexpr uint32_t idesc = tc::idesc_tf32(128, 64, 0, 0);
        for (int j = 0; j < 8; ++j) {
            const uint32_t a = tc::smem_u32(a_s) + (j >> 2) * 32768 + shift * 128 + (j & 3) * 32;
            const uint32_t b = tc::smem_u32(b_s) + (j >> 2) * 8192 + (j & 3) * 32;
            uint64_t ad = tc::smem_desc_sw128(a, 16, pitch * 128);
            if (base_mode == 1) ad |= (uint64_t)((a >> 7) & 7) << 49;
            tc::umma_tf32(tmem, ad, tc::smem_desc_sw128(b, 16, 1024), idesc, j > 0 ? 1u : 0u);
        }
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

extern "C" int dcase_selftest_umma_shift(dcase_ctx* ctx, int shift, int pitch, int base_mode, const float* A,
                                         const float* B, float* D, void* stream) {
    DCASE_REQUIRE(ctx && A && B && D, "null argument");
    DCASE_REQUIRE(shift >= 0 && pitch >= 8 && shift + 15 * pitch + 8 <= 256, "operand does not fit the 256-row buffer");
    static bool attr_set = false;
    if (!attr_set) {
        DCASE_CUDA_CHECK(cudaFuncSetAttribute(umma_shift_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 83968));
        attr_set = true;
    }
    umma_shift_selftest_kernel<<<1, 128, 83968, (cudaStream_t)stream>>>(shift, pitch, base_mode, A, B, D);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}
