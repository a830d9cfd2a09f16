// Self-test of the tcgen05 primitives in tc.cuh (descriptor encodings, SW128 operand layout, TMEM read-back).
// Exposed through the C ABI so the GPU test-suite pins them against a plain matmul (tests/test_gpu_tcgen05.py).
#include "../../include/dcase_b200.h"
#include "common.cuh"
#include "ctx.h"
#include "tc.cuh"
#include <cuda_bf16.h>

namespace {

// mode 0: D[128][64] = A[128][64] * B[64][64]^T          (A, B K-major)
// mode 1: raw TMEM dump [128 lanes][64 cols] of  D[m][n] = sum_p A[p][m] * B[p][n],  p < 128  (A, B MN-major, M = 64)
// mode 2: D[128][64] = (A B^T) B^T with the second GEMM's A operand read straight from the first one's accumulator in
//         tensor memory (tc::umma_tf32_tmem_a_elect), both GEMMs issued back to back by one warp
// mode 3: the same, followed by a third GEMM that overwrites (doubles) the first accumulator once the chained GEMM has
//         COMPLETED, the order cnn0 uses for the next tile's conv
// mode 4: D[m][j] = sum_p [A | B][p][m] * A[p][j], m < 128, j < 16, p < 128 with bf16 operands (kind::f16), both MN-major in
//         the 16-bit SWIZZLE_128B layout: rows = K index p, a 128-byte row = 64 consecutive M / N elements, 16-byte chunk c of
//         row p at chunk c ^ (p & 7); the two 64-wide M blocks are LBO = 16 KB apart, 8-row K groups SBO = 1 KB apart, one
//         instruction consumes K = 16 rows (2 KB).  The layout cnn0's backward accumulates its parameter gradients with.
__global__ void __launch_bounds__(128)
umma_selftest_kernel(int mode, const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // SW128 atoms: 1024-B aligned
    unsigned char* a_s = smem;                 // 2 blocks x 128 rows x 128 B = 32 KB
    unsigned char* b_s = smem + 32768;         // up to 2 blocks x 128 rows
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int b_rows = mode == 1 ? 128 : 64;
    if (mode == 4) {
        __shared__ uint64_t bar4;
        // A operand: block 0 = A[p][0..63], block 1 = B[p][0..63]; B operand: A[p][0..15] in the first 32 bytes of its row
        unsigned char* e_s = smem + 32768;
        for (int blk = 0; blk < 2; ++blk) {
            const float* src = (blk ? B : A) + tid * 64;
            for (int c = 0; c < 8; ++c) {
                __nv_bfloat162 h[4];
                for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(src[8 * c + 2 * e], src[8 * c + 2 * e + 1]);
                *reinterpret_cast<uint4*>(a_s + blk * 16384 + tc::sw128_chunk(tid, c)) = *reinterpret_cast<uint4*>(h);
            }
        }
        for (int c = 0; c < 8; ++c) {
            __nv_bfloat162 h[4];
            for (int e = 0; e < 4; ++e)
                h[e] = c < 2 ? __floats2bfloat162_rn(A[tid * 64 + 8 * c + 2 * e], A[tid * 64 + 8 * c + 2 * e + 1]) : __floats2bfloat162_rn(0.f, 0.f);
            *reinterpret_cast<uint4*>(e_s + tc::sw128_chunk(tid, c)) = *reinterpret_cast<uint4*>(h);
        }
        if (tid == 0) { tc::mbar_init(&bar4, 1); tc::fence_mbar_init(); }
        if (warp == 0) tc::tmem_alloc(&tmem_base_s, 128);
        tc::fence_proxy_async();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        const uint32_t tm = tmem_base_s;
        const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
        if (warp_u == 0) {
            constexpr uint32_t idesc = tc::idesc_bf16(128, 16, 1, 1);
            const uint32_t a_lo = tc::desc_lo(tc::smem_u32(a_s), 16384), e_lo = tc::desc_lo(tc::smem_u32(e_s), 16384);
            const uint32_t hi = tc::desc_hi(1024, 2);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                tc::umma_f16_elect(tm, a_lo + (j * 2048 >> 4), hi, e_lo + (j * 2048 >> 4), hi, idesc, j > 0 ? 1u : 0u);
            tc::umma_commit_elect(&bar4);
        }
        tc::mbar_wait(&bar4, 0);
        tc::fence_after_sync();
        float v[16];
        tc::tmem_ld16(tm + ((uint32_t)(warp * 32) << 16), v);
        tc::tmem_ld_wait();
        for (int c = 0; c < 64; ++c) D[tid * 64 + c] = c < 16 ? v[c] : 0.f;
        tc::fence_before_sync();
        __syncthreads();
        if (warp == 0) tc::tmem_dealloc(tm, 128);
        return;
    }
    // fill operands (thread = row)
    for (int c = 0; c < 16; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(A + tid * 64 + 4 * c);
        *reinterpret_cast<float4*>(a_s + (c >> 3) * (128 * 128) +
                                   (mode != 1 ? tc::sw128_chunk(tid, c & 7) : tc::sw128b32_chunk(tid, c & 7))) = v;
    }
    if (tid < b_rows) {
        for (int c = 0; c < 16; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(B + tid * 64 + 4 * c);
            *reinterpret_cast<float4*>(b_s + (c >> 3) * (b_rows * 128) +
                                       (mode != 1 ? tc::sw128_chunk(tid, c & 7) : tc::sw128b32_chunk(tid, c & 7))) = v;
        }
    }
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 128);
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    if (mode >= 2) {
        // Hazards through a TMEM A operand are NOT interlocked by the tensor core (round 2 measured garbage when the reading
        // GEMM was queued directly behind the producing one): each hand-over needs tcgen05.commit + an mbarrier wait,
        // exactly like a thread reading the result.  Every thread follows all barrier phases in order.
        constexpr uint32_t idesc = tc::idesc_tf32(128, 64, 0, 0);
        const uint32_t a_lo = tc::desc_lo(tc::smem_u32(a_s), 16), b_lo = tc::desc_lo(tc::smem_u32(b_s), 16);
        const uint32_t hi = tc::desc_hi(1024, 2);
        const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
        if (warp_u == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                tc::umma_tf32_elect(tmem, a_lo + (((j >> 2) * 16384 + (j & 3) * 32) >> 4), hi,
                                    b_lo + (((j >> 2) * 8192 + (j & 3) * 32) >> 4), hi, idesc, j > 0 ? 1u : 0u);
            tc::umma_commit_elect(&bar);
        }
        tc::mbar_wait(&bar, 0);                              // the first accumulator is complete
        tc::fence_after_sync();
        if (warp_u == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j)        // A = columns [8 j, 8 j + 8) of the first accumulator
                tc::umma_tf32_tmem_a_elect(tmem + 64, tmem + 8 * j, b_lo + (((j >> 2) * 8192 + (j & 3) * 32) >> 4), hi, idesc,
                                           j > 0 ? 1u : 0u);
            tc::umma_commit_elect(&bar);
        }
        tc::mbar_wait(&bar, 1);                              // the chained GEMM is complete (and has read its A operand)
        tc::fence_after_sync();
        if (mode == 3) {
            if (warp_u == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    tc::umma_tf32_elect(tmem, a_lo + (((j >> 2) * 16384 + (j & 3) * 32) >> 4), hi,
                                        b_lo + (((j >> 2) * 8192 + (j & 3) * 32) >> 4), hi, idesc, 1u);
                tc::umma_commit_elect(&bar);
            }
            tc::mbar_wait(&bar, 0);
            tc::fence_after_sync();
        }
    } else {
        if (tid == 0) {
            if (mode == 0) tc::umma_128x64x64_kmajor(tmem, tc::smem_u32(a_s), tc::smem_u32(b_s), false);
            else tc::umma_64x64_mnmajor(tmem, tc::smem_u32(a_s), tc::smem_u32(b_s), 128, false);
            tc::umma_commit(&bar);
        }
        tc::mbar_wait(&bar, 0);
        tc::fence_after_sync();
    }
    float v[64];
    tc::tmem_ld_row64(tmem, warp, mode >= 2 ? 64 : 0, v);
    for (int c = 0; c < 64; ++c) D[tid * 64 + c] = v[c];
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

}  // namespace

extern "C" int dcase_selftest_umma(dcase_ctx* ctx, int mode, const float* A, const float* B, float* D, void* stream) {
    DCASE_REQUIRE(ctx && A && B && D, "null argument");
    DCASE_REQUIRE(mode >= 0 && mode <= 4, "mode must be 0 .. 4");
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

namespace {
// Micro-benchmark: one warp per CTA issues `reps` back-to-back tcgen05.mma kind::tf32 (shape M x N x 8, both operands
// in shared memory) and reports cycles per MMA (clock64 around issue .. completion).  Used to size the kernels:
// small-N tf32 MMAs are bound by the tensor core's shared-memory operand fetch, not by its math rate.
__global__ void __launch_bounds__(32)
umma_bench_kernel(int M, int N, int a_mn, int b_mn, int reps, int a_sbo, int a_shift, int commit_every, float* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tmem_base_s;
    if (threadIdx.x == 0) tc::mbar_init(&bar2, 1);
    for (int i = threadIdx.x; i < 128 * 1024 / 16; i += 32) reinterpret_cast<float4*>(smem)[i] = make_float4(1.f, 0.5f, 0.25f, 2.f);
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    tc::tmem_alloc(&tmem_base_s, 256);
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncwarp();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t a_base = tc::smem_u32(smem), b_base = a_base + 65536;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
                           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t a_hi = a_mn ? tc::desc_hi(512, 1) : tc::desc_hi(a_sbo, 2), b_hi = b_mn ? tc::desc_hi(512, 1) : tc::desc_hi(1024, 2);
    const uint32_t a_lo = tc::desc_lo(a_base + a_shift, 16384), b_lo = tc::desc_lo(b_base, 16384);
    const uint32_t a_step = a_mn ? 64 : 2, b_step = b_mn ? 64 : 2;      // next K step, in 16-byte units
    if (commit_every < 0) {
        // the conv3x3 issue pattern: 9 taps x 4 K steps with compile-time operand offsets, one commit per 36 MMAs
        constexpr int PITCH = 10;
        const long long c0 = clock64();
        for (int r = 0; r < reps; r += 36) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int dy = tap / 3 - 1, dx = tap % 3 - 1;
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)
                    tc::umma_tf32_elect(tmem, a_lo + ((((1 + dy) * PITCH + 1 + dx) * 128 + k4 * 32) >> 4), a_hi,
                                        b_lo + ((tap * 4096 + k4 * 32) >> 4), b_hi, idesc, (tap > 0 || k4 > 0) ? 1u : 0u);
            }
            tc::umma_commit_elect(&bar2);
        }
        tc::umma_commit_elect(&bar);
        tc::mbar_wait(&bar, 0);
        const long long c1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = (float)(c1 - c0) / (float)reps;
        tc::fence_before_sync();
        __syncwarp();
        tc::tmem_dealloc(tmem, 256);
        return;
    }
    const long long t0 = clock64();
    for (int r = 0; r < reps; r += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            tc::umma_tf32_elect(tmem, a_lo + j * a_step, a_hi, b_lo + j * b_step, b_hi, idesc, 1u);
        if (commit_every > 0 && (r + 4) % commit_every == 0) tc::umma_commit_elect(&bar2);    // nobody waits on bar2
    }
    tc::umma_commit_elect(&bar);
    tc::mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (float)(t1 - t0) / (float)reps;
    tc::fence_before_sync();
    __syncwarp();
    tc::tmem_dealloc(tmem, 256);
}
}  // namespace

extern "C" int dcase_bench_umma(dcase_ctx* ctx, int M, int N, int a_mn_major, int b_mn_major, int reps, int n_ctas,
                                int a_sbo_bytes, int a_shift_bytes, int commit_every, float* cycles_per_mma, void* stream) {
    DCASE_REQUIRE(ctx && cycles_per_mma, "null argument");
    DCASE_REQUIRE((M == 64 || M == 128) && N >= 8 && N <= 256 && N % 8 == 0 && (M == 64 || N % 16 == 0), "illegal MMA shape");
    DCASE_REQUIRE(reps >= 4 && reps % 4 == 0 && n_ctas >= 1, "reps must be a positive multiple of 4");
    static bool attr_set = false;
    if (!attr_set) {
        DCASE_CUDA_CHECK(cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
        attr_set = true;
    }
    DCASE_REQUIRE(a_sbo_bytes >= 1024 && a_sbo_bytes % 128 == 0 && a_shift_bytes % 128 == 0 && a_shift_bytes + 16 * a_sbo_bytes <= 49152,
                  "A operand does not fit the 64 KB buffer");
    umma_bench_kernel<<<n_ctas, 32, 131072, (cudaStream_t)stream>>>(M, N, a_mn_major, b_mn_major, reps, a_sbo_bytes, a_shift_bytes,
                                                                   commit_every, cycles_per_mma);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}
