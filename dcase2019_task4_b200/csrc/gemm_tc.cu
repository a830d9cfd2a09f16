// GRU input projections and their data gradients on the 5th-generation tensor cores (tcgen05, fp32-accurate).
//
// Replaces (reference file:line): the `x @ W_ih^T + b_ih` half of nn.GRU (baseline/models/RNN.py:12-15) for both
// directions of a layer, and its autograd counterpart  dX = sum_dir dGi_dir @ W_ih_dir.  These four small GEMMs sit on
// the step's critical chain between the recurrence kernels; they used to run as fp32 FMA tiles (sgemm_batch_kernel).
//
//   C[M][N] (+ bias[n]) = sum over parts p of  A_p[M][K] * B_p          tile 128 x 64, one CTA per tile
//     forward : parts = 1, B = W_ih [N][K] row-major  (K-major operand)
//     backward: parts = 2 (the two directions accumulate in tensor memory), B = W_ih [K][N] row-major (MN-major operand)
//
// Precision: kind::tf32 reads 10 mantissa bits, the GRU parity bar is 2e-5 against fp32.  Every operand tile is therefore
// split in shared memory into hi (what the tensor core sees of the raw fp32 value) and lo = x - hi, and three MMAs per
// K step accumulate  hi*hi + lo*hi + hi*lo  in fp32 (error ~2^-22, the classic 3xTF32 scheme).
//
// Pipeline per CTA (256 threads): TMA (cp.async.bulk.tensor, SWIZZLE_128B / 128B_ATOM_32B, zero fill past M) brings a
// K = 64 stage of A and B straight into the operand layout, double buffered; all threads derive the lo tiles (the
// split is element-wise, so it ignores the swizzle), one elected lane issues the 24 MMAs of the stage, the next
// stage's TMA is already in flight.  Epilogue: TMEM -> registers -> per-warp swizzled staging -> full 128-byte lines.
#include "gemm_tc.cuh"

#include "tc.cuh"
#include "tma.cuh"

namespace {

constexpr int kStageK = 64;                          // K per pipeline stage (two 32-wide SW128 blocks)
constexpr int kABlock = 128 * 128;                   // bytes of one [128 rows][32 fp32] A block
constexpr int kBBlock = 64 * 128;                    // bytes of one B block: K-major [64 n][32 k] or MN-major 2 x [32 k][32 n]
constexpr int kStageBytes = 2 * kABlock + 2 * kBBlock;              // 48 KB
constexpr int kGemmThreads = 256;
constexpr int kGemmSmem = 3 * kStageBytes + 1024 /* alignment slack */ + 64;

// blockIdx.z = problem; operand pair (problem, part) uses maps[problem * parts + part] (problems * parts <= 2)
struct GemmTcArgs {
    CUtensorMap a_map[2];     // [M][K] row-major, box {32, 128}, SWIZZLE_128B
    CUtensorMap b_map[2];     // K-major [N][K] box {32, 64} SWIZZLE_128B | MN-major [K][N] box {32, 32} SWIZZLE_128B_ATOM_32B
    float* C[2];
    const float* bias[2];     // nullable
    int M, N, K, ldc, parts;
};

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

template <bool B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ GemmTcArgs g) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* hi[2] = {smem, smem + kStageBytes};
    unsigned char* lo = smem + 2 * kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * kStageBytes);     // [0,1] stage landed, [2] MMAs of a stage done
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 3);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int m0 = blockIdx.x * 128, n0 = blockIdx.y * 64;
    const int map0 = blockIdx.z * g.parts;
    float* const C = g.C[blockIdx.z];
    const float* const bias = g.bias[blockIdx.z];
    const int stages_per_part = g.K / kStageK;
    const int n_stages = stages_per_part * g.parts;

    if (tid == 0) {
        tc::mbar_init(&bars[0], 1); tc::mbar_init(&bars[1], 1); tc::mbar_init(&bars[2], 1);
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_base_s, 64);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_base_s;

    auto load_stage = [&](int s) {              // one thread: four (K-major B) or six (MN-major B) TMA boxes
        const int part = s / stages_per_part, k0 = (s - part * stages_per_part) * kStageK;
        unsigned char* dst = hi[s & 1];
        mbar_expect_tx(&bars[s & 1], (uint32_t)kStageBytes);
        tma_load_2d(dst, &g.a_map[map0 + part], k0, m0, &bars[s & 1]);
        tma_load_2d(dst + kABlock, &g.a_map[map0 + part], k0 + 32, m0, &bars[s & 1]);
        unsigned char* b = dst + 2 * kABlock;
        if (!B_MN) {
            tma_load_2d(b, &g.b_map[map0 + part], k0, n0, &bars[s & 1]);
            tma_load_2d(b + kBBlock, &g.b_map[map0 + part], k0 + 32, n0, &bars[s & 1]);
        } else {                                // per 32-wide K chunk: n-block 0 | n-block 1, 4 KB each
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                tma_load_2d(b + c * kBBlock, &g.b_map[map0 + part], n0, k0 + 32 * c, &bars[s & 1]);
                tma_load_2d(b + c * kBBlock + 4096, &g.b_map[map0 + part], n0 + 32, k0 + 32 * c, &bars[s & 1]);
            }
        }
    };
    if (tid == 0) load_stage(0);

    uint32_t ph_full[2] = {0, 0}, ph_done = 0;
    for (int s = 0; s < n_stages; ++s) {
        if (s > 0) {                            // the MMAs of stage s - 1 have read hi[(s + 1) & 1] and lo
            tc::mbar_wait(&bars[2], ph_done);
            ph_done ^= 1;
        }
        if (tid == 0 && s + 1 < n_stages) load_stage(s + 1);
        tc::mbar_wait(&bars[s & 1], ph_full[s & 1]);
        ph_full[s & 1] ^= 1;
        // lo = x - hi(x), element-wise over the whole stage (layout agnostic)
        const float4* src = reinterpret_cast<const float4*>(hi[s & 1]);
        float4* dst = reinterpret_cast<float4*>(lo);
#pragma unroll 4
        for (int i = tid; i < kStageBytes / 16; i += kGemmThreads) {
            const float4 v = src[i];
            dst[i] = make_float4(v.x - tf32_hi(v.x), v.y - tf32_hi(v.y), v.z - tf32_hi(v.z), v.w - tf32_hi(v.w));
        }
        tc::fence_proxy_async();
        __syncthreads();
        if (warp == 0) {
            tc::fence_after_sync();
            constexpr uint32_t idesc = tc::idesc_tf32(128, 64, 0, B_MN ? 1 : 0);
            const uint32_t a_hi_lo = tc::desc_lo(tc::smem_u32(hi[s & 1]), 16), a_lo_lo = tc::desc_lo(tc::smem_u32(lo), 16);
            const uint32_t k_hi = tc::desc_hi(1024, 2);
            const uint32_t bh = tc::smem_u32(hi[s & 1]) + 2 * kABlock, bl = tc::smem_u32(lo) + 2 * kABlock;
            const uint32_t b_hi_lo = B_MN ? tc::desc_lo(bh, 4096) : tc::desc_lo(bh, 16);
            const uint32_t b_lo_lo = B_MN ? tc::desc_lo(bl, 4096) : tc::desc_lo(bl, 16);
            const uint32_t b_hi = B_MN ? tc::desc_hi(512, 1) : k_hi;
#pragma unroll
            for (int j = 0; j < 8; ++j) {       // K steps of 8 inside the stage
                const uint32_t a_off = (uint32_t)(((j >> 2) * kABlock + (j & 3) * 32) >> 4);
                const uint32_t b_off = B_MN ? (uint32_t)(((j >> 2) * kBBlock + (j & 3) * 1024) >> 4)
                                            : (uint32_t)(((j >> 2) * kBBlock + (j & 3) * 32) >> 4);
                tc::umma_tf32_elect(tmem, a_hi_lo + a_off, k_hi, b_hi_lo + b_off, b_hi, idesc, (s > 0 || j > 0) ? 1u : 0u);
                tc::umma_tf32_elect(tmem, a_lo_lo + a_off, k_hi, b_hi_lo + b_off, b_hi, idesc, 1u);
                tc::umma_tf32_elect(tmem, a_hi_lo + a_off, k_hi, b_lo_lo + b_off, b_hi, idesc, 1u);
            }
            tc::umma_commit_elect(&bars[2]);
        }
    }
    tc::mbar_wait(&bars[2], ph_done);
    tc::fence_after_sync();

    // ---- epilogue: warps 0..3 own the four TMEM lane quadrants (rows 32 w .. 32 w + 31); rows leave as full lines ----
    if (warp < 4) {
        float acc[64];
        tc::tmem_ld_row64(tmem, warp, 0, acc);
        const uint32_t stage_a = tc::smem_u32(hi[0]) + (uint32_t)warp * 2048u;      // the operand buffers are free now
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
            for (int rh = 0; rh < 2; ++rh) {
                if ((lane >> 4) == rh) {
                    const uint32_t rb = stage_a + (uint32_t)(lane & 15) * 128u;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int e = 32 * ch + 4 * q;
                        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (bias) b4 = __ldg(reinterpret_cast<const float4*>(bias + n0 + e));
                        st_shared_v4(rb + (uint32_t)((q ^ (lane & 7)) << 4), acc[e] + b4.x, acc[e + 1] + b4.y, acc[e + 2] + b4.z,
                                     acc[e + 3] + b4.w);
                    }
                }
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int idx = lane + 32 * q, r = idx >> 3, c = idx & 7;
                    const float4 v = ld_shared_v4(stage_a + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4));
                    const int m = m0 + 32 * warp + 16 * rh + r;
                    if (m < g.M) *reinterpret_cast<float4*>(C + (long long)m * g.ldc + n0 + 32 * ch + 4 * c) = v;
                }
                __syncwarp();
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 64);
}

}  // namespace

#define DCASE_TRY_RC(expr) do { int rc__ = (expr); if (rc__ != DCASE_OK) return rc__; } while (0)

int gemm_tc_init() {
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    return DCASE_OK;
}

int launch_gemm_tc(const GemmTcBatch& p, cudaStream_t s) {
    DCASE_PROF("gemm_tc", s);
    DCASE_REQUIRE(p.problems >= 1 && p.parts >= 1 && p.problems * p.parts <= 2, "at most two operand pairs per launch");
    DCASE_REQUIRE(p.M > 0 && p.N % 64 == 0 && p.K % kStageK == 0, "N must be a multiple of 64 and K of 64");
    GemmTcArgs g{};
    for (int i = 0; i < p.problems * p.parts; ++i) {
        DCASE_REQUIRE((reinterpret_cast<uintptr_t>(p.A[i]) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.B[i]) & 15) == 0 &&
                      p.lda % 4 == 0 && p.ldb % 4 == 0, "TMA needs 16-byte aligned operands and row strides");
        DCASE_TRY_RC(make_matrix_map(&g.a_map[i], p.A[i], p.K, p.M, p.lda, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B));
        if (p.b_mn_major)
            DCASE_TRY_RC(make_matrix_map(&g.b_map[i], p.B[i], p.N, p.K, p.ldb, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
        else
            DCASE_TRY_RC(make_matrix_map(&g.b_map[i], p.B[i], p.K, p.N, p.ldb, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B));
    }
    for (int i = 0; i < p.problems; ++i) {
        DCASE_REQUIRE(p.ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(p.C[i]) & 15) == 0, "C rows must be 16-byte aligned");
        g.C[i] = p.C[i];
        g.bias[i] = p.bias[i];
    }
    g.M = p.M; g.N = p.N; g.K = p.K; g.ldc = p.ldc; g.parts = p.parts;
    const dim3 grid((p.M + 127) / 128, p.N / 64, p.problems);
    if (p.b_mn_major) gemm_tc_kernel<true><<<grid, kGemmThreads, kGemmSmem, s>>>(g);
    else gemm_tc_kernel<false><<<grid, kGemmThreads, kGemmSmem, s>>>(g);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}
