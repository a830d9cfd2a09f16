// CNN blocks 1 and 2, forward:  BatchNorm apply -> GLU (Linear64->64 over channels, times sigmoid of the un-projected
// input) -> Dropout(0.5) -> AvgPool(2,4), one pass over the conv output ypre [pixels][64].
//
// Replaces (reference file:line): baseline/models/CNN.py:49,56 (batchnorm, glu), :5-16 (GLU), :60-61,67 (dropout, pool).
//
// Tile = 128 consecutive pixels = 16 pool windows.  Nothing is staged through registers:
//   * the [128][64] fp32 tile arrives by TMA directly in the tensor core's K-major operand layout (two 16 KB boxes),
//   * BatchNorm is folded into the GEMM:  lin = Wg (a v + s) + bg = (Wg diag(a)) v + (Wg s + bg), so MMA1 runs on the
//     RAW tile the moment it lands (issued by the control warp, no thread pass in front of it); the folded, swizzled
//     weight image and the pooling matrix are prepared once per launch by bn_finalize and bulk-copied (25 KB),
//   * the 8 compute warps read v back from the same tile for the gate  z = lin * sigmoid(a v + s) * keep,  write z over
//     it (MN-major) and the (2,4) average pool is a second MMA with a 0/1 window matrix (as in cnn0.cu).
// Control warp (warp 8): TMA ring of two input buffers, MMA1 one tile ahead, MMA2, all through mbarriers; the compute
// warps never meet a block-wide barrier.  Two CTAs per SM.
#include "cnn.cuh"
#include "tc.cuh"
#include "tma.cuh"

namespace {

constexpr int kTile = 128;
constexpr float kTruncComp = 1.f + 3.5221e-4f;      // mean shrink of tf32 operand truncation, see cnn0.cu
constexpr float kLog2e = 1.4426950408889634f;

struct GluFwdArgs {
    long long n_pix;
    int F;                 // 16 or 4
    const float* img;      // operand image written by bn_finalize (cnn.cuh: kGluImg*)
    DropoutCfg drop;
    float* out;            // [n_pix / 8][64]
};

__device__ __forceinline__ float ex2_ftz(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_ftz(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void bar_sync_named(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ uint32_t krow_base(uint32_t region, int r) {    // K-major SWIZZLE_128B (tc::sw128_chunk)
    return region + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((r & 7) << 4));
}
__device__ __forceinline__ uint32_t mnrow_base(uint32_t region, int r) {   // MN-major SWIZZLE_128B_BASE32B
    return region + (uint32_t)(r * 128 + ((r & 3) << 5));
}
__device__ __forceinline__ uint32_t chunk_addr(uint32_t row_base, int c4) {
    return (row_base ^ (uint32_t)((c4 & 7) << 4)) + (uint32_t)(c4 >> 3) * 16384u;
}

// smem (1024-B aligned, all dynamic): Wb 16 KB | in[2] 2 x 32 KB | P 8 KB | bias'[64] | scale[64] | shift[64] |
//                                     keep_lo[2][128] | 8 mbarriers | tmem base
// TMEM (256 columns): lin[2] at 0 / 64, pooled[2] at 128 / 144 (M = 64 channels, N = 16 windows)
constexpr int kWb = 0, kIn = 16384, kP = 16384 + 65536, kMisc = kP + 8192;
constexpr int kSmemBytes = kMisc + 3 * 64 * 4 + 2 * 128 * 4 + 9 * 8 + 16;
constexpr int kThreads = 288;              // warps 0..7 compute (row = tid & 127, half = tid >> 7), warp 8 control

__global__ void __launch_bounds__(kThreads, 2)
glu_pool_fwd_tma_kernel(const __grid_constant__ CUtensorMap in_map, GluFwdArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* Wb = smem + kWb;
    unsigned char* Pm = smem + kP;
    float* bias_s = reinterpret_cast<float*>(smem + kMisc);
    float* scale_s = bias_s + 64;
    float* shift_s = scale_s + 64;
    uint32_t* keep_lo = reinterpret_cast<uint32_t*>(shift_s + 64);            // [2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(keep_lo + 256);
    uint64_t* in_full = bars;          // [2] TMA landed                      (tx)
    uint64_t* lin_full = bars + 2;     // [2] MMA1 complete                   (1)
    uint64_t* z_ready = bars + 4;      // [2] compute warps wrote z           (256)
    uint64_t* pool_full = bars + 6;    // [2] MMA2 complete: pooled tile ready, input buffer free   (1)
    uint64_t* img_bar = bars + 8;      // operand image landed (tx)
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 9);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if ((tc::smem_u32(smem) & 1023u) != 0) __trap();

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&in_full[i], 1); tc::mbar_init(&lin_full[i], 1);
            tc::mbar_init(&z_ready[i], 256); tc::mbar_init(&pool_full[i], 1);
        }
        tc::mbar_init(img_bar, 1);
        tc::fence_mbar_init();
        // W' | P | misc are laid out in the image exactly as in shared memory: Wb at 0, P at kP, misc at kMisc
        mbar_expect_tx(img_bar, kGluImgBytes);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(a.img);
        bulk_g2s(Wb, src, 16384, img_bar);
        bulk_g2s(Pm, src + kGluImgP, 8192, img_bar);
        bulk_g2s(bias_s, src + kGluImgMisc, 768, img_bar);
    }
    if (warp == 8) tc::tmem_alloc(tmem_base_s, 256);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    tc::mbar_wait(img_bar, 0);
    const uint32_t tmem = *tmem_base_s;
    const long long n_tiles = (a.n_pix + kTile - 1) / kTile;
    const long long stride = gridDim.x;
    const uint32_t in_a = tc::smem_u32(smem + kIn);

    if (warp == 8) {
        // ---------------- control warp: all lanes run the loop, one lane is elected per TMA / MMA / commit ----------------
        const uint32_t wb_lo = tc::desc_lo(tc::smem_u32(Wb), 16), p_lo = tc::desc_lo(tc::smem_u32(Pm), 16);
        const uint32_t k_hi = tc::desc_hi(1024, 2), mn_hi = tc::desc_hi(512, 1);
        constexpr uint32_t idesc1 = tc::idesc_tf32(128, 64, 0, 0), idesc2 = tc::idesc_tf32(64, 16, 1, 0);
        auto issue_tma = [&](long long tile, int buf) {
            if (lane == 0) {
                mbar_expect_tx(&in_full[buf], 32768);
                unsigned char* dst = smem + kIn + buf * 32768;
                tma_load_2d(dst, &in_map, 0, (int)(tile * kTile), &in_full[buf]);
                tma_load_2d(dst + 16384, &in_map, 32, (int)(tile * kTile), &in_full[buf]);
            }
            __syncwarp();
        };
        auto issue_mma1 = [&](int it) {          // lin = v W'^T as soon as the tile has landed
            const int buf = it & 1;
            tc::mbar_wait(&in_full[buf], (it >> 1) & 1);
            tc::fence_after_sync();
            const uint32_t a_lo = tc::desc_lo(in_a + buf * 32768, 16);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                tc::umma_tf32_elect(tmem + buf * 64, a_lo + (((j >> 2) * 16384 + (j & 3) * 32) >> 4), k_hi,
                                    wb_lo + (((j >> 2) * 8192 + (j & 3) * 32) >> 4), k_hi, idesc1, j > 0 ? 1u : 0u);
            tc::umma_commit_elect(&lin_full[buf]);
        };
        long long tile = blockIdx.x;
        if (tile < n_tiles) issue_tma(tile, 0);
        if (tile + stride < n_tiles) issue_tma(tile + stride, 1);
        if (tile < n_tiles) issue_mma1(0);
        int it = 0;
        for (; tile < n_tiles; tile += stride, ++it) {
            const int buf = it & 1;
            if (tile + stride < n_tiles) issue_mma1(it + 1);
            tc::mbar_wait(&z_ready[buf], (it >> 1) & 1);       // z of tile `it` is in the input buffer (MN-major)
            tc::fence_after_sync();
            const uint32_t z_lo = tc::desc_lo(in_a + buf * 32768, 16384);
#pragma unroll
            for (int j = 0; j < 16; ++j)                       // pooled[n][w] = sum_r z[r][n] P[w][r]
                tc::umma_tf32_elect(tmem + 128 + buf * 16, z_lo + (j * 1024 >> 4), mn_hi,
                                    p_lo + (((j >> 2) * 2048 + (j & 3) * 32) >> 4), k_hi, idesc2, j > 0 ? 1u : 0u);
            tc::umma_commit_elect(&pool_full[buf]);
            if (tile + 2 * stride < n_tiles) {
                tc::mbar_wait(&pool_full[buf], (it >> 1) & 1);  // MMA2 has read the buffer: refill it
                issue_tma(tile + 2 * stride, buf);
            }
        }
    } else {
        // ---------------- compute warps ----------------
        const int row = tid & 127, half = tid >> 7;
        const int wq = warp & 3;
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
        const bool drop = a.drop.enabled != 0;
        uint64_t seed = a.drop.seed; uint32_t step = a.drop.step;
        if (a.drop.sc) { seed = a.drop.sc->seed; step = a.drop.sc->step; }
        const float pool_scale = (drop ? 0.25f : 0.125f) * kTruncComp;   // 1/8 window, x2 inverted dropout, z truncation
        const long long n_out = a.n_pix >> 3;
        const bool round_out = a.F != 4;                        // blocks 1's output feeds the next tensor-core conv
        auto pooled_epilogue = [&](long long ptile, int pit) {
            const int pb = pit & 1;
            tc::mbar_wait(&pool_full[pb], (pit >> 1) & 1);
            tc::fence_after_sync();
            if (warp < 4) {
                float v[16];
                tc::tmem_ld16(tmem + 128 + pb * 16 + lane_base, v);
                tc::tmem_ld_wait();
                if (lane < 16) {                  // accumulator row m of an M=64 MMA lives in lane 32*(m/16) + m%16
                    const long long o0 = ptile * 16;
                    float* dst = a.out + o0 * 64 + 16 * warp + lane;
#pragma unroll
                    for (int w = 0; w < 16; ++w) {
                        const float o = pool_scale * v[w];
                        if (o0 + w < n_out) dst[w * 64] = round_out ? tc::tf32_rn(o) : o;
                    }
                }
            }
            tc::fence_before_sync();
        };
        long long prev = -1;
        int it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += stride, ++it) {
            const int buf = it & 1;
            uint32_t keep = 0xffffffffu;
            if (drop) {
                if (half == 1) {
                    const uint4 r = philox4x32_10((uint64_t)(tile * kTile + row), a.drop.stream, step, seed);
                    keep_lo[buf * 128 + row] = r.x;
                    keep = r.y;
                }
                bar_sync_named(1, 256);
                if (half == 0) keep = keep_lo[buf * 128 + row];
            }
            tc::mbar_wait(&in_full[buf], (it >> 1) & 1);
            const uint32_t v_rb = krow_base(in_a + buf * 32768, row), z_rb = mnrow_base(in_a + buf * 32768, row);
            float g[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 v = ld_shared_v4(chunk_addr(v_rb, 8 * half + q));
                const float4 sc = *reinterpret_cast<const float4*>(scale_s + 32 * half + 4 * q);    // pre-multiplied by -log2(e)
                const float4 sh = *reinterpret_cast<const float4*>(shift_s + 32 * half + 4 * q);
                g[4 * q + 0] = ex2_ftz(fmaf(sc.x, v.x, sh.x));
                g[4 * q + 1] = ex2_ftz(fmaf(sc.y, v.y, sh.y));
                g[4 * q + 2] = ex2_ftz(fmaf(sc.z, v.z, sh.z));
                g[4 * q + 3] = ex2_ftz(fmaf(sc.w, v.w, sh.w));
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) g[i] = rcp_ftz(1.f + g[i]);
            if (drop) {
#pragma unroll
                for (int i = 0; i < 32; ++i) g[i] = (keep & (1u << i)) ? g[i] : 0.f;
            }
            if (prev >= 0) pooled_epilogue(prev, it - 1);       // drains tile it - 1 while MMA1 of this tile finishes
            tc::mbar_wait(&lin_full[buf], (it >> 1) & 1);
            tc::fence_after_sync();
            {
                float lin[32];
                tc::tmem_ld16(tmem + buf * 64 + lane_base + 32 * half, lin);
                tc::tmem_ld16(tmem + buf * 64 + lane_base + 32 * half + 16, lin + 16);
                tc::tmem_ld_wait();
                tc::fence_before_sync();
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bias_s + 32 * half + 4 * q);
                    st_shared_v4(chunk_addr(z_rb, 8 * half + q),          // overwrites v (MMA1 has completed, gate done)
                                 (lin[4 * q + 0] + b4.x) * g[4 * q + 0], (lin[4 * q + 1] + b4.y) * g[4 * q + 1],
                                 (lin[4 * q + 2] + b4.z) * g[4 * q + 2], (lin[4 * q + 3] + b4.w) * g[4 * q + 3]);
                }
            }
            tc::fence_proxy_async();
            mbar_arrive(&z_ready[buf]);
            prev = tile;
        }
        if (prev >= 0) pooled_epilogue(prev, it - 1);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 8) tc::tmem_dealloc(tmem, 256);
}

}  // namespace

int glu_tma_kernels_init() {
    { const int rc = dcase_tma_init(); if (rc != DCASE_OK) return rc; }
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(glu_pool_fwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    return DCASE_OK;
}

int launch_glu_pool_fwd(const float* ypre, long long n_pix, int F, const float* glu_img, DropoutCfg drop, float* out,
                        int num_sms, cudaStream_t s) {
    DCASE_PROF(F == 16 ? "glu_pool_fwd_l1" : "glu_pool_fwd_l2", s);
    DCASE_REQUIRE(F == 16 || F == 4, "glu_pool is built for the 16- and 4-bin blocks of cfg.crnn_kwargs");
    DCASE_REQUIRE(n_pix > 0 && n_pix % (2 * F) == 0 && n_pix < (1ll << 31), "pixel count must be whole frame pairs");
    GluFwdArgs a{};
    a.n_pix = n_pix; a.F = F; a.img = glu_img; a.drop = drop; a.out = out;
    CUtensorMap in_map;
    { const int rc = make_rows_map(&in_map, ypre, n_pix, kTile, CU_TENSOR_MAP_SWIZZLE_128B); if (rc != DCASE_OK) return rc; }
    const long long n_tiles = (n_pix + kTile - 1) / kTile;
    const long long grid = n_tiles < 2ll * num_sms ? n_tiles : 2ll * num_sms;
    glu_pool_fwd_tma_kernel<<<(int)grid, kThreads, kSmemBytes, s>>>(in_map, a);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}
