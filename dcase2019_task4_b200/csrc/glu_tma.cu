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
#include <cuda_fp16.h>

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
    unsigned short* out_h; // the same as fp16 bits (operand of the next block's forward conv), nullable
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
        mbar_expect_tx(img_bar, 16384 + 8192 + 768);    // W' | P | misc (the image also holds the backward's Wm)
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
                    unsigned short* dst_h = a.out_h ? a.out_h + o0 * 64 + 16 * warp + lane : nullptr;
#pragma unroll
                    for (int w = 0; w < 16; ++w) {
                        const float o = pool_scale * v[w];
                        if (o0 + w < n_out) {
                            dst[w * 64] = round_out ? tc::tf32_rn(o) : o;
                            if (dst_h) dst_h[w * 64] = __half_as_ushort(__float2half_rn(o));
                        }
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

// ---------------------------------------------------------------------------------------------------------------
// backward of blocks 1, 2 (recomputes the forward per tile).  One CTA per SM: 8 compute warps + a control warp.
//   G1  lin [p][n] = sum_k v[p][k] W'[n][k]                       (BatchNorm folded as in the forward)     D1
//   G2  dY  [p][k] = sum_n DL[p][n] Wg[n][k]   (+ the gate path, added by the threads)                      D2
//   G3  [DLv | DLs][n][j] += sum_p DL[p][n] [v | 1][p][j]       -> dWg, db_g (affine fix-up at the end)     D3
//   G4  S1[c]            += sum_p dY[p][c]                                                                   D4
//   G5  C[c][k]          += sum_p dY[p][c] v[p][k]                -> its diagonal gives sum dY xhat          D5
// The raw tile v arrives TWICE by TMA (K-major for G1, MN-major for G3 / G5): because y = a v + s is affine per
// channel, every reduction can run on v and be corrected once per CTA, so the threads never write y.  DL (K-major)
// and later dY (MN-major) overwrite the K-major input buffer; dY leaves through a TMA store.  Inputs are double
// buffered: the next tile's loads and its G1 overlap the current tile's gradient phases.
// smem: Wb 16 | Wm 16 | vK[2] 64 | vMN[2] 64 | EXT 16 | Q2 (DL MN-major) 32 = 208 KB + misc
// TMEM (512 columns): D1 0, D2 64, D3 128 (80), D4 208 (16), D5 224 (64)
// ---------------------------------------------------------------------------------------------------------------
#ifdef DCASE_MBAR_WATCHDOG
__device__ __forceinline__ void mbar_wait_id(uint64_t* bar, uint32_t parity, int id) {
    uint32_t done = 0;
    for (long long spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(tc::smem_u32(bar)), "r"(parity) : "memory");
        if (spin > 4000000) { printf("mbar watchdog: wait %d block %d thread %d parity %u\n", id, blockIdx.x, threadIdx.x, parity); __trap(); }
    }
}
#define BWAIT(bar, parity, id) mbar_wait_id(bar, parity, id)
#else
#define BWAIT(bar, parity, id) tc::mbar_wait(bar, parity)
#endif

struct GluBwdArgs {
    long long n_pix;
    int F;
    const float* bn;       // scale, shift, mean, invstd (cnn.cuh)
    const float* img;      // W' | P | misc | Wm (cnn.cuh: kGluImg*)
    DropoutCfg drop;
    const float* d_out;    // [n_pix / 8][64]
    float* s12;            // [2][64]: sum dY, sum dY xhat (atomics)
    float* g_glu_w;        // [64][64] (atomics)
    float* g_glu_b;        // [64]
};

constexpr int kBWb = 0, kBWm = 16384, kBvK = 32768, kBvMN = kBvK + 65536, kBExt = kBvMN + 65536, kBQ2 = kBExt + 16384,
              kBMisc = kBQ2 + 32768;
constexpr int kBwdSmemBytes = kBMisc + 3 * 64 * 4 + 2 * 128 * 4 + 10 * 8 + 16;

__global__ void __launch_bounds__(kThreads, 1)
glu_pool_bwd_tma_kernel(const __grid_constant__ CUtensorMap in_k_map, const __grid_constant__ CUtensorMap in_mn_map,
                        const __grid_constant__ CUtensorMap dy_map, GluBwdArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* Wb = smem + kBWb;
    unsigned char* Wm = smem + kBWm;
    unsigned char* EXT = smem + kBExt;
    float* bias_s = reinterpret_cast<float*>(smem + kBMisc);
    float* scale_s = bias_s + 64;
    float* shift_s = scale_s + 64;
    uint32_t* keep_lo = reinterpret_cast<uint32_t*>(shift_s + 64);            // [2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(keep_lo + 256);
    uint64_t* in_full = bars;          // [2] both TMA copies of a tile landed        (tx)
    uint64_t* lin_full = bars + 2;     // G1 complete                                 (1)
    uint64_t* dl_ready = bars + 3;     // threads wrote DL (K-major + MN-major)       (256)
    uint64_t* g2_full = bars + 4;      // G2 complete                                 (1)
    uint64_t* g3_done = bars + 5;      // G3 complete: Q2 may be rewritten            (1)
    uint64_t* dy_ready = bars + 6;     // threads wrote dY (MN-major)                 (256)
    uint64_t* g45_done = bars + 7;     // G4, G5 complete: the tile's buffers are free (1)
    uint64_t* img_bar = bars + 8;
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 10);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if ((tc::smem_u32(smem) & 1023u) != 0) __trap();

    for (int r = tid; r < 128; r += kThreads) {          // EXT[p][0] = 1, other columns 0 (MN-major block, 16 columns used)
        *reinterpret_cast<float4*>(EXT + tc::sw128b32_chunk(r, 0)) = make_float4(1.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(EXT + tc::sw128b32_chunk(r, 1)) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(EXT + tc::sw128b32_chunk(r, 2)) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(EXT + tc::sw128b32_chunk(r, 3)) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) tc::mbar_init(&in_full[i], 1);
        tc::mbar_init(lin_full, 1); tc::mbar_init(dl_ready, 256); tc::mbar_init(g2_full, 1); tc::mbar_init(g3_done, 1);
        tc::mbar_init(dy_ready, 256); tc::mbar_init(g45_done, 1); tc::mbar_init(img_bar, 1);
        tc::fence_mbar_init();
        mbar_expect_tx(img_bar, 16384 + 768 + 16384);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(a.img);
        bulk_g2s(Wb, src, 16384, img_bar);
        bulk_g2s(bias_s, src + kGluImgMisc, 768, img_bar);
        bulk_g2s(Wm, src + kGluImgWm, 16384, img_bar);
    }
    if (warp == 8) tc::tmem_alloc(tmem_base_s, 512);
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    BWAIT(img_bar, 0, 1);
    const uint32_t tmem = *tmem_base_s;
    const long long n_tiles = (a.n_pix + kTile - 1) / kTile;
    const long long stride = gridDim.x;
    const uint32_t vk_a = tc::smem_u32(smem + kBvK), vmn_a = tc::smem_u32(smem + kBvMN), q2_a = tc::smem_u32(smem + kBQ2);
    const bool any = (long long)blockIdx.x < n_tiles;

    if (warp == 8) {
        const uint32_t k_hi = tc::desc_hi(1024, 2), mn_hi = tc::desc_hi(512, 1);
        const uint32_t wb_lo = tc::desc_lo(tc::smem_u32(Wb), 16), wm_lo = tc::desc_lo(tc::smem_u32(Wm), 8192);
        const uint32_t ext_lo = tc::desc_lo(tc::smem_u32(EXT), 16384), q2_lo = tc::desc_lo(q2_a, 16384);
        auto issue_tma = [&](long long tile, int buf) {
            if (lane == 0) {
                mbar_expect_tx(&in_full[buf], 65536);
                const int r0 = (int)(tile * kTile);
                tma_load_2d(smem + kBvK + buf * 32768, &in_k_map, 0, r0, &in_full[buf]);
                tma_load_2d(smem + kBvK + buf * 32768 + 16384, &in_k_map, 32, r0, &in_full[buf]);
                tma_load_2d(smem + kBvMN + buf * 32768, &in_mn_map, 0, r0, &in_full[buf]);
                tma_load_2d(smem + kBvMN + buf * 32768 + 16384, &in_mn_map, 32, r0, &in_full[buf]);
            }
            __syncwarp();
        };
        auto issue_g1 = [&](int it) {
            const int buf = it & 1;
            BWAIT(&in_full[buf], (it >> 1) & 1, 2);
            tc::fence_after_sync();
            constexpr uint32_t idesc = tc::idesc_tf32(128, 64, 0, 0);
            const uint32_t a_lo = tc::desc_lo(vk_a + buf * 32768, 16);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                tc::umma_tf32_elect(tmem, a_lo + (((j >> 2) * 16384 + (j & 3) * 32) >> 4), k_hi,
                                    wb_lo + (((j >> 2) * 8192 + (j & 3) * 32) >> 4), k_hi, idesc, j > 0 ? 1u : 0u);
            tc::umma_commit_elect(lin_full);
        };
        long long tile = blockIdx.x;
        if (tile < n_tiles) issue_tma(tile, 0);
        if (tile + stride < n_tiles) issue_tma(tile + stride, 1);
        if (tile < n_tiles) issue_g1(0);
        int it = 0;
        for (; tile < n_tiles; tile += stride, ++it) {
            const int buf = it & 1;
            const uint32_t acc1 = it > 0 ? 1u : 0u;
            const uint32_t vmn_lo = tc::desc_lo(vmn_a + buf * 32768, 16384);
            BWAIT(dl_ready, it & 1, 3);
            tc::fence_after_sync();
            {   // G2: D2[p][k] = sum_n DL[p][n] Wg[n][k];  A K-major (DL in the vK buffer), B MN-major (Wm)
                constexpr uint32_t idesc = tc::idesc_tf32(128, 64, 0, 1);
                const uint32_t a_lo = tc::desc_lo(vk_a + buf * 32768, 16);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    tc::umma_tf32_elect(tmem + 64, a_lo + (((j >> 2) * 16384 + (j & 3) * 32) >> 4), k_hi, wm_lo + (j * 1024 >> 4), mn_hi, idesc,
                                        j > 0 ? 1u : 0u);
                tc::umma_commit_elect(g2_full);
            }
            {   // G3: D3[n][j] += sum_p DL[p][n] [v | EXT][p][j];  A = Q2 (M = 64), B = vMN blocks 0, 1 then EXT (N = 80)
                constexpr uint32_t idesc64 = tc::idesc_tf32(64, 64, 1, 1), idesc16 = tc::idesc_tf32(64, 16, 1, 1);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    tc::umma_tf32_elect(tmem + 128, q2_lo + (j * 1024 >> 4), mn_hi, vmn_lo + (j * 1024 >> 4), mn_hi, idesc64, j > 0 ? 1u : acc1);
                    tc::umma_tf32_elect(tmem + 192, q2_lo + (j * 1024 >> 4), mn_hi, ext_lo + (j * 1024 >> 4), mn_hi, idesc16, j > 0 ? 1u : acc1);
                }
                tc::umma_commit_elect(g3_done);
            }
            if (tile + stride < n_tiles) issue_g1(it + 1);        // D1 has been drained (dl_ready): next tile's lin
            BWAIT(dy_ready, it & 1, 4);
            tc::fence_after_sync();
            {   // G4: D4[c][j] += sum_p dY[p][c] EXT[p][j];  G5: D5[c][k] += sum_p dY[p][c] v[p][k];  dY MN-major in the vK buffer
                constexpr uint32_t idesc64 = tc::idesc_tf32(64, 64, 1, 1), idesc16 = tc::idesc_tf32(64, 16, 1, 1);
                const uint32_t dy_lo = tc::desc_lo(vk_a + buf * 32768, 16384);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    tc::umma_tf32_elect(tmem + 208, dy_lo + (j * 1024 >> 4), mn_hi, ext_lo + (j * 1024 >> 4), mn_hi, idesc16, j > 0 ? 1u : acc1);
                    tc::umma_tf32_elect(tmem + 224, dy_lo + (j * 1024 >> 4), mn_hi, vmn_lo + (j * 1024 >> 4), mn_hi, idesc64, j > 0 ? 1u : acc1);
                }
                tc::umma_commit_elect(g45_done);
            }
            if (lane == 0) {                                      // dY tile -> global (the rows beyond n_pix are clipped)
                const int r0 = (int)(tile * kTile);
                tma_store_2d(&dy_map, smem + kBvK + buf * 32768, 0, r0);
                tma_store_2d(&dy_map, smem + kBvK + buf * 32768 + 16384, 32, r0);
                bulk_commit_group();
            }
            __syncwarp();
            if (tile + 2 * stride < n_tiles) {
                BWAIT(g45_done, it & 1, 5);                  // G4 / G5 have read the tile's buffers ...
                if (lane == 0) bulk_wait_group_read0();           // ... and so has the store
                __syncwarp();
                issue_tma(tile + 2 * stride, buf);
            }
        }
        if (lane == 0) bulk_wait_group0();
        __syncwarp();
    } else {
        const int row = tid & 127, half = tid >> 7;
        const int wq = warp & 3;
        const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
        const bool drop = a.drop.enabled != 0;
        uint64_t seed = a.drop.seed; uint32_t step = a.drop.step;
        if (a.drop.sc) { seed = a.drop.sc->seed; step = a.drop.sc->step; }
        const float dz_scale = drop ? 0.25f : 0.125f;
        const int wpr = a.F >> 2;
        const int tr = row / a.F, f = row - tr * a.F;
        const int win = (tr >> 1) * wpr + (f >> 2);               // pool window of this thread's pixel inside the tile
        const long long n_out = a.n_pix >> 3;
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
            // gradient of the pooled output for this pixel's window, dropout mask and 1/8 folded in
            float dz[32];
            {
                const long long op = tile * 16 + win;
                const float4* dsrc = reinterpret_cast<const float4*>(a.d_out + op * 64) + 8 * half;
                const bool ok = op < n_out;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 d = ok ? __ldg(dsrc + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                    dz[4 * q] = d.x * dz_scale; dz[4 * q + 1] = d.y * dz_scale; dz[4 * q + 2] = d.z * dz_scale; dz[4 * q + 3] = d.w * dz_scale;
                }
                if (drop) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) dz[i] = (keep & (1u << i)) ? dz[i] : 0.f;
                }
            }
            BWAIT(&in_full[buf], (it >> 1) & 1, 6);
            const uint32_t vk_rb = krow_base(vk_a + buf * 32768, row), dymn_rb = mnrow_base(vk_a + buf * 32768, row),
                           dlmn_rb = mnrow_base(q2_a, row);
            float g[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 v = ld_shared_v4(chunk_addr(vk_rb, 8 * half + q));
                const float4 sc = *reinterpret_cast<const float4*>(scale_s + 32 * half + 4 * q);    // pre-multiplied by -log2(e)
                const float4 sh = *reinterpret_cast<const float4*>(shift_s + 32 * half + 4 * q);
                g[4 * q + 0] = ex2_ftz(fmaf(sc.x, v.x, sh.x));
                g[4 * q + 1] = ex2_ftz(fmaf(sc.y, v.y, sh.y));
                g[4 * q + 2] = ex2_ftz(fmaf(sc.z, v.z, sh.z));
                g[4 * q + 3] = ex2_ftz(fmaf(sc.w, v.w, sh.w));
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) g[i] = rcp_ftz(1.f + g[i]);
            BWAIT(lin_full, it & 1, 7);
            tc::fence_after_sync();
            float direct[32];                                     // dz (lin + b') g (1 - g): gradient wrt y through the gate
            {
                float lin[32];
                tc::tmem_ld16(tmem + lane_base + 32 * half, lin);
                tc::tmem_ld16(tmem + lane_base + 32 * half + 16, lin + 16);
                tc::tmem_ld_wait();
                tc::fence_before_sync();
                if (it > 0) BWAIT(g3_done, (it - 1) & 1, 8);   // the previous tile's G3 has read Q2
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bias_s + 32 * half + 4 * q);
                    const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
                    float dl[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int i = 4 * q + e;
                        const float t = lin[i] + bb[e];
                        dl[e] = dz[i] * g[i];
                        direct[i] = dl[e] * fmaf(-t, g[i], t);
                    }
                    st_shared_v4(chunk_addr(vk_rb, 8 * half + q), dl[0], dl[1], dl[2], dl[3]);      // K-major, over v (G1 done)
                    st_shared_v4(chunk_addr(dlmn_rb, 8 * half + q), dl[0], dl[1], dl[2], dl[3]);   // MN-major
                }
            }
            tc::fence_proxy_async();
            mbar_arrive(dl_ready);
            BWAIT(g2_full, it & 1, 9);
            tc::fence_after_sync();
            {
                float d2[32];
                tc::tmem_ld16(tmem + 64 + lane_base + 32 * half, d2);
                tc::tmem_ld16(tmem + 64 + lane_base + 32 * half + 16, d2 + 16);
                tc::tmem_ld_wait();
                tc::fence_before_sync();
#pragma unroll
                for (int q = 0; q < 8; ++q)                        // dY, MN-major, over DL (G2 has read it)
                    st_shared_v4(chunk_addr(dymn_rb, 8 * half + q), direct[4 * q] + d2[4 * q], direct[4 * q + 1] + d2[4 * q + 1],
                                 direct[4 * q + 2] + d2[4 * q + 2], direct[4 * q + 3] + d2[4 * q + 3]);
            }
            tc::fence_proxy_async();
            mbar_arrive(dy_ready);
        }
        // ---- read-out: affine fix-ups of the reductions over v, staged in shared memory (the operand tiles are dead), then
        //      COALESCED 16-byte reductions: a warp instruction adds four full 128-byte lines.  148 CTAs add into the same
        //      4,288 floats: as scalar atomics from threads that own a row each (stride 256 B) every line received ~4,700
        //      serialised transactions and the tail cost more than the kernel's tile loop (round 2, measured).
        float* outs = reinterpret_cast<float*>(smem + kBQ2);      // [64][64] dWg | [64] db_g | [64] sum dY | [64] sum dY xhat (Q2: read by G3 only)
        if (any && warp < 4) {
            BWAIT(g45_done, (it - 1) & 1, 10);                // last commit: every MMA of the CTA has completed
            tc::fence_after_sync();
            constexpr float c2 = kTruncComp * kTruncComp;
            const int m = 16 * warp + lane;                       // M = 64: row m in lane 32 (m / 16) + m % 16
            const bool own = lane < 16;
            float v[16];
            tc::tmem_ld16(tmem + 192 + lane_base, v);             // D3 columns 64..79: col 64 = sum_p DL[p][n]
            tc::tmem_ld_wait();
            const float dbg = kTruncComp * v[0];
            if (own) outs[4096 + m] = dbg;
#pragma unroll 1
            for (int j0 = 0; j0 < 64; j0 += 16) {                 // dWg[n][k] = a_k (DL^T v)[n][k] + s_k db_g[n]
                tc::tmem_ld16(tmem + 128 + j0 + lane_base, v);
                tc::tmem_ld_wait();
                if (own) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        outs[m * 64 + j0 + j] = fmaf(__ldg(a.bn + kBnScale + j0 + j), c2 * v[j], __ldg(a.bn + kBnShift + j0 + j) * dbg);
                }
            }
            tc::tmem_ld16(tmem + 208 + lane_base, v);
            tc::tmem_ld_wait();
            const float s1 = kTruncComp * v[0];                   // sum dY
            float diag = 0.f;
#pragma unroll 1
            for (int j0 = 0; j0 < 64; j0 += 16) {
                tc::tmem_ld16(tmem + 224 + j0 + lane_base, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) diag = (j0 + j == m) ? v[j] : diag;
            }
            if (own) {
                outs[4160 + m] = s1;
                // sum dY xhat, xhat = (v - mean) invstd
                outs[4224 + m] = (c2 * diag - __ldg(a.bn + kBnMean + m) * s1) * __ldg(a.bn + kBnInvstd + m);
            }
            tc::fence_before_sync();
            bar_sync_named(2, 128);                               // warps 0..3: the staged results are complete
            auto red4 = [](float* dst, const float* src) {
                const float4 x = *reinterpret_cast<const float4*>(src);
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
            };
            for (int i = tid; i < 1024; i += 128) red4(a.g_glu_w + 4 * i, outs + 4 * i);
            if (tid < 16) red4(a.g_glu_b + 4 * tid, outs + 4096 + 4 * tid);
            else if (tid < 48) red4(a.s12 + 4 * (tid - 16), outs + 4160 + 4 * (tid - 16));
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 8) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

int glu_tma_kernels_init() {
    { const int rc = dcase_tma_init(); if (rc != DCASE_OK) return rc; }
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(glu_pool_fwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(glu_pool_bwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmemBytes));
    return DCASE_OK;
}

int launch_glu_pool_bwd(const float* ypre, long long n_pix, int F, const float* bn, const float* glu_img, DropoutCfg drop,
                        const float* d_out, float* d_y, float* s12, float* g_glu_w, float* g_glu_b, int num_sms,
                        cudaStream_t s) {
    DCASE_PROF(F == 16 ? "glu_pool_bwd_l1" : "glu_pool_bwd_l2", s);
    DCASE_REQUIRE(F == 16 || F == 4, "glu_pool is built for the 16- and 4-bin blocks of cfg.crnn_kwargs");
    DCASE_REQUIRE(n_pix > 0 && n_pix % (2 * F) == 0 && n_pix < (1ll << 31), "pixel count must be whole frame pairs");
    GluBwdArgs a{};
    a.n_pix = n_pix; a.F = F; a.bn = bn; a.img = glu_img; a.drop = drop; a.d_out = d_out; a.s12 = s12;
    a.g_glu_w = g_glu_w; a.g_glu_b = g_glu_b;
    CUtensorMap in_k_map, in_mn_map, dy_map;
    { const int rc = make_rows_map(&in_k_map, ypre, n_pix, kTile, CU_TENSOR_MAP_SWIZZLE_128B); if (rc != DCASE_OK) return rc; }
    { const int rc = make_rows_map(&in_mn_map, ypre, n_pix, kTile, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B); if (rc != DCASE_OK) return rc; }
    { const int rc = make_rows_map(&dy_map, d_y, n_pix, kTile, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B); if (rc != DCASE_OK) return rc; }
    const long long n_tiles = (n_pix + kTile - 1) / kTile;
    const long long grid = n_tiles < num_sms ? n_tiles : num_sms;
    glu_pool_bwd_tma_kernel<<<(int)grid, kThreads, kBwdSmemBytes, s>>>(in_k_map, in_mn_map, dy_map, a);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int launch_glu_pool_fwd(const float* ypre, long long n_pix, int F, const float* glu_img, DropoutCfg drop, float* out,
                        void* out_h, int num_sms, cudaStream_t s) {
    DCASE_PROF(F == 16 ? "glu_pool_fwd_l1" : "glu_pool_fwd_l2", s);
    DCASE_REQUIRE(F == 16 || F == 4, "glu_pool is built for the 16- and 4-bin blocks of cfg.crnn_kwargs");
    DCASE_REQUIRE(n_pix > 0 && n_pix % (2 * F) == 0 && n_pix < (1ll << 31), "pixel count must be whole frame pairs");
    GluFwdArgs a{};
    a.n_pix = n_pix; a.F = F; a.img = glu_img; a.drop = drop; a.out = out; a.out_h = (unsigned short*)out_h;
    CUtensorMap in_map;
    { const int rc = make_rows_map(&in_map, ypre, n_pix, kTile, CU_TENSOR_MAP_SWIZZLE_128B); if (rc != DCASE_OK) return rc; }
    const long long n_tiles = (n_pix + kTile - 1) / kTile;
    const long long grid = n_tiles < 2ll * num_sms ? n_tiles : 2ll * num_sms;
    glu_pool_fwd_tma_kernel<<<(int)grid, kThreads, kSmemBytes, s>>>(in_map, a);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}
