// CNN block 0, fused:  conv3x3(1->64) + bias -> BatchNorm (folded) -> GLU -> Dropout(0.5) -> AvgPool(2,4)
// forward and backward, never materialising the [B,64,T,64] activation (340 MB at B = 24).
//
// Replaces (reference file:line): baseline/models/CNN.py:42-67 (block 0), :5-16 (GLU) and their autograd backward.
//
// Tile = 128 pixels = 2 frames x 64 mel bins of one clip = 16 pool windows.  Per tile, on the tensor core
// (tcgen05.mma kind::tf32, fp32 accumulators in TMEM):
//   MMA0  y   [p][c] = sum_j T[p][j] W0[c][j]         T = [9 taps | 1], W0 = [BN-folded conv weights | bias], K = 16
//   MMA1  lin [p][n] = sum_k y[p][k] Wg[n][k]         the GLU linear (K = 64)
//   MMA2  out [n][w] = sum_p z[p][n] P[w][p]          forward: the (2,4) average pool as a 0/1 matrix (K = 128)
//   MMA3  [U|S][m][j] += sum_p [DL|D2][p][m] T[p][j]  backward: every parameter gradient of the block (K = 128)
// CUDA cores only do the gate: g = sigmoid(y), dropout select, z = (lin + b) g, and in the backward
//   DL = dz g (grad wrt lin),  D2 = dz (lin + b) g (1 - g) (grad wrt y through the gate).
// Backward algebra: y is LINEAR in the 10 columns of T, so with U = DL^T T and S2 = D2^T T
//   dWg[n][k] = sum_p DL[p][n] y[p][k]  = sum_j U[n][j] W0[k][j]
//   S[c][j]   = sum_p dY[p][c] T[p][j]  = sum_n Wg[n][c] U[n][j] + S2[c][j]      (dY = DL Wg + D2)
// and S gives dbeta, dgamma and the conv0 weight gradient (cnn0_bwd_finalize).  No [pixels x 64] gradient tile is
// ever multiplied by a [pixels x 64] activation tile: one M=128, N=16 accumulator per tile stream.
//
// tf32 operands: kind::tf32 truncates fp32 operands to 10 mantissa bits, a mean relative shrink of
// 0.7213 * 2^-11 (log-uniform mantissa).  Tiles written per element (y, z, DL, D2) are stored unrounded and the
// shrink is undone in the constant operand (Wg) or in the epilogue scale; small per-pixel operands (taps) and
// weights are rounded to nearest.
#include "cnn.cuh"
#include "tc.cuh"

namespace {

constexpr int kTile = 128;
constexpr float kTruncComp = 1.f + 3.5221e-4f;
constexpr float kLog2e = 1.4426950408889634f;

struct Cnn0Args {
    const float* x;        // [B][T][64] z-scored log-mel
    int B, T;
    const float* fold0;    // bn0_finalize output (cnn.cuh)
    const float* glu_w;    // [64][64]
    const float* glu_b;    // [64]
    DropoutCfg drop;
    float* out;            // fwd: [B][T/2][16][64]
    const float* d_out;    // bwd: grad of out
    float* us;             // bwd: [128][16] accumulators {U[64][16], S2[64][16]} (zeroed by the caller)
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

struct KRow {    // K-major SW128 operand: two blocks of 128 rows x 32 channels
    unsigned char* base;
    int r;
    __device__ __forceinline__ float4* chunk(int c4) const {
        return reinterpret_cast<float4*>(base + (c4 >> 3) * 16384 + tc::sw128_chunk(r, c4 & 7));
    }
};
struct MnRow {   // MN-major (SWIZZLE_128B_BASE32B) operand: blocks of 128 rows (= K index) x 32 channels
    unsigned char* base;
    int r;
    __device__ __forceinline__ float4* chunk(int c4) const {
        return reinterpret_cast<float4*>(base + (c4 >> 3) * 16384 + tc::sw128b32_chunk(r, c4 & 7));
    }
};

// stage x rows t0-1 .. t0+2 (zero padded) of the tile's clip: xs[4][66]
__device__ __forceinline__ void load_xs(const float* __restrict__ x, long long tile, int T, float* xs, int t, int nt) {
    const long long r0 = 2 * tile;
    const long long b = r0 / T;
    const int t0 = (int)(r0 % T);
    for (int i = t; i < 4 * 66; i += nt) {
        const int hr = i / 66, hc = i - hr * 66;
        const int tt = t0 - 1 + hr, ff = hc - 1;
        const bool ok = tt >= 0 && tt < T && ff >= 0 && ff < 64;
        xs[i] = ok ? __ldg(x + (b * T + tt) * 64 + ff) : 0.f;
    }
}

// Wg[n][k] -> K-major SW128 B operand (two blocks of 64 rows), rounded to tf32 after the truncation compensation
__device__ __forceinline__ void stage_wg(const float* __restrict__ glu_w, unsigned char* Wb, int t, int nt) {
    for (int i = t; i < 4096; i += nt) {
        const int n = i >> 6, k = i & 63;
        *reinterpret_cast<float*>(Wb + (k >> 5) * 8192 + tc::sw128_off(n, k & 31)) = tc::tf32_rn(kTruncComp * __ldg(glu_w + i));
    }
}

// W0 (logical columns 16..31 of rows 0..63 of the T0 block): [wf[0..8][n], bf[n], 0...]; chunk 3 of the tap rows = 0
__device__ __forceinline__ void stage_w0(const float* __restrict__ fold0, unsigned char* T0, int t, int nt) {
    for (int i = t; i < 256; i += nt) {
        const int n = i >> 2, c = i & 3;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = 4 * c + e;
            v[e] = k < 9 ? __ldg(fold0 + kFold0Wf + k * 64 + n) : (k == 9 ? __ldg(fold0 + kFold0Bf + n) : 0.f);
        }
        *reinterpret_cast<float4*>(T0 + tc::sw128_chunk(n, 4 + c)) = tc::tf32_rn4(make_float4(v[0], v[1], v[2], v[3]));
    }
    for (int r = t; r < 128; r += nt) *reinterpret_cast<float4*>(T0 + tc::sw128_chunk(r, 3)) = make_float4(0.f, 0.f, 0.f, 0.f);
}

__device__ __forceinline__ void issue_mma0(uint32_t d_tmem, uint32_t t0_addr) {   // y = T W0^T, K = 16
    constexpr uint32_t idesc = tc::idesc_tf32(128, 64, 0, 0);
#pragma unroll
    for (int k = 0; k < 2; ++k)
        tc::umma_tf32(d_tmem, tc::smem_desc_sw128(t0_addr + k * 32, 16, 1024), tc::smem_desc_sw128(t0_addr + 64 + k * 32, 16, 1024),
                      idesc, k);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    tc::tmem_ld16(taddr, v);
    tc::tmem_ld16(taddr + 16, v + 16);
    tc::tmem_ld_wait();
}

// g[i] = sigmoid(y[i]) for 32 values: all ex2 first, then all rcp (independent MUFU streams)
__device__ __forceinline__ void sigmoid32(const float (&y)[32], float (&g)[32]) {
#pragma unroll
    for (int i = 0; i < 32; ++i) g[i] = ex2_ftz(-kLog2e * y[i]);
#pragma unroll
    for (int i = 0; i < 32; ++i) g[i] = rcp_ftz(1.f + g[i]);
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
// smem (1024-B aligned): Wb 16 KB | T0 16 KB | A 32 KB (y K-major, then z MN-major) | P 8 KB | bg, xs, keep
// TMEM (128 columns): [0,64) y then lin, [64,80) pooled output (M = 64 channels, N = 16 windows)
constexpr int kFwdWb = 0, kFwdT0 = 16384, kFwdA = 32768, kFwdP = 65536, kFwdMisc = 73728;
constexpr int kFwdSmemBytes = 1024 + kFwdMisc + 64 * 4 + 4 * 66 * 4 + 128 * 8;
constexpr int kFwdThreads = 256;

__global__ void __launch_bounds__(kFwdThreads, 3)
cnn0_fwd_kernel(Cnn0Args a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* Wb = smem + kFwdWb;
    unsigned char* T0 = smem + kFwdT0;
    unsigned char* A = smem + kFwdA;
    unsigned char* Pm = smem + kFwdP;
    float* bg = reinterpret_cast<float*>(smem + kFwdMisc);
    float* xs = bg + 64;
    uint2* keep_s = reinterpret_cast<uint2*>(xs + 4 * 66);
    __shared__ uint64_t bar_s, bar_pool_s;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = tid & 127, half = tid >> 7;

    stage_wg(a.glu_w, Wb, tid, kFwdThreads);
    stage_w0(a.fold0, T0, tid, kFwdThreads);
    for (int i = tid; i < 16 * 128; i += kFwdThreads) {    // P[w][p] = 1 if pixel p = (tr, f) lies in window w = f / 4
        const int w = i >> 7, p = i & 127;
        *reinterpret_cast<float*>(Pm + (p >> 5) * 2048 + tc::sw128_off(w, p & 31)) = (((p & 63) >> 2) == w) ? 1.f : 0.f;
    }
    if (tid < 64) bg[tid] = __ldg(a.glu_b + tid);
    if (tid == 0) { tc::mbar_init(&bar_s, 1); tc::mbar_init(&bar_pool_s, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 128);
    uint64_t seed = a.drop.seed; uint32_t step = a.drop.step;
    if (a.drop.sc) { seed = a.drop.sc->seed; step = a.drop.sc->step; }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t wb_a = tc::smem_u32(Wb), t0_a = tc::smem_u32(T0), a_a = tc::smem_u32(A), p_a = tc::smem_u32(Pm);
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t phase = 0, phase_pool = 0;

    const long long n_tiles = (long long)a.B * a.T / 2;
    // 1/8 window, x2 inverted dropout, truncation of z undone here (P is exactly 1)
    const float pool_scale = (a.drop.enabled ? 0.25f : 0.125f) * kTruncComp;
    const KRow y_row{A, row};
    const MnRow z_row{A, row};
    long long prev_tile = -1;

    auto pooled_epilogue = [&]() {     // pooled tile `prev_tile`: TMEM -> out; also frees the z buffer
        tc::mbar_wait(&bar_pool_s, phase_pool);
        phase_pool ^= 1;
        tc::fence_after_sync();
        if (warp < 4) {
            float v[16];
            tc::tmem_ld16(tmem + 64 + lane_base, v);
            tc::tmem_ld_wait();
            if (lane < 16) {               // accumulator row m of an M=64 MMA lives in lane 32*(m/16) + m%16
                float* dst = a.out + prev_tile * 16 * 64 + 16 * warp + lane;
#pragma unroll
                for (int w = 0; w < 16; ++w) dst[w * 64] = tc::tf32_rn(pool_scale * v[w]);   // conv1 MMA operand
            }
        }
        tc::fence_before_sync();
    };

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        load_xs(a.x, tile, a.T, xs, tid, kFwdThreads);
        __syncthreads();
        if (half == 0) {                   // operand rows of MMA0: [tap0..8, 1, 0...]
            const int tr = row >> 6, f = row & 63;
            float tap[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) tap[k] = tc::tf32_rn(xs[(tr + k / 3) * 66 + f + (k % 3)]);
            *reinterpret_cast<float4*>(T0 + tc::sw128_chunk(row, 0)) = make_float4(tap[0], tap[1], tap[2], tap[3]);
            *reinterpret_cast<float4*>(T0 + tc::sw128_chunk(row, 1)) = make_float4(tap[4], tap[5], tap[6], tap[7]);
            *reinterpret_cast<float4*>(T0 + tc::sw128_chunk(row, 2)) = make_float4(tap[8], 1.f, 0.f, 0.f);
        } else if (a.drop.enabled) {       // dropout keep bits of the pixel (64 channels)
            const uint4 r = philox4x32_10((uint64_t)(tile * kTile + row), a.drop.stream, step, seed);
            keep_s[row] = make_uint2(r.x, r.y);
        }
        tc::fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            issue_mma0(tmem, t0_a);
            tc::umma_commit(&bar_s);
        }
        if (prev_tile >= 0) pooled_epilogue();          // overlaps MMA0
        tc::mbar_wait(&bar_s, phase);
        phase ^= 1;
        tc::fence_after_sync();
        float g[32];
        {
            float y[32];
            tmem_ld32(tmem + lane_base + 32 * half, y);
            tc::fence_before_sync();
#pragma unroll
            for (int q = 0; q < 8; ++q)
                *y_row.chunk(8 * half + q) = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
            sigmoid32(y, g);
        }
        tc::fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            tc::umma_128x64x64_kmajor(tmem, a_a, wb_a, false);
            tc::umma_commit(&bar_s);
        }
        if (a.drop.enabled) {
            const uint2 kw = keep_s[row];
            const uint32_t keep = half ? kw.y : kw.x;
#pragma unroll
            for (int i = 0; i < 32; ++i) g[i] = (keep & (1u << i)) ? g[i] : 0.f;
        }
        tc::mbar_wait(&bar_s, phase);
        phase ^= 1;
        tc::fence_after_sync();
        {
            float lin[32];
            tmem_ld32(tmem + lane_base + 32 * half, lin);
            tc::fence_before_sync();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 b4 = *reinterpret_cast<const float4*>(bg + 32 * half + 4 * q);
                float4 z;
                z.x = (lin[4 * q + 0] + b4.x) * g[4 * q + 0];
                z.y = (lin[4 * q + 1] + b4.y) * g[4 * q + 1];
                z.z = (lin[4 * q + 2] + b4.z) * g[4 * q + 2];
                z.w = (lin[4 * q + 3] + b4.w) * g[4 * q + 3];
                *z_row.chunk(8 * half + q) = z;          // overwrites y (MMA1 has completed)
            }
        }
        tc::fence_proxy_async();
        __syncthreads();
        if (tid == 0) {                    // MMA2: pooled[n][w] = sum_p z[p][n] P[w][p];  A MN-major (z), B K-major (P)
            tc::fence_after_sync();
            constexpr uint32_t idesc = tc::idesc_tf32(64, 16, 1, 0);
#pragma unroll
            for (int j = 0; j < 16; ++j)
                tc::umma_tf32(tmem + 64, tc::smem_desc(a_a + j * 1024, 16384, 512, 1),
                              tc::smem_desc_sw128(p_a + (j >> 2) * 2048 + (j & 3) * 32, 16, 1024), idesc, j > 0 ? 1u : 0u);
            tc::umma_commit(&bar_pool_s);
        }
        prev_tile = tile;
    }
    if (prev_tile >= 0) pooled_epilogue();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

// ---------------------------------------------------------------------------------------------
// backward (recomputes the forward per tile)
// ---------------------------------------------------------------------------------------------
// One CTA per SM, 512 threads = two independent groups of 256 (named barriers), each streaming its own tiles so
// one group's CUDA-core phase overlaps the other's tensor-core round trips.
// smem (1024-B aligned): Wb 16 KB | per group: T0 16 KB | E 16 KB (MN-major [taps | 1]) | DL 32 KB (first y,
//                        K-major) | D2 32 KB (DL and D2 contiguous: one M = 128 MN-major A operand) | misc
// TMEM (512 columns): group g at 256 g: [0,64) y then lin, [64,80) accumulator {U | S2}[128][16]
constexpr int kBwdThreads = 512;
constexpr int kBwdGroupBytes = 16384 + 16384 + 32768 + 32768;
constexpr int kBwdMiscOff = 16384 + 2 * kBwdGroupBytes;
constexpr int kBwdMiscGroupFloats = 4 * 66 + 2 * 128;      // xs | keep (uint2 [128])
constexpr int kBwdSmemBytes = 1024 + kBwdMiscOff + (64 + 2 * kBwdMiscGroupFloats) * 4;

__global__ void __launch_bounds__(kBwdThreads, 1)
cnn0_bwd_kernel(Cnn0Args a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, grp = tid >> 8, gt = tid & 255;
    const int warp = tid >> 5, lane = tid & 31;
    const int row = gt & 127, half = gt >> 7;
    unsigned char* Wb = smem;
    unsigned char* gbase = smem + 16384 + grp * kBwdGroupBytes;
    unsigned char* T0 = gbase;
    unsigned char* E = gbase + 16384;
    unsigned char* DL = gbase + 32768;
    unsigned char* D2 = gbase + 65536;
    float* bg = reinterpret_cast<float*>(smem + kBwdMiscOff);
    float* xs = bg + 64 + grp * kBwdMiscGroupFloats;
    uint2* keep_s = reinterpret_cast<uint2*>(xs + 4 * 66);
    __shared__ uint64_t bar_s[2], bar_acc_s[2];
    __shared__ uint32_t tmem_base_s;

    stage_wg(a.glu_w, Wb, tid, kBwdThreads);
    stage_w0(a.fold0, T0, gt, 256);
    for (int r = gt; r < 128; r += 256)                    // E chunk 3 (columns 12..15) stays zero
        *reinterpret_cast<float4*>(E + tc::sw128b32_chunk(r, 3)) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < 64) bg[tid] = __ldg(a.glu_b + tid);
    if (gt == 0) { tc::mbar_init(&bar_s[grp], 1); tc::mbar_init(&bar_acc_s[grp], 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    uint64_t seed = a.drop.seed; uint32_t step = a.drop.step;
    if (a.drop.sc) { seed = a.drop.sc->seed; step = a.drop.sc->step; }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s + 256u * grp;
    const uint32_t wb_a = tc::smem_u32(Wb), t0_a = tc::smem_u32(T0), e_a = tc::smem_u32(E), dl_a = tc::smem_u32(DL);
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint64_t* bar = &bar_s[grp];
    uint64_t* bar_acc = &bar_acc_s[grp];
    const int bar_id = 1 + grp;
    uint32_t phase = 0, phase_acc = 0;
    bool pending = false, first = true;

    const long long n_tiles = (long long)a.B * a.T / 2;
    const float dz_scale = a.drop.enabled ? 0.25f : 0.125f;
    const KRow y_row{DL, row};
    const MnRow dl_row{DL, row}, d2_row{D2, row};

    for (long long tile = (long long)blockIdx.x * 2 + grp; tile < n_tiles; tile += (long long)gridDim.x * 2) {
        load_xs(a.x, tile, a.T, xs, gt, 256);
        if (pending) { tc::mbar_wait(bar_acc, phase_acc); phase_acc ^= 1; pending = false; }   // E / DL / D2 free again
        bar_sync_named(bar_id, 256);
        if (half == 0) {
            const int tr = row >> 6, f = row & 63;
            float tap[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) tap[k] = tc::tf32_rn(xs[(tr + k / 3) * 66 + f + (k % 3)]);
            const float4 c0 = make_float4(tap[0], tap[1], tap[2], tap[3]), c1 = make_float4(tap[4], tap[5], tap[6], tap[7]),
                         c2 = make_float4(tap[8], 1.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(T0 + tc::sw128_chunk(row, 0)) = c0;
            *reinterpret_cast<float4*>(T0 + tc::sw128_chunk(row, 1)) = c1;
            *reinterpret_cast<float4*>(T0 + tc::sw128_chunk(row, 2)) = c2;
            *reinterpret_cast<float4*>(E + tc::sw128b32_chunk(row, 0)) = c0;
            *reinterpret_cast<float4*>(E + tc::sw128b32_chunk(row, 1)) = c1;
            *reinterpret_cast<float4*>(E + tc::sw128b32_chunk(row, 2)) = c2;
        } else if (a.drop.enabled) {
            const uint4 r = philox4x32_10((uint64_t)(tile * kTile + row), a.drop.stream, step, seed);
            keep_s[row] = make_uint2(r.x, r.y);
        }
        tc::fence_proxy_async();
        bar_sync_named(bar_id, 256);
        if (gt == 0) {
            tc::fence_after_sync();
            issue_mma0(tmem, t0_a);
            tc::umma_commit(bar);
        }
        tc::mbar_wait(bar, phase);
        phase ^= 1;
        tc::fence_after_sync();
        float g[32];
        {
            float y[32];
            tmem_ld32(tmem + lane_base + 32 * half, y);
            tc::fence_before_sync();
#pragma unroll
            for (int q = 0; q < 8; ++q)
                *y_row.chunk(8 * half + q) = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
            sigmoid32(y, g);
        }
        tc::fence_proxy_async();
        bar_sync_named(bar_id, 256);
        if (gt == 0) {
            tc::fence_after_sync();
            tc::umma_128x64x64_kmajor(tmem, dl_a, wb_a, false);
            tc::umma_commit(bar);
        }
        // gradient of the pooled output for this pixel's window, dropout mask and 1/8 folded in (overlaps MMA1)
        float dz[32];
        {
            const int f = row & 63;
            const float4* dsrc = reinterpret_cast<const float4*>(a.d_out + (tile * 16 + (f >> 2)) * 64) + 8 * half;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 d = __ldg(dsrc + q);
                dz[4 * q] = d.x * dz_scale; dz[4 * q + 1] = d.y * dz_scale; dz[4 * q + 2] = d.z * dz_scale; dz[4 * q + 3] = d.w * dz_scale;
            }
            if (a.drop.enabled) {
                const uint2 kw = keep_s[row];
                const uint32_t keep = half ? kw.y : kw.x;
#pragma unroll
                for (int i = 0; i < 32; ++i) dz[i] = (keep & (1u << i)) ? dz[i] : 0.f;
            }
        }
        tc::mbar_wait(bar, phase);
        phase ^= 1;
        tc::fence_after_sync();
        {
            float lin[32];
            tmem_ld32(tmem + lane_base + 32 * half, lin);
            tc::fence_before_sync();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 b4 = *reinterpret_cast<const float4*>(bg + 32 * half + 4 * q);
                const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
                float dl[4], d2[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int i = 4 * q + e;
                    const float t = lin[i] + bb[e];
                    dl[e] = dz[i] * g[i];                        // grad wrt lin
                    d2[e] = dl[e] * fmaf(-t, g[i], t);           // grad wrt y through the gate: dz t g (1 - g)
                }
                *dl_row.chunk(8 * half + q) = make_float4(dl[0], dl[1], dl[2], dl[3]);   // overwrites y (MMA1 done)
                *d2_row.chunk(8 * half + q) = make_float4(d2[0], d2[1], d2[2], d2[3]);
            }
        }
        tc::fence_proxy_async();
        bar_sync_named(bar_id, 256);
        if (gt == 0) {                     // MMA3: {U | S2}[m][j] += sum_p [DL | D2][p][m] E[p][j]
            tc::fence_after_sync();
            constexpr uint32_t idesc = tc::idesc_tf32(128, 16, 1, 1);
#pragma unroll
            for (int j = 0; j < 16; ++j)
                tc::umma_tf32(tmem + 64, tc::smem_desc(dl_a + j * 1024, 16384, 512, 1), tc::smem_desc(e_a + j * 1024, 16384, 512, 1),
                              idesc, (!first || j > 0) ? 1u : 0u);
            tc::umma_commit(bar_acc);
        }
        pending = true;
        first = false;
    }
    if (pending) { tc::mbar_wait(bar_acc, phase_acc); phase_acc ^= 1; }
    tc::fence_after_sync();
    if (!first && (warp & 7) < 4) {        // warps 0..3 of each group: accumulator rows 32 q .. 32 q + 31
        float v[16];
        tc::tmem_ld16(tmem + 64 + lane_base, v);
        tc::tmem_ld_wait();
        const int m = 32 * (warp & 3) + lane;
#pragma unroll
        for (int j = 0; j < 10; ++j) atomicAdd(a.us + m * 16 + j, kTruncComp * v[j]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base_s, 512);
}

// One block of 256 threads: {U, S2} -> GLU parameter gradients and S = Wg^T U + S2; then (threads 0..63) the
// BatchNorm / conv0 parameter gradients from S and the tap moments.
__global__ void __launch_bounds__(256)
cnn0_bwd_finalize_kernel(const double* __restrict__ mom, long long n_pix, const float* __restrict__ w,
                         const float* __restrict__ b, const float* __restrict__ fold0, const float* __restrict__ glu_w,
                         const float* __restrict__ us, float* __restrict__ g_w, float* __restrict__ g_b,
                         float* __restrict__ g_gamma, float* __restrict__ g_beta, float* __restrict__ g_glu_w,
                         float* __restrict__ g_glu_b) {
    __shared__ float U[64][10], S[64][10], W0e[64][10];
    const int tid = threadIdx.x;
    for (int i = tid; i < 640; i += 256) {
        const int r = i / 10, j = i - r * 10;
        U[r][j] = us[r * 16 + j];
        W0e[r][j] = j < 9 ? fold0[kFold0Wf + j * 64 + r] : fold0[kFold0Bf + r];
    }
    __syncthreads();
    for (int i = tid; i < 4096; i += 256) {            // dWg[n][k] = sum_j U[n][j] W0e[k][j]
        const int n = i >> 6, k = i & 63;
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 10; ++j) s = fmaf(U[n][j], W0e[k][j], s);
        g_glu_w[i] = s;
    }
    if (tid < 64) g_glu_b[tid] = U[tid][9];
    for (int i = tid; i < 640; i += 256) {             // S[c][j] = sum_n Wg[n][c] U[n][j] + S2[c][j]
        const int c = i / 10, j = i - c * 10;
        float s = us[(64 + c) * 16 + j];
        for (int n = 0; n < 64; ++n) s = fmaf(__ldg(glu_w + n * 64 + c), U[n][j], s);
        S[c][j] = s;
    }
    __syncthreads();
    if (tid >= 64) return;
    const int c = tid;
    const double n = (double)n_pix;
    const double mean = fold0[kFold0Mean + c], invstd = fold0[kFold0Invstd + c], av = fold0[kFold0A + c];
    const double S1 = S[c][9];                          // sum dY
    double G[9], wG = 0.0;
    for (int k = 0; k < 9; ++k) { G[k] = S[c][k]; wG += (double)w[c * 9 + k] * G[k]; }   // sum dY * tap_k
    const double bm = (double)b[c] - mean;
    const double S2 = invstd * (wG + bm * S1);          // sum dY * xhat
    g_gamma[c] = (float)S2;
    g_beta[c] = (float)S1;
    g_b[c] = 0.f;                                       // BN cancels the conv bias
    for (int k = 0; k < 9; ++k) {
        double sxx = 0.0;                               // sum_p xhat_c * x_k
        for (int l = 0; l < 9; ++l) {
            const int lo = l <= k ? l : k, hi = l <= k ? k : l;
            sxx += (double)w[c * 9 + l] * mom[9 + lo * 9 - (lo * (lo - 1)) / 2 + (hi - lo)];
        }
        sxx = invstd * (sxx + bm * mom[k]);
        g_w[c * 9 + k] = (float)(av * (G[k] - (S1 / n) * mom[k] - (S2 / n) * sxx));
    }
}

}  // namespace

int cnn0_kernels_init() {
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(cnn0_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(cnn0_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmemBytes));
    return DCASE_OK;
}

int launch_cnn0_fwd(const float* x, int B, int T, const float* fold0, const float* glu_w, const float* glu_b,
                    DropoutCfg drop, float* out, int num_sms, cudaStream_t s) {
    DCASE_PROF("cnn0_fused_fwd", s);
    Cnn0Args a{};
    a.x = x; a.B = B; a.T = T; a.fold0 = fold0; a.glu_w = glu_w; a.glu_b = glu_b; a.drop = drop; a.out = out;
    const long long n_tiles = (long long)B * T / 2;
    const long long grid = n_tiles < 3ll * num_sms ? n_tiles : 3ll * num_sms;
    cnn0_fwd_kernel<<<(int)grid, kFwdThreads, kFwdSmemBytes, s>>>(a);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int launch_cnn0_bwd(const float* x, int B, int T, const float* fold0, const float* glu_w, const float* glu_b,
                    DropoutCfg drop, const float* d_out, float* us, int num_sms, cudaStream_t s) {
    DCASE_PROF("cnn0_fused_bwd", s);
    Cnn0Args a{};
    a.x = x; a.B = B; a.T = T; a.fold0 = fold0; a.glu_w = glu_w; a.glu_b = glu_b; a.drop = drop; a.d_out = d_out; a.us = us;
    const long long n_tiles = (long long)B * T / 2;
    const long long grid = (n_tiles + 1) / 2 < num_sms ? (n_tiles + 1) / 2 : num_sms;
    cnn0_bwd_kernel<<<(int)grid, kBwdThreads, kBwdSmemBytes, s>>>(a);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int launch_cnn0_bwd_finalize(const double* mom, long long n_pix, const float* conv_w, const float* conv_b,
                             const float* fold0, const float* glu_w, const float* us, float* g_conv_w, float* g_conv_b,
                             float* g_gamma, float* g_beta, float* g_glu_w, float* g_glu_b, cudaStream_t s) {
    DCASE_PROF("cnn0_bwd_finalize", s);
    cnn0_bwd_finalize_kernel<<<1, 256, 0, s>>>(mom, n_pix, conv_w, conv_b, fold0, glu_w, us, g_conv_w, g_conv_b, g_gamma,
                                             g_beta, g_glu_w, g_glu_b);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}
