// CNN stack of the CRNN: 3 x { conv3x3 + bias -> BatchNorm2d(eps 1e-3, momentum 0.99) -> GLU
// (Linear64->64 over channels, times sigmoid of the un-projected input) -> Dropout(0.5) -> AvgPool (2,4) }.
//
// Replaces (reference file:line):  baseline/models/CNN.py:5-16 (GLU), :42-67 (block), :85-89 (forward),
// and their autograd backward.  Activations are channels-last fp32: [B, T, F, 64].
//
// Layer 0 never materialises its [B,64,864,64] conv output (340 MB): BN batch statistics come from the 9x9
// tap moments of the input (conv0 is linear in the 9 taps), BN is folded into the conv weights, and
// conv0 -> BN -> GLU -> dropout -> pool is recomputed per 128-pixel tile in both passes.
#include "cnn.cuh"
#include "tc.cuh"

namespace {

constexpr int kTile = 128;   // pixels per tile
constexpr float kBnEps = 1e-3f;
constexpr float kBnMomentum = 0.99f;

__device__ __forceinline__ void resolve_rng(const DropoutCfg& d, uint64_t& seed, uint32_t& step) {
    seed = d.seed; step = d.step;
    if (d.sc) { seed = d.sc->seed; step = d.sc->step; }
}

// ---------------------------------------------------------------------------------------------
// layer 0: tap moments  sum x_k (9) and sum x_k x_l (45, k <= l) over all output pixels
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cnn0_moments_kernel(const float* __restrict__ x, long long n_pix, int T, double* __restrict__ mom) {
    __shared__ float red[8][54];
    float acc[54];
#pragma unroll
    for (int i = 0; i < 54; ++i) acc[i] = 0.f;
    for (long long p = blockIdx.x * 256ll + threadIdx.x; p < n_pix; p += (long long)gridDim.x * 256) {
        const int f = (int)(p & 63);
        const int t = (int)((p >> 6) % T);
        float tap[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int dy = k / 3 - 1, dx = k % 3 - 1;
            const bool ok = (t + dy >= 0) && (t + dy < T) && (f + dx >= 0) && (f + dx < 64);
            tap[k] = ok ? __ldg(x + p + dy * 64 + dx) : 0.f;
        }
        int i = 9;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            acc[k] += tap[k];
#pragma unroll
            for (int l = k; l < 9; ++l) { acc[i] = fmaf(tap[k], tap[l], acc[i]); ++i; }
        }
    }
#pragma unroll
    for (int i = 0; i < 54; ++i) {
        const float v = warp_sum(acc[i]);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 54) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += (double)red[w][threadIdx.x];
        atomicAdd(mom + threadIdx.x, s);
    }
}

__device__ __forceinline__ int tri_index(int k, int l) {  // k <= l, order of cnn0_moments_kernel
    return 9 + k * 9 - (k * (k - 1)) / 2 + (l - k);
}

// One block of 64 threads: batch (or running) statistics of conv0's output from the tap moments, BN folded
// into the conv weights, running-stat update (momentum 0.99, unbiased variance), CNN.py:49.
__global__ void bn0_finalize_kernel(const double* __restrict__ mom, long long n_pix, const float* __restrict__ w,
                                    const float* __restrict__ b, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, float* __restrict__ running, int training,
                                    float* __restrict__ fold0) {
    const int c = threadIdx.x;
    double mean, var;
    if (training) {
        const double inv_n = 1.0 / (double)n_pix;
        double m[9];
        for (int k = 0; k < 9; ++k) m[k] = mom[k] * inv_n;
        mean = b[c];
        for (int k = 0; k < 9; ++k) mean += (double)w[c * 9 + k] * m[k];
        var = 0.0;
        for (int k = 0; k < 9; ++k)
            for (int l = 0; l < 9; ++l) {
                const double M = mom[k <= l ? tri_index(k, l) : tri_index(l, k)] * inv_n;
                var += (double)w[c * 9 + k] * (double)w[c * 9 + l] * (M - m[k] * m[l]);
            }
        if (var < 0.0) var = 0.0;
        if (running) {
            running[c] = (1.f - kBnMomentum) * running[c] + kBnMomentum * (float)mean;
            const double unbiased = var * (double)n_pix / (double)(n_pix - 1);
            running[64 + c] = (1.f - kBnMomentum) * running[64 + c] + kBnMomentum * (float)unbiased;
        }
    } else {
        mean = running[c];
        var = running[64 + c];
    }
    const float invstd = (float)(1.0 / sqrt(var + (double)kBnEps));
    const float a = gamma[c] * invstd;
    for (int k = 0; k < 9; ++k) fold0[kFold0Wf + k * 64 + c] = a * w[c * 9 + k];
    fold0[kFold0Bf + c] = a * (b[c] - (float)mean) + beta[c];
    fold0[kFold0Mean + c] = (float)mean;
    fold0[kFold0Invstd + c] = invstd;
    fold0[kFold0A + c] = a;
}

// BN statistics of layers 1,2 from the per-channel sum / sum of squares (one block of 256 threads; threads 0..63 own a
// channel).  Also emits the shared-memory image the GLU forward kernel bulk-copies (glu_tma.cu): BatchNorm folded into
// the GLU weights, W'[n][k] = Wg[n][k] * scale[k] in the swizzled K-major operand layout, bias' = bg + Wg shift, the
// gate's exponent coefficients, and the 0/1 pooling-window matrix of this block's geometry.
__global__ void __launch_bounds__(256)
bn_finalize_kernel(const double* __restrict__ stats, long long n_pix, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ running, int training, float* __restrict__ bn,
                   const float* __restrict__ glu_w, const float* __restrict__ glu_b, int F, float* __restrict__ img) {
    __shared__ float sc_s[64], sh_s[64];
    const int tid = threadIdx.x;
    if (tid < 64) {
        const int c = tid;
        double mean, var;
        if (training) {
            mean = stats[c] / (double)n_pix;
            var = stats[64 + c] / (double)n_pix - mean * mean;
            if (var < 0.0) var = 0.0;
            if (running && blockIdx.x == 0) {
                running[c] = (1.f - kBnMomentum) * running[c] + kBnMomentum * (float)mean;
                const double unbiased = var * (double)n_pix / (double)(n_pix - 1);
                running[64 + c] = (1.f - kBnMomentum) * running[64 + c] + kBnMomentum * (float)unbiased;
            }
        } else {
            mean = running[c];
            var = running[64 + c];
        }
        const float invstd = (float)(1.0 / sqrt(var + (double)kBnEps));
        const float a = gamma[c] * invstd;
        const float sh = beta[c] - a * (float)mean;
        if (blockIdx.x == 0) {
            bn[kBnScale + c] = a;
            bn[kBnShift + c] = sh;
            bn[kBnMean + c] = (float)mean;
            bn[kBnInvstd + c] = invstd;
        }
        sc_s[c] = a;
        sh_s[c] = sh;
    }
    __syncthreads();
    if (!img) return;
    // the image is split over the grid's blocks (each block recomputed scale / shift above)
    constexpr float kComp = 1.f + 3.5221e-4f;          // tf32 operand truncation compensation (cnn0.cu)
    constexpr float kNegLog2e = -1.4426950408889634f;
    const int nb = gridDim.x, bid = blockIdx.x;
    unsigned char* Wb = reinterpret_cast<unsigned char*>(img);
    for (int i = bid * 256 + tid; i < 4096; i += nb * 256) {
        const int n = i >> 6, k = i & 63;
        *reinterpret_cast<float*>(Wb + (k >> 5) * 8192 + tc::sw128_off(n, k & 31)) = tc::tf32_rn(kComp * __ldg(glu_w + i) * sc_s[k]);
    }
    unsigned char* Pm = Wb + kGluImgP;
    const int wpr = F >> 2;
    for (int i = bid * 256 + tid; i < 16 * 128; i += nb * 256) {   // P[w][r] = 1 if tile row r = (tr, f) lies in pool window w
        const int w = i >> 7, r = i & 127;
        const int tr = r / F, f = r - tr * F;
        *reinterpret_cast<float*>(Pm + (r >> 5) * 2048 + tc::sw128_off(w, r & 31)) = ((tr >> 1) * wpr + (f >> 2)) == w ? 1.f : 0.f;
    }
    float* misc = reinterpret_cast<float*>(Wb + kGluImgMisc);      // bias'[64] | -log2(e) scale[64] | -log2(e) shift[64]
    for (int n = bid * 8 + (tid >> 5); n < 64; n += nb * 8) {      // bias'[n] = bg[n] + sum_k Wg[n][k] shift[k], one warp per n
        const int l = tid & 31;
        float b = __ldg(glu_w + n * 64 + l) * sh_s[l] + __ldg(glu_w + n * 64 + 32 + l) * sh_s[32 + l];
        b = warp_sum(b);
        if (l == 0) misc[n] = b + __ldg(glu_b + n);
    }
    if (bid == 0 && tid < 64) {
        misc[64 + tid] = kNegLog2e * sc_s[tid];
        misc[128 + tid] = kNegLog2e * sh_s[tid];
    }
}

// ---------------------------------------------------------------------------------------------
// shared pieces of the GLU / pool kernels
// ---------------------------------------------------------------------------------------------
struct GluArgs {
    const float* src;      // L0: x [B][T][64]; else ypre [P][64]
    long long n_pix;       // P
    int T;                 // L0 only: frames per clip
    int F;                 // 64 / 16 / 4
    const float* aff;      // L0: fold0; else bn
    const float* gamma;    // bwd, layers 1,2
    const float* beta;     // bwd, layers 1,2
    const float* glu_w;
    const float* glu_b;
    DropoutCfg drop;
    float* out;            // fwd: pooled output [P/8][64]
    const float* d_out;    // bwd: grad of pooled output
    float* d_y;            // bwd layers 1,2: grad wrt BN output [P][64]
    float* stat_acc;       // bwd: layers 1,2 -> s12 [2][64]; L0 -> acc0 {S1[64], G[64][9]}
    float* g_glu_w;
    float* g_glu_b;
};

// y row (BN output) of this thread's pixel into its smem row; L0 recomputes conv0 from the 9 taps.
// Row accessors: 16-byte chunk c4 (channels 4*c4 .. 4*c4+3) of a tile row.
struct Sw128Row {    // two SW128 blocks of 128 rows (tc.cuh): K-major tcgen05 operand, rows = pixels
    unsigned char* base;
    int r;
    __device__ __forceinline__ float4* chunk(int c4) const {
        return reinterpret_cast<float4*>(base + (c4 >> 3) * 16384 + tc::sw128_chunk(r, c4 & 7));
    }
};

// Thread mapping of the GLU kernels: 256 threads per 128-pixel tile, two threads per pixel row.
//   row  = tid & 127   (warp w reads TMEM lanes 32*(w & 3) .. +31, the lanes of its row)
//   half = tid >> 7    (channels 32*half .. 32*half + 31  = chunks 8*half .. 8*half + 7)
constexpr int kThreads = 256;

// y (BN output) of this thread's 32 channels into its smem row; L0 recomputes conv0 from the 9 taps.
template <bool L0, typename RowT>
__device__ __forceinline__ void produce_y_half(const GluArgs& a, long long p, bool valid, const float* aff_s,
                                               const float* xs, RowT a_row, int row, int half, float (&tap)[9]) {
    if (L0) {
        const int tr = row >> 6, f = row & 63;
#pragma unroll
        for (int k = 0; k < 9; ++k) tap[k] = xs[(tr + k / 3) * 66 + f + (k % 3)];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int c4 = 8 * half + q;
            float4 y = *reinterpret_cast<const float4*>(aff_s + kFold0Bf + 4 * c4);
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const float4 w = *reinterpret_cast<const float4*>(aff_s + kFold0Wf + k * 64 + 4 * c4);
                y.x = fmaf(w.x, tap[k], y.x); y.y = fmaf(w.y, tap[k], y.y);
                y.z = fmaf(w.z, tap[k], y.z); y.w = fmaf(w.w, tap[k], y.w);
            }
            *a_row.chunk(c4) = tc::tf32_rn4(y);
        }
    } else {
        const float4* src = reinterpret_cast<const float4*>(a.src + p * 64);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int c4 = 8 * half + q;
            float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                const float4 v = __ldg(src + c4);
                const float4 sc = *reinterpret_cast<const float4*>(aff_s + kBnScale + 4 * c4);
                const float4 sh = *reinterpret_cast<const float4*>(aff_s + kBnShift + 4 * c4);
                y.x = fmaf(sc.x, v.x, sh.x); y.y = fmaf(sc.y, v.y, sh.y);
                y.z = fmaf(sc.z, v.z, sh.z); y.w = fmaf(sc.w, v.w, sh.w);
            }
            *a_row.chunk(c4) = tc::tf32_rn4(y);
        }
    }
}

// stage x rows t0-1 .. t0+2 (zero padded) of clip b for a layer-0 tile (2 rows x 64 mel bins)
__device__ __forceinline__ void load_xs(const float* __restrict__ x, long long tile, int T, float* xs) {
    const long long r0 = 2 * tile;
    const long long b = r0 / T;
    const int t0 = (int)(r0 % T);
    for (int i = threadIdx.x; i < 4 * 66; i += kThreads) {
        const int hr = i / 66, hc = i % 66;
        const int tt = t0 - 1 + hr, ff = hc - 1;
        const bool ok = tt >= 0 && tt < T && ff >= 0 && ff < 64;
        xs[i] = ok ? __ldg(x + (b * T + tt) * 64 + ff) : 0.f;
    }
}

// 32 accumulator columns [col, col + 32) of this thread's row
__device__ __forceinline__ void tmem_ld_row32(uint32_t tmem_base, int warp, int col, float (&v)[32]) {
    const uint32_t t = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col;
    tc::tmem_ld16(t, v);
    tc::tmem_ld16(t + 16, v + 16);
    tc::tmem_ld_wait();
}

// ---------------------------------------------------------------------------------------------
// forward: [conv0 | BN apply] -> GLU -> dropout -> avg-pool (2,4)
// ---------------------------------------------------------------------------------------------
// smem (1024-B aligned): Wb  K-major B operand  Wg[n][k]     2 x [64][128 B]   16 KB
//                        A   K-major A operand  y[p][k]      2 x [128][128 B]  32 KB (reused for z)
//                        bg[64] | aff (fold0 / bn) | xs[4][66]
// The 64x64 channel GEMM  lin = y Wg^T  runs on the tensor core (tcgen05.mma kind::tf32, fp32 accumulate in
// TMEM, 64 columns); CUDA cores do conv0 / BN, the gate, dropout and the pooling.  Four CTAs (32 warps) per SM
// overlap one CTA's MMA with the others' CUDA-core phases.
constexpr int kGluFwdSmemBytes = 1024 + 16384 + 32768 + (64 + 832 + 4 * 66) * 4;

// Layer 0 also runs conv0 (+ folded BN) on the tensor core: per tile the 9 taps of every pixel (+ a constant 1 for
// the bias) form a [128][16] operand T0, the folded weights a [64][16] operand W0; both live in the second 16 KB
// block of A (T0 = logical columns 0..15, W0 = logical columns 16..31 of rows 0..63, same swizzle) until y = T0 W0^T
// has been read back from TMEM and written over them as the A operand of the GLU GEMM.
template <bool L0>
__global__ void __launch_bounds__(kThreads, 4)
glu_pool_fwd_kernel(GluArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* Wb = smem;
    unsigned char* A = smem + 16384;
    float* bg = reinterpret_cast<float*>(smem + 16384 + 32768);
    float* aff_s = bg + 64;           // fold0 (832) or bn (256)
    float* xs = aff_s + 832;          // [4][66] (L0)
    __shared__ uint64_t mma_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ uint2 keep_s[kTile];   // dropout keep words of the tile's pixels (computed once per pixel)
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & 127, half = tid >> 7;

    for (int i = tid; i < 4096; i += kThreads) {      // Wg[n][k] -> block k/32, row n, swizzled
        const int n = i >> 6, k = i & 63;
        *reinterpret_cast<float*>(Wb + (k >> 5) * 8192 + tc::sw128_off(n, k & 31)) = tc::tf32_rn(__ldg(a.glu_w + i));
    }
    if (tid < 64) bg[tid] = __ldg(a.glu_b + tid);
    for (int i = tid; i < (L0 ? kFold0Size : kBnSize); i += kThreads) aff_s[i] = a.aff[i];
    if (tid == 0) { tc::mbar_init(&mma_bar, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 128);
    uint64_t seed; uint32_t step;
    resolve_rng(a.drop, seed, step);
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t a_addr = tc::smem_u32(A), b_addr = tc::smem_u32(Wb);
    uint32_t phase = 0;

    // W0 chunk owned by this thread (row n = tid >> 2, logical chunk 4 + (tid & 3)): folded conv0 weights + bias
    float4 w0_chunk = make_float4(0.f, 0.f, 0.f, 0.f);
    if (L0) {
        const int n = tid >> 2, c = tid & 3;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = 4 * c + e;
            v[e] = k < 9 ? aff_s[kFold0Wf + k * 64 + n] : (k == 9 ? aff_s[kFold0Bf + n] : 0.f);
        }
        w0_chunk = tc::tf32_rn4(make_float4(v[0], v[1], v[2], v[3]));
    }
    unsigned char* T0 = A + 16384;                    // block 1 of A

    const long long n_tiles = (a.n_pix + kTile - 1) / kTile;
    const long long n_out = a.n_pix >> 3;
    const int wpr = a.F >> 2;
    const float pool_scale = a.drop.enabled ? 0.25f : 0.125f;   // 1/8 window, x2 inverted dropout
    const Sw128Row a_row{A, row};
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long p = tile * kTile + row;
        const bool valid = p < a.n_pix;
        if (L0) {
            load_xs(a.src, tile, a.T, xs);
            __syncthreads();
            // operands of y = T0 W0^T
            *reinterpret_cast<float4*>(T0 + tc::sw128_chunk(tid >> 2, 4 + (tid & 3))) = w0_chunk;
            if (half == 0) {
                const int tr = row >> 6, f = row & 63;
                float tap[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) tap[k] = tc::tf32_rn(xs[(tr + k / 3) * 66 + f + (k % 3)]);
                *reinterpret_cast<float4*>(T0 + tc::sw128_chunk(row, 0)) = make_float4(tap[0], tap[1], tap[2], tap[3]);
                *reinterpret_cast<float4*>(T0 + tc::sw128_chunk(row, 1)) = make_float4(tap[4], tap[5], tap[6], tap[7]);
                *reinterpret_cast<float4*>(T0 + tc::sw128_chunk(row, 2)) = make_float4(tap[8], 1.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(T0 + tc::sw128_chunk(row, 3)) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            tc::fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                tc::fence_after_sync();
                constexpr uint32_t idesc = tc::idesc_tf32(128, 64, 0, 0);
                const uint32_t t0a = a_addr + 16384;
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    tc::umma_tf32(tmem + 64, tc::smem_desc_sw128(t0a + k * 32, 16, 1024),
                                  tc::smem_desc_sw128(t0a + 64 + k * 32, 16, 1024), idesc, k);
                tc::umma_commit(&mma_bar);
            }
        }
        if (a.drop.enabled && half == 0) {            // one Philox call per pixel (overlaps the conv0 MMA)
            const uint4 r = philox4x32_10((uint64_t)p, a.drop.stream, step, seed);
            keep_s[row] = make_uint2(r.x, r.y);
        }
        if (L0) {
            tc::mbar_wait(&mma_bar, phase);
            phase ^= 1;
            tc::fence_after_sync();
            float y[32];
            tmem_ld_row32(tmem, warp, 64 + 32 * half, y);
            tc::fence_before_sync();
#pragma unroll
            for (int q = 0; q < 8; ++q)
                *a_row.chunk(8 * half + q) = tc::tf32_rn4(make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]));
        } else {
            float tap[9];
            produce_y_half<false>(a, p, valid, aff_s, xs, a_row, row, half, tap);
        }
        tc::fence_proxy_async();                 // y tile -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            tc::umma_128x64x64_kmajor(tmem, a_addr, b_addr, false);
            tc::umma_commit(&mma_bar);
        }
        uint32_t keep = 0xffffffffu;             // keep bits of channels 32*half .. +31
        if (a.drop.enabled) { const uint2 kw = keep_s[row]; keep = half ? kw.y : kw.x; }
        tc::mbar_wait(&mma_bar, phase);
        phase ^= 1;
        tc::fence_after_sync();
        float acc[32];
        tmem_ld_row32(tmem, warp, 32 * half, acc);
        tc::fence_before_sync();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float4* slot = a_row.chunk(8 * half + q);
            const float4 y = *slot;
            const float4 b4 = *reinterpret_cast<const float4*>(bg + 32 * half + 4 * q);
            const uint32_t bits = keep >> (4 * q);
            float4 z;
            z.x = (bits & 1u) ? (acc[4 * q + 0] + b4.x) * sigmoid_fast(y.x) : 0.f;
            z.y = (bits & 2u) ? (acc[4 * q + 1] + b4.y) * sigmoid_fast(y.y) : 0.f;
            z.z = (bits & 4u) ? (acc[4 * q + 2] + b4.z) * sigmoid_fast(y.z) : 0.f;
            z.w = (bits & 8u) ? (acc[4 * q + 3] + b4.w) * sigmoid_fast(y.w) : 0.f;
            if (!valid) z = make_float4(0.f, 0.f, 0.f, 0.f);
            *slot = z;
        }
        __syncthreads();
        {   // pooling: 16 windows x 16 channel quads
            const int w = tid >> 4, cq = tid & 15;
            const int wr = w / wpr, wc = w - wr * wpr;
            const int r0 = (2 * wr) * a.F + 4 * wc;
            float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const Sw128Row rr{A, r0 + i * a.F + j};
                    const float4 u = *rr.chunk(cq);
                    s0.x += u.x; s0.y += u.y; s0.z += u.z; s0.w += u.w;
                }
            const long long op = tile * 16 + w;
            if (op < n_out) {
                float4 o = make_float4(pool_scale * s0.x, pool_scale * s0.y, pool_scale * s0.z, pool_scale * s0.w);
                if (a.F != 4) o = tc::tf32_rn4(o);            // input of the next block's tensor-core conv
                *reinterpret_cast<float4*>(a.out + op * 64 + 4 * cq) = o;
            }
        }
        __syncthreads();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

// ---------------------------------------------------------------------------------------------
// backward of [conv0 | BN apply] -> GLU -> dropout -> pool, recomputing the forward per tile
// ---------------------------------------------------------------------------------------------
// All five contractions of this pass run on the tensor core (tcgen05.mma kind::tf32, fp32 accumulators in TMEM):
//   G1  lin  [p][n]  = sum_k Y[p][k]  Wg[n][k]            (recompute)            D1: cols   0.. 63, M=128
//   G2  dY   [p][k] += sum_n DL[p][n] Wg[n][k]            (through the linear)   D2: cols  64..127, M=128
//   G3  dWg' [n][j]  = sum_p DL[p][n] [Y | 1][p][j]       (dWg and db_g)         D3: cols 128..207, M=64, N=80
//   G4  S    [c][j]  = sum_p dY[p][c] [1 | taps][p][j]    (dbeta; L0: conv0 dW)  D4: cols 208..223, M=64, N=16
//   G5  C    [c][k]  = sum_p dY[p][c] Y[p][k]             (diag -> dgamma)       D5: cols 224..287, M=64 (layers 1,2)
// G3..G5 accumulate in TMEM over all tiles of the persistent CTA and are read once at the end.  Reductions over
// pixels need MN-major operands (SWIZZLE_128B_BASE32B for fp32), so Y / DL / dY are staged in both layouts.
struct B32Row {      // MN-major operand blocks of 128 rows: rows = pixels (the K index), 32 channels per block
    unsigned char* base;
    int r;
    __device__ __forceinline__ float4* chunk(int c4) const {
        return reinterpret_cast<float4*>(base + (c4 >> 3) * 16384 + tc::sw128b32_chunk(r, c4 & 7));
    }
};

constexpr int kBwdWb = 0, kBwdWm = 16384, kBwdP = 32768, kBwdQ1 = 65536, kBwdExt = 65536 + 32768, kBwdQ2 = 114688,
              kBwdMisc = 147456;
constexpr int kGluBwdSmemBytes = 1024 + kBwdMisc + (64 + 832 + 128 + 4 * 66) * 4;

template <bool L0>
__global__ void __launch_bounds__(kThreads, 1)
glu_pool_bwd_kernel(GluArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* Wb = smem + kBwdWb;     // Wg[n][k], K-major SW128 (rows n)
    unsigned char* Wm = smem + kBwdWm;     // Wg[n][k], MN-major B32 (rows n = K index of G2, k contiguous)
    unsigned char* P = smem + kBwdP;       // K-major: Y, then DL
    unsigned char* Q1 = smem + kBwdQ1;     // MN-major: Y (2 blocks) | EXT block
    unsigned char* EXT = smem + kBwdExt;
    unsigned char* Q2 = smem + kBwdQ2;     // MN-major: DL, then dY
    float* bg = reinterpret_cast<float*>(smem + kBwdMisc);
    float* aff_s = bg + 64;                // fold0 (832) or bn (256)
    float* gb = aff_s + 832;               // [2][64]: 1/gamma, beta (layers 1,2)
    float* xs = gb + 128;                  // [4][66] (L0)
    __shared__ uint64_t mma_bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = tid & 127, half = tid >> 7;

    for (int i = tid; i < 4096; i += kThreads) {
        const int n = i >> 6, k = i & 63;
        const float w = tc::tf32_rn(__ldg(a.glu_w + i));
        *reinterpret_cast<float*>(Wb + (k >> 5) * 8192 + tc::sw128_off(n, k & 31)) = w;
        *reinterpret_cast<float*>(Wm + (k >> 5) * 8192 + tc::sw128b32_chunk(n, (k & 31) >> 2) + (k & 3) * 4) = w;
    }
    if (tid < 64) {
        bg[tid] = __ldg(a.glu_b + tid);
        if (!L0) {
            const float g = __ldg(a.gamma + tid);
            gb[tid] = fabsf(g) > 1e-20f ? 1.f / g : 0.f;
            gb[64 + tid] = __ldg(a.beta + tid);
        }
    }
    for (int i = tid; i < (L0 ? kFold0Size : kBnSize); i += kThreads) aff_s[i] = a.aff[i];
    if (half == 0) {   // EXT row of this pixel: [1, 0...] (L0 rewrites it with the taps every tile)
        const B32Row ext{EXT, row};
        *ext.chunk(0) = make_float4(1.f, 0.f, 0.f, 0.f);
        *ext.chunk(1) = make_float4(0.f, 0.f, 0.f, 0.f);
        *ext.chunk(2) = make_float4(0.f, 0.f, 0.f, 0.f);
        *ext.chunk(3) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid == 0) { tc::mbar_init(&mma_bar, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    uint64_t seed; uint32_t step;
    resolve_rng(a.drop, seed, step);
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t wb_a = tc::smem_u32(Wb), wm_a = tc::smem_u32(Wm), p_a = tc::smem_u32(P), q1_a = tc::smem_u32(Q1),
                   ext_a = tc::smem_u32(EXT), q2_a = tc::smem_u32(Q2);
    uint32_t phase = 0;
    bool pending = false;      // G4/G5 of the previous tile still reading Q1 / Q2
    bool first = true;

    const long long n_tiles = (a.n_pix + kTile - 1) / kTile;
    const int wpr = a.F >> 2;
    const Sw128Row p_row{P, row};
    const B32Row y32_row{Q1, row}, q2_row{Q2, row}, ext_row{EXT, row};

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long p = tile * kTile + row;
        const bool valid = p < a.n_pix;
        if (L0) { load_xs(a.src, tile, a.T, xs); }
        if (pending) { tc::mbar_wait(&mma_bar, phase); phase ^= 1; pending = false; }
        if (L0) __syncthreads();
        float tap[9];
        produce_y_half<L0>(a, p, valid, aff_s, xs, p_row, row, half, tap);
#pragma unroll
        for (int q = 0; q < 8; ++q) *y32_row.chunk(8 * half + q) = *p_row.chunk(8 * half + q);
        if (L0 && half == 0) {
            *ext_row.chunk(0) = tc::tf32_rn4(make_float4(1.f, tap[0], tap[1], tap[2]));
            *ext_row.chunk(1) = tc::tf32_rn4(make_float4(tap[3], tap[4], tap[5], tap[6]));
            *ext_row.chunk(2) = tc::tf32_rn4(make_float4(tap[7], tap[8], 0.f, 0.f));
        }
        tc::fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            tc::umma_128x64x64_kmajor(tmem, p_a, wb_a, false);            // G1
            tc::umma_commit(&mma_bar);
        }
        // pooled-output gradient of this pixel's window, dropout mask and 1/8 folded in (overlaps G1)
        float dz[32];
        {
            uint32_t keep = 0xffffffffu;
            float scale = 0.125f;
            if (a.drop.enabled) {
                const uint4 r = philox4x32_10((uint64_t)p, a.drop.stream, step, seed);
                keep = half ? r.y : r.x; scale = 0.25f;
            }
            const int tr = row / a.F, f = row - tr * a.F;
            const long long op = tile * 16 + (tr >> 1) * wpr + (f >> 2);
            const float4* dsrc = reinterpret_cast<const float4*>(a.d_out + op * 64) + 8 * half;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 d = valid ? __ldg(dsrc + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                const uint32_t bits = keep >> (4 * q);
                dz[4 * q + 0] = (bits & 1u) ? d.x * scale : 0.f;
                dz[4 * q + 1] = (bits & 2u) ? d.y * scale : 0.f;
                dz[4 * q + 2] = (bits & 4u) ? d.z * scale : 0.f;
                dz[4 * q + 3] = (bits & 8u) ? d.w * scale : 0.f;
            }
        }
        tc::mbar_wait(&mma_bar, phase);
        phase ^= 1;
        tc::fence_after_sync();
        float acc[32];
        tmem_ld_row32(tmem, warp, 32 * half, acc);                            // lin (without bias)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float4* slot = p_row.chunk(8 * half + q);
            const float4 y = *slot;
            const float4 b4 = *reinterpret_cast<const float4*>(bg + 32 * half + 4 * q);
            const float gx = sigmoid_fast(y.x), gy = sigmoid_fast(y.y), gz = sigmoid_fast(y.z), gw = sigmoid_fast(y.w);
            const float4 dl = tc::tf32_rn4(make_float4(dz[4 * q] * gx, dz[4 * q + 1] * gy, dz[4 * q + 2] * gz, dz[4 * q + 3] * gw));
            *slot = dl;                       // P: Y -> DL (G1 has completed)
            *q2_row.chunk(8 * half + q) = dl;
            // direct path through the gate: dz * lin * g * (1 - g)
            acc[4 * q + 0] = dz[4 * q + 0] * (acc[4 * q + 0] + b4.x) * gx * (1.f - gx);
            acc[4 * q + 1] = dz[4 * q + 1] * (acc[4 * q + 1] + b4.y) * gy * (1.f - gy);
            acc[4 * q + 2] = dz[4 * q + 2] * (acc[4 * q + 2] + b4.z) * gz * (1.f - gz);
            acc[4 * q + 3] = dz[4 * q + 3] * (acc[4 * q + 3] + b4.w) * gw * (1.f - gw);
        }
        tc::fence_proxy_async();
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            {   // G2: D2[p][k] = sum_n DL[p][n] Wg[n][k];  A K-major (P), B MN-major (Wm)
                constexpr uint32_t idesc = tc::idesc_tf32(128, 64, 0, 1);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    tc::umma_tf32(tmem + 64, tc::smem_desc_sw128(p_a + (j >> 2) * 16384 + (j & 3) * 32, 16, 1024),
                                  tc::smem_desc(wm_a + j * 1024, 8192, 512, 1), idesc, j > 0 ? 1u : 0u);
            }
            {   // G3: D3[n][j] (+)= sum_p DL[p][n] [Y | EXT][p][j]
                constexpr uint32_t idesc = tc::idesc_tf32(64, 80, 1, 1);
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    tc::umma_tf32(tmem + 128, tc::smem_desc(q2_a + j * 1024, 16384, 512, 1),
                                  tc::smem_desc(q1_a + j * 1024, 16384, 512, 1), idesc, (!first || j > 0) ? 1u : 0u);
            }
            tc::umma_commit(&mma_bar);
        }
        tc::mbar_wait(&mma_bar, phase);
        phase ^= 1;
        tc::fence_after_sync();
        {
            float d2[32];
            tmem_ld_row32(tmem, warp, 64 + 32 * half, d2);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 v = make_float4(acc[4 * q] + d2[4 * q], acc[4 * q + 1] + d2[4 * q + 1],
                                             acc[4 * q + 2] + d2[4 * q + 2], acc[4 * q + 3] + d2[4 * q + 3]);
                *q2_row.chunk(8 * half + q) = tc::tf32_rn4(v);        // Q2: DL -> dY (G3 has completed)
                if (!L0 && valid) reinterpret_cast<float4*>(a.d_y + p * 64)[8 * half + q] = v;
            }
        }
        tc::fence_proxy_async();
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            {   // G4: D4[c][j] (+)= sum_p dY[p][c] EXT[p][j]
                constexpr uint32_t idesc = tc::idesc_tf32(64, 16, 1, 1);
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    tc::umma_tf32(tmem + 208, tc::smem_desc(q2_a + j * 1024, 16384, 512, 1),
                                  tc::smem_desc(ext_a + j * 1024, 16384, 512, 1), idesc, (!first || j > 0) ? 1u : 0u);
            }
            if (!L0) {   // G5: D5[c][k] (+)= sum_p dY[p][c] Y[p][k]
                constexpr uint32_t idesc = tc::idesc_tf32(64, 64, 1, 1);
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    tc::umma_tf32(tmem + 224, tc::smem_desc(q2_a + j * 1024, 16384, 512, 1),
                                  tc::smem_desc(q1_a + j * 1024, 16384, 512, 1), idesc, (!first || j > 0) ? 1u : 0u);
            }
            tc::umma_commit(&mma_bar);
        }
        pending = true;
        first = false;
    }
    if (pending) { tc::mbar_wait(&mma_bar, phase); phase ^= 1; }
    tc::fence_after_sync();
    if (!first && warp < 4) {
        // accumulator row m of an M=64 MMA lives in TMEM lane 32*(m/16) + m%16: warp w, lanes 0..15 -> m = 16w + lane
        const int m = 16 * warp + lane;
        const bool own = lane < 16;
        const uint32_t tbase = tmem + ((uint32_t)(warp * 32) << 16);
        float v[16];
#pragma unroll 1
        for (int j0 = 0; j0 < 80; j0 += 16) {
            tc::tmem_ld16(tbase + 128 + j0, v);
            tc::tmem_ld_wait();
            if (own) {
                if (j0 < 64) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) atomicAdd(a.g_glu_w + m * 64 + j0 + j, v[j]);
                } else {
                    atomicAdd(a.g_glu_b + m, v[0]);
                }
            }
        }
        tc::tmem_ld16(tbase + 208, v);
        tc::tmem_ld_wait();
        const float s1 = v[0];
        if (own) {
            atomicAdd(a.stat_acc + m, s1);                                   // S1 = sum dY
            if (L0) {
#pragma unroll
                for (int k = 0; k < 9; ++k) atomicAdd(a.stat_acc + 64 + m * 9 + k, v[1 + k]);
            }
        }
        if (!L0) {
            float diag = 0.f;
#pragma unroll 1
            for (int j0 = 0; j0 < 64; j0 += 16) {
                tc::tmem_ld16(tbase + 224 + j0, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) diag = (j0 + j == m) ? v[j] : diag;
            }
            if (own) atomicAdd(a.stat_acc + 64 + m, (diag - gb[64 + m] * s1) * gb[m]);   // S2 = sum dY * xhat
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// d_pre = a * (d_y - S1/N - xhat * S2/N)   (BatchNorm backward, batch statistics), in place.
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(float* __restrict__ d_y, const float* __restrict__ ypre, long long n_pix,
                    const float* __restrict__ bn, const float* __restrict__ s12, float* __restrict__ g_gamma,
                    float* __restrict__ g_beta, float* __restrict__ g_conv_b) {
    __shared__ float sa[64], sm[64], si[64], s1[64], s2[64];
    if (threadIdx.x < 64) {
        const int c = threadIdx.x;
        const float inv_n = 1.f / (float)n_pix;
        sa[c] = bn[kBnScale + c]; sm[c] = bn[kBnMean + c]; si[c] = bn[kBnInvstd + c];
        s1[c] = s12[c] * inv_n; s2[c] = s12[64 + c] * inv_n;
        if (blockIdx.x == 0) { g_gamma[c] = s12[64 + c]; g_beta[c] = s12[c]; g_conv_b[c] = 0.f; }
    }
    __syncthreads();
    const long long total = n_pix * 16;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int c = (int)(i & 15) * 4;
        float4 d = reinterpret_cast<float4*>(d_y)[i];
        const float4 y = __ldg(reinterpret_cast<const float4*>(ypre) + i);
        d.x = sa[c + 0] * (d.x - s1[c + 0] - (y.x - sm[c + 0]) * si[c + 0] * s2[c + 0]);
        d.y = sa[c + 1] * (d.y - s1[c + 1] - (y.y - sm[c + 1]) * si[c + 1] * s2[c + 1]);
        d.z = sa[c + 2] * (d.z - s1[c + 2] - (y.z - sm[c + 2]) * si[c + 2] * s2[c + 2]);
        d.w = sa[c + 3] * (d.w - s1[c + 3] - (y.w - sm[c + 3]) * si[c + 3] * s2[c + 3]);
        reinterpret_cast<float4*>(d_y)[i] = tc::tf32_rn4(d);       // operand of the conv dgrad / wgrad MMAs
    }
}

constexpr size_t kGluFwdSmem = kGluFwdSmemBytes;
constexpr size_t kGluBwdSmem = kGluBwdSmemBytes;

}  // namespace

int cnn_kernels_init() {
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(glu_pool_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGluBwdSmem));
    return DCASE_OK;
}

int launch_cnn0_moments(const float* x, int B, int T, double* mom, int num_sms, cudaStream_t s) {
    DCASE_PROF("cnn0_moments", s);
    DCASE_CUDA_CHECK(cudaMemsetAsync(mom, 0, 54 * sizeof(double), s));
    const long long n_pix = (long long)B * T * 64;
    long long blocks = (n_pix + 255) / 256;
    if (blocks > num_sms * 4) blocks = num_sms * 4;
    cnn0_moments_kernel<<<(int)blocks, 256, 0, s>>>(x, n_pix, T, mom);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int launch_bn0_finalize(const double* mom, long long n_pix, const float* conv_w, const float* conv_b,
                        const float* gamma, const float* beta, float* running, int training, float* fold0,
                        cudaStream_t s) {
    DCASE_PROF("bn0_finalize", s);
    bn0_finalize_kernel<<<1, 64, 0, s>>>(mom, n_pix, conv_w, conv_b, gamma, beta, running, training, fold0);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

static int grid_for(long long n_tiles, int num_sms, int per_sm) {
    long long g = (long long)num_sms * per_sm;
    return (int)(n_tiles < g ? n_tiles : g);
}

int launch_bn_finalize(const double* stats, long long n_pix, const float* gamma, const float* beta, float* running,
                       int training, float* bn, const float* glu_w, const float* glu_b, int F, float* glu_img, cudaStream_t s) {
    DCASE_PROF("bn_finalize", s);
    bn_finalize_kernel<<<glu_img ? 8 : 1, 256, 0, s>>>(stats, n_pix, gamma, beta, running, training, bn, glu_w, glu_b, F, glu_img);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int launch_glu_pool_bwd(const float* ypre, long long n_pix, int F, const float* bn, const float* gamma,
                        const float* beta, const float* glu_w, const float* glu_b, DropoutCfg drop,
                        const float* d_out, float* d_y, float* s12, float* g_glu_w, float* g_glu_b, int num_sms,
                        cudaStream_t s) {
    DCASE_PROF(F == 16 ? "glu_pool_bwd_l1" : "glu_pool_bwd_l2", s);
    GluArgs a{};
    a.src = ypre; a.n_pix = n_pix; a.F = F; a.aff = bn; a.gamma = gamma; a.beta = beta;
    a.glu_w = glu_w; a.glu_b = glu_b; a.drop = drop; a.d_out = d_out; a.d_y = d_y; a.stat_acc = s12;
    a.g_glu_w = g_glu_w; a.g_glu_b = g_glu_b;
    const long long n_tiles = (n_pix + kTile - 1) / kTile;
    glu_pool_bwd_kernel<false><<<grid_for(n_tiles, num_sms, 1), kThreads, kGluBwdSmem, s>>>(a);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int launch_bn_bwd_apply(float* d_y, const float* ypre, long long n_pix, const float* bn, const float* gamma,
                        const float* s12, float* g_gamma, float* g_beta, float* g_conv_b, int num_sms,
                        cudaStream_t s) {
    DCASE_PROF("bn_bwd_apply", s);
    (void)gamma;
    long long blocks = (n_pix * 16 + 255) / 256;
    if (blocks > num_sms * 8) blocks = num_sms * 8;
    bn_bwd_apply_kernel<<<(int)blocks, 256, 0, s>>>(d_y, ypre, n_pix, bn, s12, g_gamma, g_beta, g_conv_b);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}
