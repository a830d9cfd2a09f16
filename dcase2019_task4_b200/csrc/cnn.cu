// CNN stack of the CRNN, the small kernels around the tensor-core ones (cnn0.cu, conv_tc.cu, glu_tma.cu):
// 3 x { conv3x3 + bias -> BatchNorm2d(eps 1e-3, momentum 0.99) -> GLU
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
// Thread = one mel bin f of a 16-frame segment of one clip: it walks down the frames with a rolling 3 x 3 window in
// registers (3 new loads per pixel instead of 9) and accumulates the 54 sums in fp32; fp64 from the block level on.
__device__ __forceinline__ int tri_index(int k, int l) {  // k <= l, order of cnn0_moments_kernel
    return 9 + k * 9 - (k * (k - 1)) / 2 + (l - k);
}

// ---- BN fold of block 0 from the tap moments (CNN.py:49) --------------------------------------------------------------
// Task (c, k): row k of the 9 x 9 quadratic form  var_c = sum_kl w_ck w_cl (M_kl - m_k m_l)  and term k of the mean (fp64:
// B200's fp64 rate is low, so the 81 terms of a channel are spread over nine tasks).
__device__ __forceinline__ void bn0_task(int c, int k, const double* mom, long long n_pix, const float* __restrict__ w,
                                         double (*pm)[64], double (*pv)[64]) {
    const double inv_n = 1.0 / (double)n_pix;
    const double mk = mom[k] * inv_n, wk = (double)w[c * 9 + k];
    double v = 0.0;
    for (int l = 0; l < 9; ++l) {
        const double M = mom[k <= l ? tri_index(k, l) : tri_index(l, k)] * inv_n;
        v += (double)w[c * 9 + l] * (M - mk * (mom[l] * inv_n));
    }
    pm[k][c] = wk * mk;
    pv[k][c] = wk * v;
}
// Channel c: statistics (batch: from the task partials; eval: running), running-stat update (momentum 0.99, unbiased
// variance), BatchNorm folded into the conv weights.
__device__ __forceinline__ void bn0_channel(int c, double (*pm)[64], double (*pv)[64], long long n_pix,
                                            const float* __restrict__ w, const float* __restrict__ b,
                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                            float* __restrict__ running, int training, float* __restrict__ fold0) {
    double mean, var;
    if (training) {
        mean = b[c];
        var = 0.0;
        for (int j = 0; j < 9; ++j) { mean += pm[j][c]; var += pv[j][c]; }
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
    for (int j = 0; j < 9; ++j) fold0[kFold0Wf + j * 64 + c] = a * w[c * 9 + j];
    fold0[kFold0Bf + c] = a * (b[c] - (float)mean) + beta[c];
    fold0[kFold0Mean + c] = (float)mean;
    fold0[kFold0Invstd + c] = invstd;
    fold0[kFold0A + c] = a;
}
struct Bn0FinalizeArgs {      // training: the LAST block of the moments kernel folds BatchNorm (mom[54] is its ticket)
    long long n_pix;
    const float *w, *b, *gamma, *beta;
    float *running, *fold0;
};

constexpr int kMomSeg = 18;      // 864 frames = 48 segments per clip: 288 blocks of four strips at B = 24, one wave of two CTAs per SM
constexpr int kMomPitch = 4 * 65 + 1;      // floats per row of the reduction buffer (4 strips of 64 threads, padded)
constexpr int kMomSmemBytes = 54 * kMomPitch * 4;
__global__ void __launch_bounds__(256)
cnn0_moments_kernel(const float* __restrict__ x, int B, int T, double* __restrict__ mom, Bn0FinalizeArgs fin) {
    extern __shared__ float tr[];            // [54][kMomPitch]: the block's partial sums, transposed (see the reduction below)
    __shared__ int is_last;
    __shared__ double mom_s[54], pm[9][64], pv[9][64];
    float acc[54];
#pragma unroll
    for (int i = 0; i < 54; ++i) acc[i] = 0.f;
    const int f = threadIdx.x & 63, strip = threadIdx.x >> 6;
    const int segs_per_clip = (T + kMomSeg - 1) / kMomSeg;
    const int n_seg = B * segs_per_clip;
    for (int sg = blockIdx.x * 4 + strip; sg < n_seg; sg += gridDim.x * 4) {
        const int b = sg / segs_per_clip, t0 = (sg - b * segs_per_clip) * kMomSeg;
        const float* xb = x + (long long)b * T * 64 + f;
        // all 18 x 3 values of the segment are requested up front (independent loads in flight together): with the rolling
        // three-row window every iteration waited out an L2 / HBM round trip, 16 of them in a row -- the kernel's 19 us
        float r[kMomSeg + 2][3];
#pragma unroll
        for (int k = 0; k < kMomSeg + 2; ++k) {
            const int t = t0 - 1 + k;
            const bool tv = t >= 0 && t < T && k <= (T - t0 < kMomSeg ? T - t0 : kMomSeg) + 1;
            const float* row = xb + (long long)t * 64;
            r[k][0] = (tv && f > 0) ? __ldg(row - 1) : 0.f;
            r[k][1] = tv ? __ldg(row) : 0.f;
            r[k][2] = (tv && f < 63) ? __ldg(row + 1) : 0.f;
        }
        const int n_rows = T - t0 < kMomSeg ? T - t0 : kMomSeg;
#pragma unroll
        for (int k = 0; k < kMomSeg; ++k) {
            if (k < n_rows) {
                const float tap[9] = {r[k][0], r[k][1], r[k][2], r[k + 1][0], r[k + 1][1], r[k + 1][2],
                                      r[k + 2][0], r[k + 2][1], r[k + 2][2]};                        // k = (dy+1)*3 + (dx+1)
                int i = 9;
#pragma unroll
                for (int a = 0; a < 9; ++a) {
                    acc[a] += tap[a];
#pragma unroll
                    for (int l = a; l < 9; ++l) { acc[i] = fmaf(tap[a], tap[l], acc[i]); ++i; }
                }
            }
        }
    }
    // Block reduction through shared memory: every thread parks its 54 partials (conflict-free, transposed), then thread
    // (i, q) adds a quarter of the block's 256 values of sum i and the quad folds -- ~190 instructions per thread instead
    // of the 540 of 54 five-step warp butterflies, which cost as much as the accumulation itself.
    {
        const int col = (threadIdx.x >> 6) * 65 + (threadIdx.x & 63);
#pragma unroll
        for (int i = 0; i < 54; ++i) tr[i * kMomPitch + col] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x < 224) {                 // whole warps (the quad fold shuffles): sums 54, 55 do not exist
        const int i = threadIdx.x >> 2, q = threadIdx.x & 3;
        const float* src = tr + (i < 54 ? i : 53) * kMomPitch + q * 65;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
        for (int k = 0; k < 64; k += 4) { s0 += src[k]; s1 += src[k + 1]; s2 += src[k + 2]; s3 += src[k + 3]; }
        double s = (double)(s0 + s1) + (double)(s2 + s3);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (q == 0 && i < 54) atomicAdd(mom + i, s);
    }
    // the last block to arrive folds BatchNorm into the conv weights: saves the one-block launch that sat between this
    // kernel and cnn0_fwd on the forward chain (SyncBN: the sums cross the ranks first, bn0_finalize_kernel folds)
    if (fin.fold0 == nullptr) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
        is_last = atomicAdd(reinterpret_cast<unsigned long long*>(mom + 54), 1ull) == (unsigned long long)gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x < 54) mom_s[threadIdx.x] = __ldcg(mom + threadIdx.x);     // from L2: the other blocks' atomics
    __syncthreads();
    for (int task = threadIdx.x; task < 576; task += 256) bn0_task(task & 63, task >> 6, mom_s, fin.n_pix, fin.w, pm, pv);
    __syncthreads();
    if (threadIdx.x < 64)
        bn0_channel(threadIdx.x, pm, pv, fin.n_pix, fin.w, fin.b, fin.gamma, fin.beta, fin.running, 1, fin.fold0);
}

// Eval mode (running statistics, no moments pass), or a caller that wants the fold alone: one block of 64 x 9 threads.
__global__ void __launch_bounds__(576)
bn0_finalize_kernel(const double* __restrict__ mom, long long n_pix, const float* __restrict__ w,
                    const float* __restrict__ b, const float* __restrict__ gamma, const float* __restrict__ beta,
                    float* __restrict__ running, int training, float* __restrict__ fold0, double* __restrict__ mom_copy) {
    __shared__ double pm[9][64], pv[9][64];
    const int c = threadIdx.x & 63, k = threadIdx.x >> 6;
    if (mom_copy && threadIdx.x < 54) mom_copy[threadIdx.x] = mom[threadIdx.x];   // moments computed ahead of the step: the
    if (training) bn0_task(c, k, mom, n_pix, w, pm, pv);                           // backward reads them from the workspace
    __syncthreads();
    if (k == 0) bn0_channel(c, pm, pv, n_pix, w, b, gamma, beta, running, training, fold0);
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
    unsigned char* Wm = Wb + kGluImgWm;                             // Wg[n][k]: rows n (= K of dY = DL Wg), k contiguous
    for (int i = bid * 256 + tid; i < 4096; i += nb * 256) {
        const int n = i >> 6, k = i & 63;
        *reinterpret_cast<float*>(Wm + (k >> 5) * 8192 + tc::sw128b32_chunk(n, (k & 31) >> 2) + (k & 3) * 4) = tc::tf32_rn(kComp * __ldg(glu_w + i));
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

// d_pre = a * (d_y - S1/N - xhat * S2/N)   (BatchNorm backward, batch statistics), in place.
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(float* __restrict__ d_y, const float* __restrict__ ypre, long long n_pix, long long n_stat,
                    const float* __restrict__ bn, const float* __restrict__ s12, float param_grad_scale,
                    float* __restrict__ g_gamma, float* __restrict__ g_beta, float* __restrict__ g_conv_b) {
    __shared__ float sa[64], sm[64], si[64], s1[64], s2[64];
    if (threadIdx.x < 64) {
        const int c = threadIdx.x;
        const float inv_n = 1.f / (float)n_stat;
        sa[c] = bn[kBnScale + c]; sm[c] = bn[kBnMean + c]; si[c] = bn[kBnInvstd + c];
        s1[c] = s12[c] * inv_n; s2[c] = s12[64 + c] * inv_n;
        if (blockIdx.x == 0) { g_gamma[c] = param_grad_scale * s12[64 + c]; g_beta[c] = param_grad_scale * s12[c]; g_conv_b[c] = 0.f; }
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


}  // namespace

int cnn_kernels_init() {
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(cnn0_moments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMomSmemBytes));
    return DCASE_OK;
}

int launch_cnn0_moments(const float* x, int B, int T, double* mom, const float* conv_w, const float* conv_b,
                        const float* gamma, const float* beta, float* running, float* fold0, int num_sms, cudaStream_t s) {
    // fold0 == NULL: the moments alone (the caller reduces them over the ranks, then launch_bn0_finalize)
    DCASE_PROF("cnn0_moments", s);
    DCASE_CUDA_CHECK(cudaMemsetAsync(mom, 0, 55 * sizeof(double), s));       // 54 sums + the completion ticket
    const long long n_seg = (long long)B * ((T + kMomSeg - 1) / kMomSeg);
    DCASE_REQUIRE(n_seg < (1ll << 30), "batch too large");
    long long blocks = (n_seg + 3) / 4;
    if (blocks > num_sms * 2) blocks = num_sms * 2;      // 127 registers: two CTAs per SM, one wave
    Bn0FinalizeArgs fin{(long long)B * T * 64, conv_w, conv_b, gamma, beta, running, fold0};
    cnn0_moments_kernel<<<(int)blocks, 256, kMomSmemBytes, s>>>(x, B, T, mom, fin);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int launch_bn0_finalize(const double* mom, long long n_pix, const float* conv_w, const float* conv_b,
                        const float* gamma, const float* beta, float* running, int training, float* fold0,
                        double* mom_copy, cudaStream_t s) {
    DCASE_PROF("bn0_finalize", s);
    bn0_finalize_kernel<<<1, 576, 0, s>>>(mom, n_pix, conv_w, conv_b, gamma, beta, running, training, fold0, mom_copy);
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

int launch_bn_bwd_apply(float* d_y, const float* ypre, long long n_pix, long long n_stat, const float* bn, const float* gamma,
                        const float* s12, float param_grad_scale, float* g_gamma, float* g_beta, float* g_conv_b,
                        int num_sms, cudaStream_t s) {
    DCASE_PROF("bn_bwd_apply", s);
    (void)gamma;
    long long blocks = (n_pix * 16 + 255) / 256;
    if (blocks > num_sms * 8) blocks = num_sms * 8;
    bn_bwd_apply_kernel<<<(int)blocks, 256, 0, s>>>(d_y, ypre, n_pix, n_stat, bn, s12, param_grad_scale, g_gamma, g_beta, g_conv_b);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}
