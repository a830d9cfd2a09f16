// K1/K2: waveform -> Hamming STFT magnitude -> Slaney mel (amplitude) -> noise / dB / top_db / z-score.
//
// Replaces (reference file:line):
//   baseline/DatasetDcase2019Task4.py:197-231   calculate_mel_spec            -> dcase_logmel_fwd
//   baseline/DataLoad.py:274-287,192-207,210-259,302-321 + utils/Scaler.py:99-105
//   (order: utils/utils.py:397-412 get_transforms)                            -> dcase_logmel_finish
//
// K1 layout: one CTA = 8 consecutive frames of one clip. The contiguous waveform span
// (7*511 + 2048 samples, reflect-padded at the clip ends) is staged once in shared memory, so the
// 4x frame overlap (2048/511) is served on-chip (HBM / L2 see each sample ~1.4x); 59 KB of shared memory per
// CTA keep 3 CTAs (24 warps) resident per SM. Two real frames are
// packed into one complex 2048-point Stockham FFT (radix 8,8,8,4) held in shared memory.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/dcase_b200.h"
#include "common.cuh"
#include "ctx.h"
#include "fft2048.cuh"

namespace {

constexpr int kNfft = 2048;
constexpr int kHop = 511;
constexpr int kBins = 1025;
constexpr int kMel = 64;
constexpr int kFramesPerCta = 8;
constexpr int kSpan = (kFramesPerCta - 1) * kHop + kNfft;  // 5625
constexpr int kSpanPad = (kSpan + 15) / 16 * 16;
constexpr int kMagPitch = 1032;

struct MelTables {
    const float* window;
    const cf32* twiddle;
    const float* mel_w;
    const int4* mel_work;    // [128] per-thread work items {bin start, weight offset, count, band}
    const int2* mel_owner;   // [64]  {first slot, slot count} per band
};

__device__ __forceinline__ int reflect_index(int s, int L) {
    if (s < 0) s = -s;
    if (s >= L) s = 2 * (L - 1) - s;
    return s;
}

template <typename WaveT>
__device__ __forceinline__ float load_sample(const WaveT* p, int i);
template <>
__device__ __forceinline__ float load_sample<float>(const float* p, int i) { return __ldg(p + i); }
template <>
__device__ __forceinline__ float load_sample<int16_t>(const int16_t* p, int i) {
    return (float)__ldg(p + i) * (1.0f / 32768.0f);  // soundfile's int16 -> float scaling
}

// sqrt(x) = x * rsqrt(x) (2 ulp), exact 0 for x = 0: the IEEE sqrtf sequence was 9 % of the kernel's instructions
__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return x > 0.f ? x * r : 0.f;
}

template <typename WaveT>
__global__ void __launch_bounds__(256, 3)
stft_mel_kernel(const WaveT* __restrict__ wave, int L, int T, MelTables tab, float* __restrict__ mel_amp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* span = reinterpret_cast<float*>(smem_raw);                 // [kSpan] (+pad)
    cf32* bufA = reinterpret_cast<cf32*>(span + kSpanPad);            // [2304] (padded, fft_pad)
    cf32* bufB = bufA + kFftPaddedSize;                               // [2304]
    __shared__ float part[256];
    const cf32* __restrict__ tw = tab.twiddle;                        // [2048], read through L1 (16 KB, hot)

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * kFramesPerCta;
    const WaveT* clip = wave + (size_t)b * L;

    const int s0 = t0 * kHop - kNfft / 2;
    if (sizeof(WaveT) == 4) {
        // float clips: asynchronous 4-byte copies straight into shared memory (all 22 per thread in flight at once)
        const uint32_t span_a = (uint32_t)__cvta_generic_to_shared(span);
        for (int i = tid; i < kSpan; i += 256) {
            const int s = reflect_index(s0 + i, L);
            if (s >= 0 && s < L)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(span_a + 4u * i), "l"(clip + s) : "memory");
            else
                span[i] = 0.f;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else {
        for (int i = tid; i < kSpan; i += 256) {
            const int s = reflect_index(s0 + i, L);
            span[i] = (s >= 0 && s < L) ? load_sample<WaveT>(clip, s) : 0.f;
        }
    }
    __syncthreads();

    // this thread's run of Slaney weights (<= 22 of them) is the same for every frame: the first 16 are loaded once per
    // CTA and kept in registers, the few longer runs read their tail through L1
    const int4 mel_e = __ldg(tab.mel_work + (tid & 127));           // {bin start, weight offset, count, band}
    float mel_wt[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) mel_wt[i] = i < mel_e.z ? __ldg(tab.mel_w + mel_e.y + i) : 0.f;

    const int n_pairs = min(kFramesPerCta, T - t0 + 1) / 2;  // frames [t0, T) in pairs (odd tail rounds up)
    for (int p = 0; p < n_pairs; ++p) {
        // pass 1 (Ns = 1): windowed load straight from the span, two real frames -> one complex signal
        {
            const float* fa = span + (2 * p) * kHop;
            const float* fb = fa + kHop;
            cf32 v[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int n = tid + r * 256;
                const float w = __ldg(tab.window + n);
                v[r] = cf32{fa[n] * w, fb[n] * w};
            }
            fft8(v);
#pragma unroll
            for (int r = 0; r < 8; ++r) bufA[fft_pad(tid * 8 + r)] = v[r];
        }
        __syncthreads();
        stockham_pass<8>(tid, 8, bufA, bufB, tw);
        __syncthreads();
        stockham_pass<8>(tid, 64, bufB, bufA, tw);
        __syncthreads();
        stockham_pass<4>(tid, 512, bufA, bufB, tw);
        stockham_pass<4>(tid + 256, 512, bufA, bufB, tw);
        __syncthreads();
        // separate the two real spectra and take magnitudes (bins 0..1024)
        float* mag = reinterpret_cast<float*>(bufA);  // [2][kMagPitch]
        for (int k = tid; k < kBins; k += 256) {
            const cf32 zk = bufB[fft_pad(k)];
            const cf32 zn = bufB[fft_pad((kNfft - k) & (kNfft - 1))];
            const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);
            const float br = 0.5f * (zk.y + zn.y), bi = 0.5f * (zn.x - zk.x);
            mag[k] = fast_sqrt(ar * ar + ai * ai);
            mag[kMagPitch + k] = fast_sqrt(br * br + bi * bi);
        }
        __syncthreads();
        // sparse Slaney mel projection (1,983 non-zeros per frame), load balanced: 128 threads per frame, each owns
        // a contiguous run of weights inside one band (host-built table); band owners combine the partials
        {
            const int fr = tid >> 7;
            const float* mg = mag + fr * kMagPitch + mel_e.x;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                if (i < mel_e.z) {                      // zero weights pad the run: mg stays inside the padded buffer
                    a0 = fmaf(mel_wt[i], mg[i], a0);
                    a1 = fmaf(mel_wt[i + 1], mg[i + 1], a1);
                    a2 = fmaf(mel_wt[i + 2], mg[i + 2], a2);
                    a3 = fmaf(mel_wt[i + 3], mg[i + 3], a3);
                }
            }
            for (int i = 16; i < mel_e.z; ++i) a0 = fmaf(__ldg(tab.mel_w + mel_e.y + i), mg[i], a0);
            part[tid] = (a0 + a1) + (a2 + a3);
        }
        __syncthreads();
        if ((tid & 127) < kMel) {
            const int fr = tid >> 7, m = tid & 127;
            const int2 o = __ldg(tab.mel_owner + m);            // {first slot, slot count} of band m
            float acc = 0.f;
            for (int q = 0; q < o.y; ++q) acc += part[fr * 128 + o.x + q];
            const int t = t0 + 2 * p + fr;
            if (t < T) mel_amp[((size_t)b * T + t) * kMel + m] = acc;
        }
        __syncthreads();
    }
}

// ---- K2a: per-clip max of the clean and of the noisy amplitude mel ------------------------------
__device__ __forceinline__ float4 noise_quad(uint64_t quad_index, uint32_t step, uint64_t seed) {
    const uint4 r = philox4x32_10(quad_index, DCASE_STREAM_NOISE, step, seed);
    const float k = 2.3283064365386963e-10f;  // 2^-32
    const float u0 = ((float)r.x + 0.5f) * k, u1 = ((float)r.y + 0.5f) * k;
    const float u2 = ((float)r.z + 0.5f) * k, u3 = ((float)r.w + 0.5f) * k;
    // (float)r + 0.5f can round to 2^32 -> u = 1 -> log 0 = radius 0: harmless.
    const float ra = sqrtf(-2.f * logf(fminf(u0, 0.99999994f))), rb = sqrtf(-2.f * logf(fminf(u2, 0.99999994f)));
    float sa, ca, sb, cb;
    sincospif(2.f * u1, &sa, &ca);
    sincospif(2.f * u3, &sb, &cb);
    return make_float4(0.25f * fabsf(ra * ca), 0.25f * fabsf(ra * sa), 0.25f * fabsf(rb * cb),
                       0.25f * fabsf(rb * sb));
}

// grid (slices, B): every CTA reduces a slice of the clip and folds it in with atomicMax on the float bits
// (amplitudes are >= 0, so the unsigned order is the float order; clip_max is zeroed by the launcher)
__global__ void __launch_bounds__(256)
clip_max_kernel(const float* __restrict__ mel_amp, int T_in, const float* __restrict__ noise, int want_noisy,
                uint64_t seed, uint32_t step, const DcaseStepScalars* __restrict__ sc,
                float* __restrict__ clip_max, int B) {
    __shared__ float red[2][8];
    if (sc) { seed = sc->seed; step += sc->step; }    // `step` is an offset on top of the device scalars
    const int b = blockIdx.y;
    const int n_quads = T_in * (kMel / 4);
    const float4* src = reinterpret_cast<const float4*>(mel_amp + (size_t)b * T_in * kMel);
    const float4* nz = noise ? reinterpret_cast<const float4*>(noise + (size_t)b * T_in * kMel) : nullptr;
    float mc = 0.f, mn = 0.f;
    for (int q = blockIdx.x * 256 + threadIdx.x; q < n_quads; q += gridDim.x * 256) {
        const float4 x = __ldg(src + q);
        mc = fmaxf(mc, fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w)));
        if (want_noisy) {
            const float4 e = nz ? __ldg(nz + q) : noise_quad((uint64_t)b * n_quads + q, step, seed);
            mn = fmaxf(mn, fmaxf(fmaxf(x.x + e.x, x.y + e.y), fmaxf(x.z + e.z, x.w + e.w)));
        }
    }
    mc = warp_max(mc);
    mn = warp_max(mn);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = mc; red[1][threadIdx.x >> 5] = mn; }
    __syncthreads();
    if (threadIdx.x < 32) {
        mc = threadIdx.x < 8 ? red[0][threadIdx.x] : 0.f;
        mn = threadIdx.x < 8 ? red[1][threadIdx.x] : 0.f;
        mc = warp_max(mc);
        mn = warp_max(mn);
        if (threadIdx.x == 0) {
            atomicMax(reinterpret_cast<unsigned int*>(clip_max) + b, __float_as_uint(fmaxf(mc, 0.f)));
            if (want_noisy) atomicMax(reinterpret_cast<unsigned int*>(clip_max) + B + b, __float_as_uint(fmaxf(mn, 0.f)));
        }
    }
}

__device__ __forceinline__ float amp_to_db(float x) { return 20.f * log10f(fmaxf(x, 1e-5f)); }

// ---- K2b: dB, top_db floor, pad/trunc, z-score -> clean (+ noisy) ---------------------------------
__global__ void __launch_bounds__(256)
finish_kernel(const float* __restrict__ mel_amp, int B, int T_in, int T_out, const float* __restrict__ mean,
              const float* __restrict__ stdv, const float* __restrict__ noise, uint64_t seed, uint32_t step,
              const DcaseStepScalars* __restrict__ sc, const float* __restrict__ clip_max,
              float* __restrict__ clean, float* __restrict__ noisy) {
    if (sc) { seed = sc->seed; step += sc->step; }    // `step` is an offset on top of the device scalars
    const size_t total = (size_t)B * T_out * (kMel / 4);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int q = (int)(i & 15);
        const size_t row = i >> 4;
        const int t = (int)(row % T_out);
        const int b = (int)(row / T_out);
        const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + q);
        const float4 sd = __ldg(reinterpret_cast<const float4*>(stdv) + q);
        float4 lc = make_float4(0.f, 0.f, 0.f, 0.f), ln = lc;  // pad rows are 0.0 dB (DataLoad.py:222-226)
        if (t < T_in) {
            const size_t src_q = ((size_t)b * T_in + t) * 16 + q;
            const float4 x = __ldg(reinterpret_cast<const float4*>(mel_amp) + src_q);
            const float floor_c = amp_to_db(clip_max[b]) - 80.f;
            lc.x = fmaxf(amp_to_db(x.x), floor_c);
            lc.y = fmaxf(amp_to_db(x.y), floor_c);
            lc.z = fmaxf(amp_to_db(x.z), floor_c);
            lc.w = fmaxf(amp_to_db(x.w), floor_c);
            if (noisy) {
                const float4 e = noise ? __ldg(reinterpret_cast<const float4*>(noise) + src_q)
                                       : noise_quad(src_q, step, seed);
                const float floor_n = amp_to_db(clip_max[B + b]) - 80.f;
                ln.x = fmaxf(amp_to_db(x.x + e.x), floor_n);
                ln.y = fmaxf(amp_to_db(x.y + e.y), floor_n);
                ln.z = fmaxf(amp_to_db(x.z + e.z), floor_n);
                ln.w = fmaxf(amp_to_db(x.w + e.w), floor_n);
            }
        }
        float4 o;
        o.x = (lc.x - mu.x) / sd.x; o.y = (lc.y - mu.y) / sd.y; o.z = (lc.z - mu.z) / sd.z; o.w = (lc.w - mu.w) / sd.w;
        reinterpret_cast<float4*>(clean)[i] = o;
        if (noisy) {
            o.x = (ln.x - mu.x) / sd.x; o.y = (ln.y - mu.y) / sd.y; o.z = (ln.z - mu.z) / sd.z; o.w = (ln.w - mu.w) / sd.w;
            reinterpret_cast<float4*>(noisy)[i] = o;
        }
    }
}

// ---- K2c: Scaler.means (utils/Scaler.py:34-87) on the device -----------------------------------------
// sums[0][m] += mean over the T_out frames of L[b][t][m], sums[1][m] += the same of fl32(L * L)  (the reference squares
// in the sample's own dtype, float32, and only the np.mean accumulates in float64).  apply_log: L = dB of the amplitude
// mel with the clip's top_db floor, truncated / zero-padded to T_out frames (ApplyLog -> PadOrTrunc; pad rows are
// 0.0 dB and add nothing); otherwise the rows are taken as they are.  grid (slices, B); a CTA is 16 frame lanes x 16
// quads, so one warp reads two whole 256-byte rows.
__global__ void __launch_bounds__(256)
scaler_accum_kernel(const float* __restrict__ feats, int T_in, int T_out, int apply_log,
                    const float* __restrict__ clip_max, double* __restrict__ sums) {
    __shared__ double red[2][16][kMel];
    const int b = blockIdx.y, q = threadIdx.x & 15, lane = threadIdx.x >> 4;
    const int T = T_in < T_out ? T_in : T_out;
    const int per = (T + gridDim.x - 1) / gridDim.x;
    const int t0 = blockIdx.x * per;
    const int t1 = t0 + per < T ? t0 + per : T;
    const float floor_c = apply_log ? amp_to_db(clip_max[b]) - 80.f : 0.f;
    const float4* src = reinterpret_cast<const float4*>(feats + (size_t)b * T_in * kMel);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
    for (int t = t0 + lane; t < t1; t += 16) {
        float4 x = __ldg(src + (size_t)t * 16 + q);
        if (apply_log) {
            x.x = fmaxf(amp_to_db(x.x), floor_c);
            x.y = fmaxf(amp_to_db(x.y), floor_c);
            x.z = fmaxf(amp_to_db(x.z), floor_c);
            x.w = fmaxf(amp_to_db(x.w), floor_c);
        }
        s0 += (double)x.x; s1 += (double)x.y; s2 += (double)x.z; s3 += (double)x.w;
        r0 += (double)__fmul_rn(x.x, x.x); r1 += (double)__fmul_rn(x.y, x.y);
        r2 += (double)__fmul_rn(x.z, x.z); r3 += (double)__fmul_rn(x.w, x.w);
    }
    red[0][lane][4 * q] = s0; red[0][lane][4 * q + 1] = s1; red[0][lane][4 * q + 2] = s2; red[0][lane][4 * q + 3] = s3;
    red[1][lane][4 * q] = r0; red[1][lane][4 * q + 1] = r1; red[1][lane][4 * q + 2] = r2; red[1][lane][4 * q + 3] = r3;
    __syncthreads();
    if (threadIdx.x < 2 * kMel) {
        const int which = threadIdx.x >> 6, m = threadIdx.x & (kMel - 1);
        double acc = 0.0;
#pragma unroll
        for (int l = 0; l < 16; ++l) acc += red[which][l][m];
        atomicAdd(sums + which * kMel + m, acc / (double)T_out);
    }
}

// mean_ = sums[0] / n, mean_of_square_ = sums[1] / n (Scaler.py:72-73), std_ = sqrt(mean_of_square_ - mean_^2)
// (Scaler.py:31-32, :89-97; no fused multiply-add, as numpy), plus the float32 copies dcase_logmel_finish reads.
__global__ void scaler_finalize_kernel(const double* __restrict__ sums, double n, double* __restrict__ mean_,
                                       double* __restrict__ mean_of_square_, float* __restrict__ mean_f32,
                                       float* __restrict__ std_f32) {
    const int m = threadIdx.x;
    if (m >= kMel) return;
    const double mu = sums[m] / n, sq = sums[kMel + m] / n;
    const double sd = sqrt(__dsub_rn(sq, __dmul_rn(mu, mu)));
    if (mean_) mean_[m] = mu;
    if (mean_of_square_) mean_of_square_[m] = sq;
    if (mean_f32) mean_f32[m] = (float)mu;
    if (std_f32) std_f32[m] = (float)sd;
}

// ---- read_audio's mono mix-down (utils/utils.py:187-189): mean over the channels of interleaved frames ----------
// 16-bit PCM is scaled by 1 / 32768 first (soundfile's float conversion); the mean is taken in fp32 (<= 8 channels of
// 16-bit samples are exact in fp32 up to the final division).
template <typename Tin>
__global__ void __launch_bounds__(256)
mixdown_kernel(const Tin* __restrict__ in, long long n_frames, int n_ch, float* __restrict__ mono) {
    const float scale = sizeof(Tin) == 2 ? 1.f / 32768.f : 1.f;
    const float inv = 1.f / (float)n_ch;
    for (long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x; f < n_frames;
         f += (long long)gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int c = 0; c < n_ch; ++c) acc += (float)in[f * n_ch + c] * scale;
        mono[f] = acc * inv;
    }
}

// ---- host-side constant tables -----------------------------------------------------------------
double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

constexpr size_t kStftSmemBytes = kSpanPad * sizeof(float) + 2 * kFftPaddedSize * sizeof(cf32);

}  // namespace

int dcase_logmel_tables_create(dcase_ctx* ctx) {
    const double sr = 44100.0, f_lo = 0.0, f_hi = 22050.0;
    std::vector<float> win(kNfft);
    for (int n = 0; n < kNfft; ++n) win[n] = (float)(0.54 - 0.46 * cos(2.0 * M_PI * n / (kNfft - 1)));
    std::vector<cf32> tw(kNfft);
    for (int m = 0; m < kNfft; ++m) {
        const double a = -2.0 * M_PI * m / kNfft;
        tw[m] = cf32{(float)cos(a), (float)sin(a)};
    }
    // Slaney filterbank, librosa.filters.mel(sr, n_fft, n_mels=64, fmin, fmax, htk=False, norm=None)
    std::vector<double> mel_f(kMel + 2);
    const double m_lo = hz_to_mel(f_lo), m_hi = hz_to_mel(f_hi);
    for (int i = 0; i < kMel + 2; ++i) mel_f[i] = mel_to_hz(m_lo + (m_hi - m_lo) * i / (kMel + 1));
    ctx->h_mel_dense = (float*)calloc((size_t)kMel * kBins, sizeof(float));
    std::vector<float> packed;
    std::vector<int> start(kMel), len(kMel), off(kMel);
    for (int i = 0; i < kMel; ++i) {
        int first = -1, last = -1;
        for (int k = 0; k < kBins; ++k) {
            const double f = (sr / 2) * k / (kBins - 1);
            const double lower = (f - mel_f[i]) / (mel_f[i + 1] - mel_f[i]);
            const double upper = (mel_f[i + 2] - f) / (mel_f[i + 2] - mel_f[i + 1]);
            const double w = lower < upper ? (lower > 0.0 ? lower : 0.0) : (upper > 0.0 ? upper : 0.0);
            ctx->h_mel_dense[(size_t)i * kBins + k] = (float)w;
            if ((float)w != 0.f) { if (first < 0) first = k; last = k; }
        }
        start[i] = first < 0 ? 0 : first;
        len[i] = first < 0 ? 0 : last - first + 1;
        off[i] = (int)packed.size();
        for (int k = 0; k < len[i]; ++k) packed.push_back(ctx->h_mel_dense[(size_t)i * kBins + start[i] + k]);
    }
    ctx->mel_nnz = (int)packed.size();
    DCASE_CUDA_CHECK(cudaMalloc(&ctx->d_window, kNfft * sizeof(float)));
    DCASE_CUDA_CHECK(cudaMalloc(&ctx->d_twiddle, kNfft * sizeof(cf32)));
    DCASE_CUDA_CHECK(cudaMalloc(&ctx->d_mel_w, packed.size() * sizeof(float)));
    DCASE_CUDA_CHECK(cudaMemcpy(ctx->d_mel_w, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice));
    DCASE_CUDA_CHECK(cudaMemcpy(ctx->d_window, win.data(), kNfft * sizeof(float), cudaMemcpyHostToDevice));
    DCASE_CUDA_CHECK(cudaMemcpy(ctx->d_twiddle, tw.data(), kNfft * sizeof(cf32), cudaMemcpyHostToDevice));
    // balanced work split: band m gets n_m of the 128 slots (proportional to its length, at least one)
    {
        std::vector<int> n_slots(kMel, 1);
        int used = kMel;
        while (used < 128) {            // give the next slot to the band with the largest per-slot load
            int best = 0;
            double best_load = -1.0;
            for (int m = 0; m < kMel; ++m) {
                const double load = (double)len[m] / n_slots[m];
                if (load > best_load) { best_load = load; best = m; }
            }
            ++n_slots[best];
            ++used;
        }
        std::vector<int> work(128 * 4), owner(kMel * 2);
        int slot = 0;
        for (int m = 0; m < kMel; ++m) {
            owner[2 * m] = slot;
            owner[2 * m + 1] = n_slots[m];
            for (int q = 0; q < n_slots[m]; ++q, ++slot) {
                const int beg = (int)((long long)len[m] * q / n_slots[m]), end = (int)((long long)len[m] * (q + 1) / n_slots[m]);
                work[4 * slot] = start[m] + beg;
                work[4 * slot + 1] = off[m] + beg;
                work[4 * slot + 2] = end - beg;
                work[4 * slot + 3] = m;
            }
        }
        DCASE_CUDA_CHECK(cudaMalloc(&ctx->d_mel_work, work.size() * sizeof(int)));
        DCASE_CUDA_CHECK(cudaMalloc(&ctx->d_mel_owner, owner.size() * sizeof(int)));
        DCASE_CUDA_CHECK(cudaMemcpy(ctx->d_mel_work, work.data(), work.size() * sizeof(int), cudaMemcpyHostToDevice));
        DCASE_CUDA_CHECK(cudaMemcpy(ctx->d_mel_owner, owner.data(), owner.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(stft_mel_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)kStftSmemBytes));
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(stft_mel_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)kStftSmemBytes));
    return DCASE_OK;
}

void dcase_logmel_tables_destroy(dcase_ctx* ctx) {
    cudaFree(ctx->d_window); cudaFree(ctx->d_twiddle); cudaFree(ctx->d_mel_w);
    cudaFree(ctx->d_mel_work); cudaFree(ctx->d_mel_owner);
    free(ctx->h_mel_dense);
}

extern "C" {

int dcase_logmel_num_frames(int n_samples) { return 1 + n_samples / kHop; }

int dcase_mel_filterbank(dcase_ctx* ctx, float* out_host) {
    DCASE_REQUIRE(ctx && out_host, "null argument");
    memcpy(out_host, ctx->h_mel_dense, (size_t)kMel * kBins * sizeof(float));
    return DCASE_OK;
}

static int logmel_fwd_impl(dcase_ctx* ctx, const void* wave, int is_pcm16, int B, int L, float* mel_amp,
                           cudaStream_t stream) {
    DCASE_REQUIRE(ctx && wave && mel_amp, "null argument");
    DCASE_REQUIRE(B >= 0 && L > kNfft / 2, "need L > 1024 samples (single reflection)");
    if (B == 0) return DCASE_OK;
    const int T = 1 + L / kHop;
    DCASE_PROF("stft_mel", stream);
    MelTables tab{ctx->d_window, ctx->d_twiddle, ctx->d_mel_w, (const int4*)ctx->d_mel_work, (const int2*)ctx->d_mel_owner};
    dim3 grid((T + kFramesPerCta - 1) / kFramesPerCta, B);
    if (is_pcm16)
        stft_mel_kernel<int16_t><<<grid, 256, kStftSmemBytes, stream>>>((const int16_t*)wave, L, T, tab, mel_amp);
    else
        stft_mel_kernel<float><<<grid, 256, kStftSmemBytes, stream>>>((const float*)wave, L, T, tab, mel_amp);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int dcase_logmel_fwd(dcase_ctx* ctx, const float* wave, int B, int L, float* mel_amp, void* stream) {
    return logmel_fwd_impl(ctx, wave, 0, B, L, mel_amp, (cudaStream_t)stream);
}

int dcase_logmel_fwd_pcm16(dcase_ctx* ctx, const int16_t* wave, int B, int L, float* mel_amp, void* stream) {
    return logmel_fwd_impl(ctx, wave, 1, B, L, mel_amp, (cudaStream_t)stream);
}

int dcase_logmel_finish(dcase_ctx* ctx, const float* mel_amp, int B, int T_in, int T_out, const float* mean,
                        const float* stdv, const float* noise, uint64_t seed, uint32_t step, const void* scalars,
                        float* clip_max_ws, float* clean, float* noisy, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DCASE_REQUIRE(ctx && mel_amp && mean && stdv && clip_max_ws && clean, "null argument");
    DCASE_REQUIRE(B >= 0 && T_in > 0 && T_out > 0, "bad shape");
    if (B == 0) return DCASE_OK;
    const DcaseStepScalars* sc = (const DcaseStepScalars*)scalars;
    DCASE_PROF("logmel_finish", stream);
    DCASE_CUDA_CHECK(cudaMemsetAsync(clip_max_ws, 0, 2 * (size_t)B * sizeof(float), stream));
    const int n_quads = T_in * (kMel / 4);
    int slices = (n_quads + 1023) / 1024;                       // >= 4 quads per thread
    const int want = (ctx->num_sms * 8 + B - 1) / B;
    if (slices > want) slices = want;
    if (slices < 1) slices = 1;
    clip_max_kernel<<<dim3(slices, B), 256, 0, stream>>>(mel_amp, T_in, noise, noisy != nullptr, seed, step, sc,
                                                         clip_max_ws, B);
    DCASE_LAUNCH_CHECK();
    const size_t total = (size_t)B * T_out * 16;
    int blocks = (int)((total + 255) / 256);
    if (blocks > ctx->num_sms * 8) blocks = ctx->num_sms * 8;
    finish_kernel<<<blocks, 256, 0, stream>>>(mel_amp, B, T_in, T_out, mean, stdv, noise, seed, step, sc, clip_max_ws,
                                             clean, noisy);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int dcase_scaler_accumulate(dcase_ctx* ctx, const float* feats, int B, int T_in, int T_out, int apply_log,
                            float* clip_max_ws, double* sums, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DCASE_REQUIRE(ctx && feats && sums, "null argument");
    DCASE_REQUIRE(B >= 0 && T_in > 0 && T_out > 0, "bad shape");
    DCASE_REQUIRE(!apply_log || clip_max_ws, "apply_log needs the [B] clip_max scratch");
    DCASE_REQUIRE(apply_log || T_in == T_out, "finished features are reduced as they are (T_in == T_out)");
    if (B == 0) return DCASE_OK;
    DCASE_PROF("scaler_accumulate", stream);
    const int want = (ctx->num_sms * 8 + B - 1) / B;
    if (apply_log) {
        DCASE_CUDA_CHECK(cudaMemsetAsync(clip_max_ws, 0, (size_t)B * sizeof(float), stream));
        const int n_quads = T_in * (kMel / 4);
        int slices = (n_quads + 1023) / 1024;
        if (slices > want) slices = want;
        if (slices < 1) slices = 1;
        clip_max_kernel<<<dim3(slices, B), 256, 0, stream>>>(feats, T_in, nullptr, 0, 0, 0, nullptr, clip_max_ws, B);
        DCASE_LAUNCH_CHECK();
    }
    const int T = T_in < T_out ? T_in : T_out;
    int slices = (T + 63) / 64;                                  // >= 4 rows per thread
    if (slices > want) slices = want;
    if (slices < 1) slices = 1;
    scaler_accum_kernel<<<dim3(slices, B), 256, 0, stream>>>(feats, T_in, T_out, apply_log, clip_max_ws, sums);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int dcase_scaler_finalize(dcase_ctx* ctx, const double* sums, long long n_samples, double* mean,
                          double* mean_of_square, float* mean_f32, float* std_f32, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DCASE_REQUIRE(ctx && sums, "null argument");
    DCASE_REQUIRE(n_samples > 0, "Scaler.means over an empty dataset");
    DCASE_PROF("scaler_finalize", stream);
    scaler_finalize_kernel<<<1, kMel, 0, stream>>>(sums, (double)n_samples, mean, mean_of_square, mean_f32, std_f32);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int dcase_audio_mixdown(dcase_ctx* ctx, const void* interleaved, int is_pcm16, long long n_frames, int n_channels,
                        float* mono, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DCASE_REQUIRE(ctx && interleaved && mono, "null argument");
    DCASE_REQUIRE(n_frames >= 0 && n_channels >= 1 && n_channels <= 64, "bad shape");
    if (n_frames == 0) return DCASE_OK;
    DCASE_PROF("audio_mixdown", stream);
    long long blocks = (n_frames + 255) / 256;
    if (blocks > ctx->num_sms * 16) blocks = ctx->num_sms * 16;
    if (is_pcm16)
        mixdown_kernel<int16_t><<<(int)blocks, 256, 0, stream>>>((const int16_t*)interleaved, n_frames, n_channels, mono);
    else
        mixdown_kernel<float><<<(int)blocks, 256, 0, stream>>>((const float*)interleaved, n_frames, n_channels, mono);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

}  // extern "C"
