// K1/K2: waveform -> Hamming STFT magnitude -> Slaney mel (amplitude) -> noise / dB / top_db / z-score.
//
// Replaces (reference file:line):
//   baseline/DatasetDcase2019Task4.py:197-231   calculate_mel_spec            -> dcase_logmel_fwd
//   baseline/DataLoad.py:274-287,192-207,210-259,302-321 + utils/Scaler.py:99-105
//   (order: utils/utils.py:397-412 get_transforms)                            -> dcase_logmel_finish
//
// K1 layout (round 2): persistent CTAs of 8 warps; a tile = 8 consecutive frames of one clip, ONE WARP PER FRAME.
//   * the tile's contiguous waveform span (7 * 511 + 2048 samples, reflect-padded at the clip ends) is staged once per
//     tile with 16-byte global loads and stored DE-INTERLEAVED by sample parity (E[m] = x[2m], O[m] = x[2m+1]), so that
//     the packing z[n] = x[2n] + i x[2n+1] of a real frame into a 1024-point complex FFT reads two stride-1 arrays
//     (conflict-free) whatever the parity of the frame's offset (hop 511 is odd);
//   * the warp runs the whole real FFT on its own (csrc/fft1024.cuh): two 32-point DFTs in registers with one
//     32 x 32 transpose through the warp's private 8.4 KB of shared memory, the Hermitian split with the partner values
//     fetched by warp shuffles, packed f32x2 additions -- no block barrier inside a frame;
//   * magnitudes go to the warp's scratch, the sparse Slaney projection (1,983 non-zeros, weights in shared memory,
//     staged once per CTA together with the interleaved window by two cp.async.bulk copies) runs as 4 balanced work
//     items per lane, band owners add the partials and write the frame's 64 mel amplitudes as two 128-byte rows.
// 108 KB of shared memory per CTA: two CTAs (16 warps) per SM; while one stages its next span the other computes.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/dcase_b200.h"
#include "common.cuh"
#include "ctx.h"
#include "fft1024.cuh"

namespace {

constexpr int kNfft = 2048;
constexpr int kHop = 511;
constexpr int kBins = 1025;
constexpr int kMel = 64;
constexpr int kFramesPerTile = 8;
constexpr int kStftWarps = 8;
constexpr int kSpan = (kFramesPerTile - 1) * kHop + kNfft;  // 5625 samples
constexpr int kHalfSpan = 2816;                              // ceil(5625 / 2) = 2813, padded
constexpr int kMelWPad = 1984;                               // 1,983 packed weights, padded to 16 bytes
constexpr int kMelItems = 128;                               // work items of the mel projection (4 per lane)

struct MelTables {
    const float2* window2;   // [1024] {w[2n], w[2n+1]} symmetric Hamming
    const float* mel_w;      // [kMelWPad] packed non-zero Slaney weights
    const int4* mel_work;    // [128] work items {bin start, weight offset, count, band}
    const int2* mel_owner;   // [64]  {first slot, slot count} per band
};

struct StftSmem {
    float even[kHalfSpan];                 // x[s0 + 2m]
    float odd[kHalfSpan];                  // x[s0 + 2m + 1]
    float2 window2[kNfft / 2];
    float mel_w[kMelWPad];
    cpx xchg[kStftWarps][kXchgSize];       // per-warp transpose buffer, re-used for the frame's 1025 magnitudes
    float part[kStftWarps][kMelItems];
    unsigned long long bar;
};

__device__ __forceinline__ int reflect_index(int s, int L) {
    if (s < 0) s = -s;
    if (s >= L) s = 2 * (L - 1) - s;
    return s;
}

template <typename WaveT>
__device__ __forceinline__ float load_sample(const WaveT* p, long long i);
template <>
__device__ __forceinline__ float load_sample<float>(const float* p, long long i) { return __ldg(p + i); }
template <>
__device__ __forceinline__ float load_sample<int16_t>(const int16_t* p, long long i) {
    return (float)__ldg(p + i) * (1.0f / 32768.0f);  // soundfile's int16 -> float scaling
}

// one MUFU: sqrt.approx (relative error ~1e-7), exact 0 for x = 0 and exact under scaling by 4
__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Stage the span [s0, s0 + kSpan) of one clip, de-interleaved.  Interior tiles whose global address is 16-byte
// aligned-down-able read whole 16-byte vectors (4 floats / 8 PCM samples); edge tiles (reflect padding) go sample by
// sample.
template <typename WaveT>
__device__ __forceinline__ void stage_span(const WaveT* __restrict__ wave, long long clip_off, long long total, int L,
                                           int s0, bool base_aligned, StftSmem& sm, int tid) {
    constexpr int V = 16 / (int)sizeof(WaveT);
    const bool interior = s0 >= 0 && s0 + kSpan <= L;
    if (interior && base_aligned) {
        const long long g0 = clip_off + s0;
        const int a = (int)(g0 & (V - 1));
        const long long gv = g0 - a;                               // 16-byte aligned element index
        const int n_vec = (kSpan + a + V - 1) / V;
        for (int c = tid; c < n_vec; c += 32 * kStftWarps) {
            const long long g = gv + (long long)V * c;
            float x[V];
            if (g + V <= total) {
                if (sizeof(WaveT) == 4) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(wave + g));
                    x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
                } else {
                    const int4 q = __ldg(reinterpret_cast<const int4*>(wave + g));
                    const int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {              // little endian: the low half is the earlier sample
                        x[(2 * e) % V] = (float)(short)(w[e] & 0xFFFF) * (1.0f / 32768.0f);
                        x[(2 * e + 1) % V] = (float)(short)(w[e] >> 16) * (1.0f / 32768.0f);
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < V; ++e) x[e] = g + e < total ? load_sample<WaveT>(wave, g + e) : 0.f;
            }
            const int i0 = V * c - a;
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const int i = i0 + e;
                if (i >= 0 && i < kSpan) (i & 1 ? sm.odd : sm.even)[i >> 1] = x[e];
            }
        }
    } else {
        for (int i = tid; i < kSpan; i += 32 * kStftWarps) {
            const int s = reflect_index(s0 + i, L);
            const float x = (s >= 0 && s < L) ? load_sample<WaveT>(wave, clip_off + s) : 0.f;
            (i & 1 ? sm.odd : sm.even)[i >> 1] = x;
        }
    }
}

template <typename WaveT>
__global__ void __launch_bounds__(32 * kStftWarps, 2)
stft_mel_kernel(const WaveT* __restrict__ wave, int B, int L, int T, MelTables tab, float* __restrict__ mel_amp) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];   // every member of StftSmem is a multiple of 16 bytes
    StftSmem& sm = *reinterpret_cast<StftSmem*>(smem_dyn);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // constant tables: two bulk copies per CTA (window 8 KB, mel weights 7.8 KB), complete on an mbarrier
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&sm.bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&sm.bar)),
                     "r"((uint32_t)(sizeof(sm.window2) + sizeof(sm.mel_w))) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_addr(sm.window2)), "l"(tab.window2), "r"((uint32_t)sizeof(sm.window2)), "r"(smem_addr(&sm.bar))
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_addr(sm.mel_w)), "l"(tab.mel_w), "r"((uint32_t)sizeof(sm.mel_w)), "r"(smem_addr(&sm.bar))
                     : "memory");
    }

    // per-lane constants: W_1024^(lane 2^j) for the pass-1 twiddles, W_2048^lane for the Hermitian split
    Pass1Twiddles tw1;
    cpx wl;
    {
        float s, c;
        sincospif(-(float)lane / 512.f, &s, &c);   tw1.w1 = cmake(c, s);
        sincospif(-(float)lane / 256.f, &s, &c);   tw1.w2 = cmake(c, s);
        sincospif(-(float)lane / 128.f, &s, &c);   tw1.w4 = cmake(c, s);
        sincospif(-(float)lane / 64.f, &s, &c);    tw1.w8 = cmake(c, s);
        sincospif(-(float)lane / 32.f, &s, &c);    tw1.w16 = cmake(c, s);
        sincospif(-(float)lane / 1024.f, &s, &c);  wl = cmake(c, s);
    }
    const bool base_aligned = (reinterpret_cast<uintptr_t>(wave) & 15) == 0;
    const long long total = (long long)B * L;
    const int tiles_per_clip = (T + kFramesPerTile - 1) / kFramesPerTile;
    const int n_tiles = tiles_per_clip * B;
    cpx* const xb = sm.xchg[warp];
    float* const mag = reinterpret_cast<float*>(xb);
    bool tables_ready = false;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_clip;
        const int t0 = (tile - b * tiles_per_clip) * kFramesPerTile;
        __syncthreads();                                          // every warp is done with the previous span
        stage_span<WaveT>(wave, (long long)b * L, total, L, t0 * kHop - kNfft / 2, base_aligned, sm, tid);
        __syncthreads();
        if (!tables_ready) {
            uint32_t done = 0;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(smem_addr(&sm.bar)) : "memory");
            tables_ready = true;
        }
        const int t = t0 + warp;
        if (t >= T) continue;                                     // warp-uniform: the tile's tail frames do not exist

        // ---- load + window: z[n1] = (x[o + 2n], x[o + 2n + 1]) * (w[2n], w[2n + 1]), n = 32 n1 + lane ----
        const int o = warp * kHop;
        const float* pre = (o & 1) ? sm.odd + (o >> 1) : sm.even + (o >> 1);
        const float* pim = (o & 1) ? sm.even + ((o + 1) >> 1) : sm.odd + (o >> 1);
        cpx v[32];
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            const float2 w = sm.window2[32 * n1 + lane];
            v[n1] = cmul_elem(cmake(pre[32 * n1 + lane], pim[32 * n1 + lane]), cmake(w.x, w.y));
        }
        stft_pass1(v, tw1, lane, xb);
        __syncwarp();
        stft_pass2(v, lane, xb);                                  // v[k2] = Z[lane + 32 k2]
        __syncwarp();                                             // the transpose buffer becomes the magnitude buffer

        // ---- Hermitian split: partner Z[1024 - k] of k = lane + 32 p sits in lane (32 - lane) % 32, register 31 - p
        //      (lane 0 pairs with itself: Z[1024 - 32 p] is its own register 32 - p, Z[1024] = Z[0]) ----
        {
            constexpr float c64[16] = DCASE_W64_COS;
            constexpr float s64[16] = DCASE_W64_SIN;
            const int src = (32 - lane) & 31;
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                const float mx = __shfl_sync(0xffffffffu, cre(v[31 - p]), src);
                const float my = __shfl_sync(0xffffffffu, cim(v[31 - p]), src);
                const cpx own = v[(32 - p) & 31];
                const cpx zm = lane == 0 ? own : cmake(mx, my);
                float lo, hi;
                stft_post_pair(v[p], zm, wl, c64[p], s64[p], lo, hi);
                mag[lane + 32 * p] = fast_sqrt(lo);               // 2 |X[k]|: the staged mel weights carry the 1/2
                mag[1024 - lane - 32 * p] = fast_sqrt(hi);
            }
            if (lane == 0) {                                      // |X[512]| = |Z[512]|
                const float x = cre(v[16]), y = cim(v[16]);
                mag[512] = 2.f * fast_sqrt(x * x + y * y);
            }
        }
        __syncwarp();

        // ---- sparse Slaney projection: 4 work items per lane (each a contiguous run of weights inside one band) ----
        float* part = sm.part[warp];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int4 it = __ldg(tab.mel_work + lane + 32 * j);  // {bin start, weight offset, count, band}
            const float* mg = mag + it.x;
            const float* wt = sm.mel_w + it.y;
            float a0 = 0.f, a1 = 0.f;
            int i = 0;
            for (; i + 1 < it.z; i += 2) {
                a0 = fmaf(wt[i], mg[i], a0);
                a1 = fmaf(wt[i + 1], mg[i + 1], a1);
            }
            if (i < it.z) a0 = fmaf(wt[i], mg[i], a0);
            part[lane + 32 * j] = a0 + a1;
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int m = lane + 32 * j;
            const int2 ow = __ldg(tab.mel_owner + m);             // {first slot, slot count} of band m
            float acc = 0.f;
            for (int q = 0; q < ow.y; ++q) acc += part[ow.x + q];
            mel_amp[((size_t)b * T + t) * kMel + m] = acc;
        }
        __syncwarp();                                             // part / mag are re-used by the next tile
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

constexpr size_t kStftSmemBytes = sizeof(StftSmem);

}  // namespace

int dcase_logmel_tables_create(dcase_ctx* ctx) {
    const double sr = 44100.0, f_lo = 0.0, f_hi = 22050.0;
    std::vector<float> win(kNfft);            // np.hamming(2048); stored as float2 {w[2n], w[2n + 1]}
    for (int n = 0; n < kNfft; ++n) win[n] = (float)(0.54 - 0.46 * cos(2.0 * M_PI * n / (kNfft - 1)));
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
    if ((int)packed.size() > kMelWPad) { dcase_set_error("mel weight table larger than its shared-memory slot"); return DCASE_ERR_STATE; }
    packed.resize(kMelWPad, 0.f);
    for (float& w : packed) w *= 0.5f;        // the kernel's magnitudes are 2 |X[k]| (exact power-of-two rescale)
    DCASE_CUDA_CHECK(cudaMalloc(&ctx->d_window2, kNfft * sizeof(float)));
    DCASE_CUDA_CHECK(cudaMalloc(&ctx->d_mel_w, packed.size() * sizeof(float)));
    DCASE_CUDA_CHECK(cudaMemcpy(ctx->d_mel_w, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice));
    DCASE_CUDA_CHECK(cudaMemcpy(ctx->d_window2, win.data(), kNfft * sizeof(float), cudaMemcpyHostToDevice));
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
    // two 108 KB CTAs per SM need the whole carve-out
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(stft_mel_kernel<float>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(stft_mel_kernel<int16_t>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    return DCASE_OK;
}

void dcase_logmel_tables_destroy(dcase_ctx* ctx) {
    cudaFree(ctx->d_window2); cudaFree(ctx->d_mel_w);
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
    MelTables tab{ctx->d_window2, ctx->d_mel_w, (const int4*)ctx->d_mel_work, (const int2*)ctx->d_mel_owner};
    const long long n_tiles = (long long)((T + kFramesPerTile - 1) / kFramesPerTile) * B;
    DCASE_REQUIRE(n_tiles < (1ll << 31), "too many frames for one launch");
    // persistent CTAs: two per SM (108 KB of shared memory each), each looping over 8-frame tiles
    const int grid = (int)(n_tiles < 2ll * ctx->num_sms ? n_tiles : 2ll * ctx->num_sms);
    if (is_pcm16)
        stft_mel_kernel<int16_t><<<grid, 32 * kStftWarps, kStftSmemBytes, stream>>>((const int16_t*)wave, B, L, T, tab, mel_amp);
    else
        stft_mel_kernel<float><<<grid, 32 * kStftWarps, kStftSmemBytes, stream>>>((const float*)wave, B, L, T, tab, mel_amp);
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
