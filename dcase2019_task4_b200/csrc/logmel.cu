// K1/K2: waveform -> Hamming STFT magnitude -> Slaney mel (amplitude) -> noise / dB / top_db / z-score.
//
// Replaces (reference file:line):
//   baseline/DatasetDcase2019Task4.py:197-231   calculate_mel_spec            -> dcase_logmel_fwd
//   baseline/DataLoad.py:274-287,192-207,210-259,302-321 + utils/Scaler.py:99-105
//   (order: utils/utils.py:397-412 get_transforms)                            -> dcase_logmel_finish
//
// K1 layout (round 2): persistent CTAs of 8 warps; a tile = 8 consecutive frames of one clip, ONE WARP PER FRAME.
//   * the tile's contiguous waveform span (7 * 511 + 2048 samples) arrives by ONE cp.async.bulk (1-D TMA) of its 16-byte
//     aligned superset, issued for tile i + 1 as soon as every warp has pulled its frame of tile i into registers, so
//     the copy flies under the FFT / mel phase of tile i (mbarrier completion).  Tiles that touch the clip ends
//     (reflect padding), 16-bit PCM input and unaligned buffers are staged by ordinary loads instead;
//   * the warp runs the whole real FFT on its own (csrc/fft1024.cuh): two 32-point DFTs in registers with one
//     32 x 32 transpose through the warp's private 8.4 KB of shared memory, the Hermitian split with the partner values
//     fetched by warp shuffles, packed f32x2 additions -- no block barrier inside the FFT;
//   * magnitudes go to the warp's scratch; the sparse Slaney projection (1,983 non-zeros) runs as 4 work items per lane
//     (each a run of <= 22 weights inside one band) with the weights stored TRANSPOSED and zero padded in shared
//     memory ([round][i][lane]: conflict-free, fixed trip count, no predicates); band owners add the partials and
//     write the frame's 64 mel amplitudes as two 128-byte rows.  Window + mel tables: cp.async.bulk once per CTA.
// 112 KB of shared memory per CTA: two CTAs (16 warps) per SM.
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
constexpr int kStftThreads = 32 * kStftWarps;
constexpr int kSpan = (kFramesPerTile - 1) * kHop + kNfft;  // 5625 samples
constexpr int kSpanBuf = 5632;                               // + up to 3 samples of alignment slack, 16-byte multiple
// 16-bit PCM input: the raw span (+ up to 7 samples of slack, <= 5640 samples) lands in the UPPER half of the span buffer by
// the same single bulk copy and is expanded to float32 in place (every thread first holds its samples in registers)
constexpr int kPcmOff = kSpanBuf * 4 - 5640 * 2;             // byte offset, a multiple of 16
constexpr int kPcmPerThread = (kSpan + kStftThreads - 1) / kStftThreads;
static_assert(kPcmOff % 16 == 0, "bulk copies need 16-byte aligned destinations");
constexpr int kMelItems = 128;                               // work items of the mel projection (4 rounds x 32 lanes)
constexpr int kMelRun = 22;                                  // longest run of one item

struct MelTables {
    const float2* window2;   // [1024] {w[2n], w[2n+1]} symmetric Hamming
    const float* mel_wt;     // [4][kMelRun][32] weights of item (lane + 32 j), transposed, zero padded, pre-halved
    const int* mel_start;    // [128] first FFT bin of each item
    const int2* mel_owner;   // [64]  {first item, item count} per band
};

struct StftSmem {
    float span[kSpanBuf];
    float2 window2[kNfft / 2];
    float mel_wt[4 * kMelRun * 32];
    int mel_start[kMelItems];
    int2 mel_owner[kMel];
    cpx xchg[kStftWarps][kXchgSize];       // per-warp transpose buffer, re-used for the frame's 1025 magnitudes
    float part[kStftWarps][kMelItems];
    unsigned long long bar_tables, bar_span;
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
__device__ __forceinline__ void mbar_init1(unsigned long long* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(unsigned long long* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
}

// Where a tile's span sits: sm.span[i] = x[s0 - a + i] (reflect padded), a = alignment slack of the bulk copy.
struct TileStage {
    int b, t0, a;
    bool async;
    int a16;      // 16-bit PCM: slack of the raw span in the upper half of the buffer (a = 0 after the expansion)
};

// Stage the span of `tile` into sm.span.  Returns how it was staged; the async form is issued by thread 0 alone and
// completes on sm.bar_span, the synchronous form is written by all threads (the caller barriers before reading).
template <typename WaveT>
__device__ __forceinline__ TileStage stage_span(const WaveT* __restrict__ wave, int tile, int tiles_per_clip, int B, int L,
                                                bool base_aligned, StftSmem& sm, int tid) {
    TileStage st;
    st.b = tile / tiles_per_clip;
    st.t0 = (tile - st.b * tiles_per_clip) * kFramesPerTile;
    const int s0 = st.t0 * kHop - kNfft / 2;
    const long long clip_off = (long long)st.b * L;
    st.a = 0;
    st.a16 = 0;
    st.async = false;
    if (sizeof(WaveT) == 2 && base_aligned && s0 >= 0 && s0 + kSpan <= L) {
        const long long g0 = clip_off + s0;
        const int a = (int)(g0 & 7);
        const uint32_t n = (uint32_t)((kSpan + a + 7) & ~7);
        if (g0 - a + n <= (long long)B * L) {
            st.a16 = a;
            st.async = true;
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect(&sm.bar_span, 2u * n);
                bulk_load(reinterpret_cast<unsigned char*>(sm.span) + kPcmOff, wave + (g0 - a), 2u * n, &sm.bar_span);
            }
            return st;
        }
    }
    if (sizeof(WaveT) == 4 && base_aligned && s0 >= 0 && s0 + kSpan <= L) {
        const long long g0 = clip_off + s0;
        const int a = (int)(g0 & 3);
        const uint32_t n = (uint32_t)((kSpan + a + 3) & ~3);
        if (g0 - a + n <= (long long)B * L) {
            st.a = a;
            st.async = true;
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads / writes of the buffer
                mbar_expect(&sm.bar_span, 4u * n);
                bulk_load(sm.span, wave + (g0 - a), 4u * n, &sm.bar_span);
            }
            return st;
        }
    }
    for (int i = tid; i < kSpan; i += kStftThreads) {
        const int s = reflect_index(s0 + i, L);
        sm.span[i] = (s >= 0 && s < L) ? load_sample<WaveT>(wave, clip_off + s) : 0.f;
    }
    return st;
}

template <typename WaveT>
__global__ void __launch_bounds__(kStftThreads, 2)
stft_mel_kernel(const WaveT* __restrict__ wave, int B, int L, int T, MelTables tab, float* __restrict__ mel_amp) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];   // every member of StftSmem is a multiple of 16 bytes
    StftSmem& sm = *reinterpret_cast<StftSmem*>(smem_dyn);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // constant tables: four bulk copies per CTA (window 8 KB, mel weights 11 KB, item starts, band owners)
    if (tid == 0) {
        mbar_init1(&sm.bar_tables);
        mbar_init1(&sm.bar_span);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect(&sm.bar_tables, (uint32_t)(sizeof(sm.window2) + sizeof(sm.mel_wt) + sizeof(sm.mel_start) + sizeof(sm.mel_owner)));
        bulk_load(sm.window2, tab.window2, (uint32_t)sizeof(sm.window2), &sm.bar_tables);
        bulk_load(sm.mel_wt, tab.mel_wt, (uint32_t)sizeof(sm.mel_wt), &sm.bar_tables);
        bulk_load(sm.mel_start, tab.mel_start, (uint32_t)sizeof(sm.mel_start), &sm.bar_tables);
        bulk_load(sm.mel_owner, tab.mel_owner, (uint32_t)sizeof(sm.mel_owner), &sm.bar_tables);
    }
    __syncthreads();                                              // the mbarriers exist before anybody polls them

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
    const int tiles_per_clip = (T + kFramesPerTile - 1) / kFramesPerTile;
    const int n_tiles = tiles_per_clip * B;
    cpx* const xb = sm.xchg[warp];
    float* const mag = reinterpret_cast<float*>(xb);
    float* const part = sm.part[warp];
    uint32_t span_parity = 0;

    int tile = blockIdx.x;
    TileStage cur{0, 0, 0, false, 0};
    if (tile < n_tiles) cur = stage_span<WaveT>(wave, tile, tiles_per_clip, B, L, base_aligned, sm, tid);
    mbar_wait_parity(&sm.bar_tables, 0);

    for (; tile < n_tiles; tile += gridDim.x) {
        if (cur.async) { mbar_wait_parity(&sm.bar_span, span_parity); span_parity ^= 1; }
        if (sizeof(WaveT) == 2 && cur.async) {                    // raw 16-bit span -> float32 in place (soundfile's 1/32768)
            const int16_t* raw = reinterpret_cast<const int16_t*>(reinterpret_cast<const unsigned char*>(sm.span) + kPcmOff) + cur.a16;
            int16_t smp[kPcmPerThread];
#pragma unroll
            for (int k = 0; k < kPcmPerThread; ++k) {
                const int i = tid + k * kStftThreads;
                smp[k] = i < kSpan ? raw[i] : (int16_t)0;
            }
            __syncthreads();                                      // everybody holds its samples: the buffer may be rewritten
#pragma unroll
            for (int k = 0; k < kPcmPerThread; ++k) {
                const int i = tid + k * kStftThreads;
                if (i < kSpan) sm.span[i] = (float)smp[k] * (1.0f / 32768.0f);
            }
            __syncthreads();
        } else if (!cur.async) __syncthreads();                   // the span was written with ordinary stores
        const int t = cur.t0 + warp;
        const bool active = t < T;                                // warp-uniform: the tile's tail frames do not exist
        const int b = cur.b;

        // ---- load + window: z[n1] = (x[o + 2n], x[o + 2n + 1]) * (w[2n], w[2n + 1]), n = 32 n1 + lane ----
        cpx v[32];
        if (active) {
            const float* fr = sm.span + cur.a + warp * kHop + 2 * lane;
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) {
                const float2 w = sm.window2[32 * n1 + lane];
                v[n1] = cmul_elem(cmake(fr[64 * n1], fr[64 * n1 + 1]), cmake(w.x, w.y));
            }
        }
        __syncthreads();                                          // every warp holds its frame: the span buffer is free
        const int next = tile + gridDim.x;
        if (next < n_tiles) cur = stage_span<WaveT>(wave, next, tiles_per_clip, B, L, base_aligned, sm, tid);
        if (!active) continue;

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

        // ---- sparse Slaney projection: item (lane + 32 j) is a run of <= 22 weights inside one band; the zero padding
        //      of the transposed weight table makes the trip count fixed (magnitudes past the run are finite) ----
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float* mg = mag + sm.mel_start[lane + 32 * j];
            const float* wt = sm.mel_wt + j * (kMelRun * 32) + lane;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int i = 0; i < kMelRun; i += 2) {
                a0 = fmaf(wt[32 * i], mg[i], a0);
                a1 = fmaf(wt[32 * i + 32], mg[i + 1], a1);
            }
            part[lane + 32 * j] = a0 + a1;
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int m = lane + 32 * j;
            const int2 ow = sm.mel_owner[m];                      // {first item, item count} of band m
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

// ---- read_audio's resampler (utils/utils.py:190-192 -> librosa.resample -> resampy kaiser_best) -----------------------
// One thread per output sample: band-limited sinc interpolation with the half window `win` (and its forward differences
// `delta`) sampled 512 times per zero crossing; the table index advances by int(scale * 512) per input sample.  Times in
// fp64 (resampy's own arithmetic), weights and the sum in fp32.  Samples past resampy's int(n_in * ratio) are librosa's
// fix_length zero padding.
constexpr int kRsZeros = 64, kRsTable = 512, kRsWin = kRsZeros * kRsTable + 1;
// resampy's clock is the running sum  time_register += 1 / ratio  in float64.  Its rounding matters: whenever the exact time
// is an integer (every 147th output sample at 48 -> 44.1 kHz) the algorithm's own index truncation makes the result jump by
// ~1e-3 depending on which side of the integer the rounded sum falls.  The sum is reproduced EXACTLY without a serial
// loop: while the register stays inside one binade every addition rounds the same way, so it advances by a constant
// representable increment -- a table of <= kRsSegs linear segments {first sample, register there, increment}.
constexpr int kRsSegs = 160;
struct ResampleClock {
    long long t0[kRsSegs];
    double s0[kRsSegs];
    double inc[kRsSegs];
    int n;
};
int build_resample_clock(long long n_steps, double time_increment, ResampleClock* c) {
    c->n = 0;
    double s = 0.0;
    long long t = 0;
    while (t < n_steps) {
        if (c->n >= kRsSegs) return -1;
        const double nxt = s + time_increment;               // the one rounding resampy performs per sample
        int es = 0, en = 0;
        frexp(s, &es);
        frexp(nxt, &en);
        long long m = 1;
        const double inc = nxt - s;                          // exact: both are multiples of the smaller ulp
        if (s != 0.0 && es == en) {                          // same binade before and after: constant increment inside it
            const double top = ldexp(1.0, es);               // frexp: s in [2^(es-1), 2^es)
            m = (long long)((top - s) / inc);
            while (m > 1 && s + (double)m * inc >= top) --m;
            if (m < 1) m = 1;
        }
        c->t0[c->n] = t; c->s0[c->n] = s; c->inc[c->n] = inc; ++c->n;
        t += m;
        s += (double)m * inc;                                // exact (a multiple of the binade's ulp below its top)
    }
    return 0;
}

__global__ void __launch_bounds__(256)
resample_kernel(const float* __restrict__ x, long long n_in, float* __restrict__ y, long long n_out, long long n_valid,
                const __grid_constant__ ResampleClock clk, double scale, int index_step, float gain,
                const float* __restrict__ win, const float* __restrict__ delta) {
    const long long t = blockIdx.x * 256ll + threadIdx.x;
    if (t >= n_out) return;
    if (t >= n_valid) { y[t] = 0.f; return; }
    int sg = 0;
    for (int i = 1; i < clk.n; ++i) sg = clk.t0[i] <= t ? i : sg;
    const double time_register = clk.s0[sg] + (double)(t - clk.t0[sg]) * clk.inc[sg];
    const long long n = (long long)time_register;
    double frac = scale * (time_register - (double)n);
    float acc = 0.f;
    {
        const double index_frac = frac * kRsTable;
        const int offset = (int)index_frac;
        const float eta = (float)(index_frac - offset);
        long long i_max = (kRsWin - offset) / index_step;
        if (i_max > n + 1) i_max = n + 1;
        for (long long i = 0; i < i_max; ++i) {
            const int k = offset + (int)i * index_step;
            acc = fmaf(fmaf(eta, __ldg(delta + k), __ldg(win + k)), __ldg(x + n - i), acc);
        }
    }
    {
        frac = scale - frac;
        const double index_frac = frac * kRsTable;
        const int offset = (int)index_frac;
        const float eta = (float)(index_frac - offset);
        long long k_max = (kRsWin - offset) / index_step;
        if (k_max > n_in - n - 1) k_max = n_in - n - 1;
        for (long long q = 0; q < k_max; ++q) {
            const int k = offset + (int)q * index_step;
            acc = fmaf(fmaf(eta, __ldg(delta + k), __ldg(win + k)), __ldg(x + n + q + 1), acc);
        }
    }
    y[t] = gain * acc;
}

// modified Bessel function I0 by its power series (numpy.kaiser's i0), double precision
double bessel_i0(double x) {
    double sum = 1.0, term = 1.0;
    const double q = x * x / 4.0;
    for (int k = 1; k < 500; ++k) {
        term *= q / ((double)k * (double)k);
        sum += term;
        if (term < 1e-17 * sum) break;
    }
    return sum;
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
    DCASE_CUDA_CHECK(cudaMalloc(&ctx->d_window2, kNfft * sizeof(float)));
    DCASE_CUDA_CHECK(cudaMemcpy(ctx->d_window2, win.data(), kNfft * sizeof(float), cudaMemcpyHostToDevice));
    {   // resampy.filters.sinc_window(num_zeros=64, precision=9, window=kaiser(14.769656459379492), rolloff=0.9475937167399596)
        const double beta = 14.769656459379492, rolloff = 0.9475937167399596;
        const int n = kRsZeros * kRsTable;
        std::vector<double> w(n + 1);
        const double i0b = bessel_i0(beta);
        for (int i = 0; i <= n; ++i) {
            const double xz = rolloff * ((double)kRsZeros * i / n);           // rolloff * linspace(0, num_zeros, n + 1)
            const double sinc = xz == 0.0 ? 1.0 : sin(M_PI * xz) / (M_PI * xz);
            const double r = (double)i / n;                                    // np.kaiser(2 n + 1, beta)[n + i]
            const double taper = bessel_i0(beta * sqrt(1.0 - r * r < 0.0 ? 0.0 : 1.0 - r * r)) / i0b;
            w[i] = taper * rolloff * sinc;
        }
        std::vector<float> tab(2 * (size_t)kRsWin, 0.f);
        for (int i = 0; i <= n; ++i) tab[i] = (float)w[i];
        for (int i = 0; i < n; ++i) tab[kRsWin + i] = (float)(w[i + 1] - w[i]);
        DCASE_CUDA_CHECK(cudaMalloc(&ctx->d_resample_win, tab.size() * sizeof(float)));
        DCASE_CUDA_CHECK(cudaMemcpy(ctx->d_resample_win, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    // balanced work split: band m gets n_m of the 128 items (proportional to its length, at least one); item q is a
    // contiguous run of weights inside its band.  The kernel reads the weights of item (lane + 32 j) transposed
    // ([j][i][lane]), zero padded to kMelRun and halved (its magnitudes are 2 |X[k]|: exact power-of-two rescale).
    {
        std::vector<int> n_slots(kMel, 1);
        int used = kMel;
        while (used < kMelItems) {      // give the next item to the band with the largest per-item load
            int best = 0;
            double best_load = -1.0;
            for (int m = 0; m < kMel; ++m) {
                const double load = (double)len[m] / n_slots[m];
                if (load > best_load) { best_load = load; best = m; }
            }
            ++n_slots[best];
            ++used;
        }
        std::vector<int> item_start(kMelItems, 0), owner(kMel * 2);
        std::vector<float> wt((size_t)4 * kMelRun * 32, 0.f);
        int slot = 0;
        for (int m = 0; m < kMel; ++m) {
            owner[2 * m] = slot;
            owner[2 * m + 1] = n_slots[m];
            for (int q = 0; q < n_slots[m]; ++q, ++slot) {
                const int beg = (int)((long long)len[m] * q / n_slots[m]), end = (int)((long long)len[m] * (q + 1) / n_slots[m]);
                if (end - beg > kMelRun) { dcase_set_error("mel work item longer than kMelRun"); return DCASE_ERR_STATE; }
                item_start[slot] = start[m] + beg;
                for (int i = 0; i < end - beg; ++i)
                    wt[((size_t)(slot >> 5) * kMelRun + i) * 32 + (slot & 31)] = 0.5f * packed[off[m] + beg + i];
            }
        }
        DCASE_CUDA_CHECK(cudaMalloc(&ctx->d_mel_wt, wt.size() * sizeof(float)));
        DCASE_CUDA_CHECK(cudaMalloc(&ctx->d_mel_start, item_start.size() * sizeof(int)));
        DCASE_CUDA_CHECK(cudaMalloc(&ctx->d_mel_owner, owner.size() * sizeof(int)));
        DCASE_CUDA_CHECK(cudaMemcpy(ctx->d_mel_wt, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice));
        DCASE_CUDA_CHECK(cudaMemcpy(ctx->d_mel_start, item_start.data(), item_start.size() * sizeof(int), cudaMemcpyHostToDevice));
        DCASE_CUDA_CHECK(cudaMemcpy(ctx->d_mel_owner, owner.data(), owner.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(stft_mel_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)kStftSmemBytes));
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(stft_mel_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)kStftSmemBytes));
    // two 112 KB CTAs per SM need the whole carve-out
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(stft_mel_kernel<float>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(stft_mel_kernel<int16_t>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    return DCASE_OK;
}

void dcase_logmel_tables_destroy(dcase_ctx* ctx) {
    cudaFree(ctx->d_window2); cudaFree(ctx->d_mel_wt); cudaFree(ctx->d_resample_win);
    cudaFree(ctx->d_mel_start); cudaFree(ctx->d_mel_owner);
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
    MelTables tab{ctx->d_window2, ctx->d_mel_wt, ctx->d_mel_start, (const int2*)ctx->d_mel_owner};
    const long long n_tiles = (long long)((T + kFramesPerTile - 1) / kFramesPerTile) * B;
    DCASE_REQUIRE(n_tiles < (1ll << 31), "too many frames for one launch");
    // persistent CTAs: two per SM (112 KB of shared memory each), each looping over 8-frame tiles
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

long long dcase_audio_resample_len(long long n_in, int sr_in, int sr_out) {
    if (n_in <= 0 || sr_in <= 0 || sr_out <= 0) return 0;
    return (long long)ceil((double)n_in * ((double)sr_out / (double)sr_in));
}

// test hook (no GPU): the time registers  sum_{j < t} 1 / ratio  as the device reconstructs them, into host memory
int dcase_audio_resample_clock(long long n, int sr_in, int sr_out, double* out_host) {
    DCASE_REQUIRE(n >= 0 && sr_in > 0 && sr_out > 0 && out_host, "bad argument");
    ResampleClock clk;
    DCASE_REQUIRE(build_resample_clock(n, 1.0 / ((double)sr_out / (double)sr_in), &clk) == 0, "clock table overflow");
    for (long long t = 0; t < n; ++t) {
        int sg = 0;
        for (int i = 1; i < clk.n; ++i) sg = clk.t0[i] <= t ? i : sg;
        out_host[t] = clk.s0[sg] + (double)(t - clk.t0[sg]) * clk.inc[sg];
    }
    return DCASE_OK;
}

int dcase_audio_resample(dcase_ctx* ctx, const float* mono, long long n_in, int sr_in, int sr_out, float* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DCASE_REQUIRE(ctx && mono && out, "null argument");
    DCASE_REQUIRE(n_in > 0 && sr_in > 0 && sr_out > 0, "bad shape / rate");
    const double ratio = (double)sr_out / (double)sr_in;
    const long long n_out = dcase_audio_resample_len(n_in, sr_in, sr_out);
    const long long n_valid = (long long)((double)n_in * ratio);              // resampy: int(n_in * ratio)
    DCASE_REQUIRE(n_valid >= 1, "input signal too short for the target rate");
    const double scale = ratio < 1.0 ? ratio : 1.0;
    const int index_step = (int)(scale * kRsTable);
    DCASE_REQUIRE(index_step >= 1, "sample ratio too small for the 512-entry filter table");
    ResampleClock clk;
    DCASE_REQUIRE(build_resample_clock(n_valid, 1.0 / ratio, &clk) == 0, "signal too long for the resampler's clock table");
    DCASE_PROF("audio_resample", stream);
    resample_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, stream>>>(mono, n_in, out, n_out, n_valid < n_out ? n_valid : n_out,
                                                                        clk, scale, index_step, ratio < 1.0 ? (float)ratio : 1.f,
                                                                        ctx->d_resample_win, ctx->d_resample_win + kRsWin);
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
