// Head (dropout -> sigmoid dense + class-softmax attention pooling), mean-teacher losses, Adam + EMA.
//
// Replaces (reference file:line):
//   baseline/models/CRNN.py:74-81                         head forward (+ autograd backward)
//   baseline/main.py:95-145                               weak / strong BCE, 2x MSE consistency, teacher meters
//   baseline/main.py:152-157, :45-49 + torch.optim.Adam   optimizer step and teacher EMA
#include "head_loss.cuh"

namespace {

constexpr int kD = 128;        // 2 * hidden
constexpr int kMaxC = 16;      // class slots
constexpr int kHeadThreads = 256;
constexpr int kHeadRows = 128;    // rows of a clip staged per pass

template <int KC>
__device__ __forceinline__ void softmax_classes(const float (&ls)[KC], int NC, float (&a_raw)[KC]) {
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < KC; ++c) if (c < NC) mx = fmaxf(mx, ls[c]);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < KC; ++c) { a_raw[c] = c < NC ? __expf(ls[c] - mx) : 0.f; sum += a_raw[c]; }
    const float inv = 1.f / sum;
#pragma unroll
    for (int c = 0; c < KC; ++c) a_raw[c] *= inv;
}

__device__ __forceinline__ void load_head_weights(const HeadArgs& a, float* Wd, float* Ws, float* bd, float* bs) {
    // all 16 loads of a thread are requested before the first is stored: as a rolled loop every iteration waited out an L2
    // round trip (kHeadThreads = 256 threads, kMaxC * kD = 2048 floats per matrix)
    float wd[kMaxC * kD / kHeadThreads], wsv[kMaxC * kD / kHeadThreads];
#pragma unroll
    for (int k = 0; k < kMaxC * kD / kHeadThreads; ++k) {
        const int i = threadIdx.x + k * kHeadThreads;
        const bool ok = i < a.NC * kD;
        wd[k] = ok ? __ldg(a.w_dense + i) : 0.f;
        wsv[k] = ok ? __ldg(a.w_soft + i) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < kMaxC * kD / kHeadThreads; ++k) {
        Wd[threadIdx.x + k * kHeadThreads] = wd[k];
        Ws[threadIdx.x + k * kHeadThreads] = wsv[k];
    }
    if (threadIdx.x < kMaxC) {
        bd[threadIdx.x] = threadIdx.x < a.NC ? __ldg(a.b_dense + threadIdx.x) : 0.f;
        bs[threadIdx.x] = threadIdx.x < a.NC ? __ldg(a.b_soft + threadIdx.x) : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// One CTA per clip, 256 threads.  The clip's [To][128] BiGRU output is staged once in shared memory (coalesced loads,
// dropout applied on the way in, rows padded to 132 floats so that a warp's 32 rows hit distinct banks); thread
// (row, head) then owns the 10 logits of one head (dense / softmax) of one frame: 1280 FMAs against weights that
// are warp-broadcast from shared memory.  The previous thread-per-row version spent its time in un-coalesced row
// reads and 2560 serial FMAs per thread.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kXPitch = 132;

// x rows [t0, t0 + n_rows) of clip b -> xm[r][kXPitch] with inverted dropout (keep bits from Philox, one call per row)
__device__ __forceinline__ void stage_rows(const HeadArgs& a, int b, int t0, int n_rows, uint64_t seed, uint32_t step, float* xm,
                                           uint4* keep_s) {
    const int tid = threadIdx.x;
    if (a.drop && tid < n_rows) keep_s[tid] = philox4x32_10((uint64_t)((long long)b * a.To + t0 + tid), a.stream, step, seed);
    if (a.drop) __syncthreads();
    const float4* src = reinterpret_cast<const float4*>(a.x + ((long long)b * a.To + t0) * kD);
    constexpr int kIters = kHeadRows * (kD / 4) / kHeadThreads;      // 16: every load is in flight before the first use
    float4 vv[kIters];
#pragma unroll
    for (int k = 0; k < kIters; ++k) {
        const int i = tid + k * kHeadThreads;
        vv[k] = i < n_rows * (kD / 4) ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < kIters; ++k) {
        const int i = tid + k * kHeadThreads;
        if (i >= n_rows * (kD / 4)) break;
        const int r = i >> 5, k4 = i & 31;
        float4 v = vv[k];
        if (a.drop) {
            const uint4 kw4 = keep_s[r];
            const uint32_t kw = k4 < 8 ? kw4.x : (k4 < 16 ? kw4.y : (k4 < 24 ? kw4.z : kw4.w));
            const uint32_t bits = kw >> ((4 * k4) & 31);
            v.x = (bits & 1u) ? 2.f * v.x : 0.f;
            v.y = (bits & 2u) ? 2.f * v.y : 0.f;
            v.z = (bits & 4u) ? 2.f * v.z : 0.f;
            v.w = (bits & 8u) ? 2.f * v.w : 0.f;
        }
        *reinterpret_cast<float4*>(xm + r * kXPitch + 4 * k4) = v;
    }
}

// the kMaxC logits of one head for one staged row: acc[c] = bias[c] + sum_k xm_row[k] W[c][k]
template <int KC>
__device__ __forceinline__ void head_logits(const float* xm_row, const float* W, const float* bias, float (&acc)[KC]) {
#pragma unroll
    for (int c = 0; c < KC; ++c) acc[c] = bias[c];
#pragma unroll 2
    for (int k4 = 0; k4 < kD / 4; ++k4) {
        const float4 v = *reinterpret_cast<const float4*>(xm_row + 4 * k4);
#pragma unroll
        for (int c = 0; c < KC; ++c) {
            const float4 w = *reinterpret_cast<const float4*>(W + c * kD + 4 * k4);
            acc[c] = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, acc[c]))));
        }
    }
}

// dynamic smem: Wd, Ws [16][128] | bd, bs [16] | keep [128] uint4 | xm [128][132] | ex [128][2 * 16]
constexpr size_t kHeadSmem = (2 * kMaxC * kD + 2 * kMaxC) * sizeof(float) + kHeadRows * sizeof(uint4) +
                             (kHeadRows * kXPitch + kHeadRows * 2 * kMaxC) * sizeof(float);

template <int KC>      // class slots computed: 10 (cfg.crnn_kwargs) or kMaxC
__global__ void __launch_bounds__(kHeadThreads)
head_fwd_kernel(HeadArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* Wd = smem;
    float* Ws = Wd + kMaxC * kD;
    float* bd = Ws + kMaxC * kD;
    float* bs = bd + kMaxC;
    uint4* keep_s = reinterpret_cast<uint4*>(bs + kMaxC);
    float* xm = reinterpret_cast<float*>(keep_s + kHeadRows);
    float* ex = xm + kHeadRows * kXPitch;              // [row][0..15] = sigmoid(dense), [16..31] = clamped class softmax
    __shared__ float red[8][2 * kMaxC];
    const int tid = threadIdx.x, b = blockIdx.x;
    load_head_weights(a, Wd, Ws, bd, bs);
    uint64_t seed = a.seed; uint32_t step = a.step;
    if (a.sc) { seed = a.sc->seed; step = a.sc->step; }
    const int r = tid & 127, head = tid >> 7;          // head 0: dense -> strong; head 1: dense_softmax -> attention
    float num = 0.f, den = 0.f;                        // threads 0..NC-1 of the final pass own a class
    float pn[KC], pd[KC];
#pragma unroll
    for (int c = 0; c < KC; ++c) { pn[c] = 0.f; pd[c] = 0.f; }
    for (int t0 = 0; t0 < a.To; t0 += kHeadRows) {
        const int n_rows = min(kHeadRows, a.To - t0);
        __syncthreads();                               // weights staged / previous chunk consumed
        stage_rows(a, b, t0, n_rows, seed, step, xm, keep_s);
        __syncthreads();
        if (r < n_rows) {
            float acc[KC];
            head_logits<KC>(xm + r * kXPitch, head ? Ws : Wd, head ? bs : bd, acc);
            if (head == 0) {
                const long long row = (long long)b * a.To + t0 + r;
#pragma unroll
                for (int c = 0; c < KC; ++c) {
                    const float sg = sigmoid_fast(acc[c]);
                    ex[r * 2 * kMaxC + c] = sg;
                    if (c < a.NC) a.strong[row * a.NC + c] = sg;
                }
            } else {
                float ar[KC];
                softmax_classes<KC>(acc, a.NC, ar);
#pragma unroll
                for (int c = 0; c < KC; ++c) ex[r * 2 * kMaxC + kMaxC + c] = fminf(fmaxf(ar[c], 1e-7f), 1.f);
            }
        }
        __syncthreads();
        // attention pooling partial sums: thread (row, head 0) folds its row, then warps reduce per class
        if (head == 0 && r < n_rows) {
#pragma unroll
            for (int c = 0; c < KC; ++c) {
                const float at = ex[r * 2 * kMaxC + kMaxC + c];
                pn[c] = fmaf(ex[r * 2 * kMaxC + c], at, pn[c]);
                pd[c] += at;
            }
        }
    }
#pragma unroll
    for (int c = 0; c < KC; ++c) {
        const float n = warp_sum(pn[c]), d = warp_sum(pd[c]);
        if ((tid & 31) == 0) { red[tid >> 5][c] = n; red[tid >> 5][kMaxC + c] = d; }
    }
    __syncthreads();
    if (tid < a.NC) {
#pragma unroll
        for (int w = 0; w < 4; ++w) { num += red[w][tid]; den += red[w][kMaxC + tid]; }     // warps 0..3 hold head 0
        a.weak[b * a.NC + tid] = num / den;
        if (a.den) a.den[b * a.NC + tid] = den;
    }
}

__device__ __forceinline__ float bce_term(float p, float y) {
    return -(y * fmaxf(logf(p), -100.f) + (1.f - y) * fmaxf(log1pf(-p), -100.f));
}
__device__ __forceinline__ float bce_grad(float p, float y) { return (p - y) / fmaxf((1.f - p) * p, 1e-12f); }

// One CTA (256 threads) per clip, grid-stride over clips: element-wise gradients + per-CTA partial sums; the last CTA to
// finish (atomic ticket) adds the partials in CTA order, so the meters do not depend on the schedule.  A device function:
// the stand-alone kernel below, or the first phase of head_bwd_kernel (HeadArgs::loss), whose CTA b then consumes the
// gradients of clip b it has just written.
__device__ __forceinline__ void mt_loss_body(const LossArgs& a, float (*red)[6], bool* is_last_p) {
    const int tid = threadIdx.x;
    const float cw = a.sc ? a.sc->cons_weight : a.cons_weight;
    const bool has_t = a.strong_t != nullptr;
    const int per_clip = a.To * a.NC;
    const long long n_all_s = (long long)a.B * per_clip;
    const int n_all_w = a.B * a.NC;
    const float inv_ns = a.strong_hi > a.strong_lo ? 1.f / ((float)(a.strong_hi - a.strong_lo) * per_clip) : 0.f;
    const float inv_nw = a.weak_hi > a.weak_lo ? 1.f / ((float)(a.weak_hi - a.weak_lo) * a.NC) : 0.f;
    const float cs_scale = has_t ? cw * 2.f / (float)n_all_s : 0.f;
    const float cwk_scale = has_t ? cw * 2.f / (float)n_all_w : 0.f;
    float s_bce = 0.f, s_bce_t = 0.f, s_cons = 0.f, w_bce = 0.f, w_bce_t = 0.f, w_cons = 0.f;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        const bool in_strong = b >= a.strong_lo && b < a.strong_hi;
        const bool in_weak = b >= a.weak_lo && b < a.weak_hi;
        const long long base = (long long)b * per_clip;
        // the clip's elements in passes of up to 8 x 256: every load of a pass is requested before the first logarithm
        // (as a rolled loop each iteration waited out its own L2 round trips)
        for (int i0 = 0; i0 < per_clip; i0 += 8 * 256) {
            float ps[8], ys[8], ts[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = i0 + tid + 256 * k;
                const bool ok = i < per_clip;
                ps[k] = ok ? a.strong_s[base + i] : 0.5f;
                ys[k] = ok && in_strong ? a.target[base + i] : 0.f;
                ts[k] = ok && has_t ? a.strong_t[base + i] : 0.5f;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = i0 + tid + 256 * k;
                if (i >= per_clip) break;
            const long long e = base + i;
            const float p = ps[k];
            float d = 0.f;
            if (in_strong) {
                const float y = ys[k];
                s_bce += bce_term(p, y);
                if (has_t) s_bce_t += bce_term(ts[k], y);
                d = bce_grad(p, y) * inv_ns;
            }
            if (has_t) {
                const float diff = p - ts[k];
                s_cons = fmaf(diff, diff, s_cons);
                d = fmaf(cs_scale, diff, d);
            }
            a.d_strong[e] = d;
            }
        }
        // weak outputs of the clip: 16 lanes per class scan the clip's targets (target.max(-2), main.py:95)
        const int c = tid >> 4, l = tid & 15;
        if (c < a.NC) {
            const int e = b * a.NC + c;
            float y = -INFINITY;
            if (in_weak) {                      // passes of eight loads in flight (a rolled loop waited out an L2 round trip each)
                for (int t0 = l; t0 < a.To; t0 += 8 * 16) {
                    float yy[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int t = t0 + 16 * k;
                        yy[k] = t < a.To ? a.target[base + (long long)t * a.NC + c] : -INFINITY;
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) y = fmaxf(y, yy[k]);
                }
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) y = fmaxf(y, __shfl_xor_sync(0xffffffffu, y, o));
            if (l == 0) {
                const float p = a.weak_s[e];
                float d = 0.f;
                if (in_weak) {
                    w_bce += bce_term(p, y);
                    if (has_t) w_bce_t += bce_term(a.weak_t[e], y);
                    d = bce_grad(p, y) * inv_nw;
                }
                if (has_t) {
                    const float diff = p - a.weak_t[e];
                    w_cons = fmaf(diff, diff, w_cons);
                    d = fmaf(cwk_scale, diff, d);
                }
                a.d_weak[e] = d;
            }
        }
    }
    float v[6] = {w_bce, w_bce_t, s_bce, s_bce_t, s_cons, w_cons};
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        v[i] = warp_sum(v[i]);
        if ((tid & 31) == 0) red[tid >> 5][i] = v[i];
    }
    __syncthreads();
    if (tid < 6) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][tid];
        a.partials[blockIdx.x * 8 + tid] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) *is_last_p = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!*is_last_p) return;
    __threadfence();
    if (tid < 32) {
        float s = 0.f;
        if (tid < 6)
            for (unsigned g = 0; g < gridDim.x; ++g) s += __ldcg(a.partials + g * 8 + tid);
#pragma unroll
        for (int i = 0; i < 6; ++i) v[i] = __shfl_sync(0xffffffffu, s, i);
        if (tid == 0) {
            const float weak_loss = v[0] * inv_nw, weak_ema = v[1] * inv_nw;
            const float strong_loss = v[2] * inv_ns, strong_ema = v[3] * inv_ns;
            const float cons_s = has_t ? cw * v[4] / (float)n_all_s : 0.f;
            const float cons_w = has_t ? cw * v[5] / (float)n_all_w : 0.f;
            a.meters[0] = weak_loss; a.meters[1] = weak_ema; a.meters[2] = strong_loss; a.meters[3] = strong_ema;
            a.meters[4] = cons_s; a.meters[5] = cons_w;
            a.meters[6] = weak_loss + strong_loss + cons_s + cons_w;
            a.meters[7] = cw;
            *a.ticket = 0u;                                  // ready for the next launch
        }
    }
}

__global__ void __launch_bounds__(256)
mt_loss_kernel(LossArgs a) {
    __shared__ float red[8][6];
    __shared__ bool is_last;
    mt_loss_body(a, red, &is_last);
}

// Backward of the head for one clip per CTA; forward recomputed from the staged rows.  Thread (row, head) owns the
// logit gradients of its head; d_x needs both heads' gradients of the row (exchanged through shared memory), and the
// weight gradients are per-thread column sums over the staged chunk.
template <int KC>
__global__ void __launch_bounds__(kHeadThreads)
head_bwd_kernel(HeadArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* Wd = smem;
    float* Ws = Wd + kMaxC * kD;
    float* bd = Ws + kMaxC * kD;
    float* bs = bd + kMaxC;
    uint4* keep_s = reinterpret_cast<uint4*>(bs + kMaxC);
    float* xm = reinterpret_cast<float*>(keep_s + kHeadRows);
    float* dl = xm + kHeadRows * kXPitch;              // [row][0..15] dense logit grads, [16..31] softmax logit grads
    __shared__ float sig_s[kHeadRows][kMaxC];          // sigmoid(dense) of the chunk (softmax branch needs s - weak)
    const int tid = threadIdx.x, b = blockIdx.x;
    if (a.fused_loss) {                                // the losses and d Loss / d outputs of clip b first (main.py:95-145)
        __shared__ float loss_red[8][6];
        __shared__ bool loss_last;
        mt_loss_body(a.loss, loss_red, &loss_last);
        __syncthreads();                               // d_strong / d_weak of clip b, written by this CTA, are visible to it
    }
    load_head_weights(a, Wd, Ws, bd, bs);
    uint64_t seed = a.seed; uint32_t step = a.step;
    if (a.sc) { seed = a.sc->seed; step = a.sc->step; }
    const int r = tid & 127, head = tid >> 7;
    // weight gradients: thread tid owns column k = tid & 127 of head (tid >> 7): 16 accumulators + bias slot
    float gw[KC];
#pragma unroll
    for (int c = 0; c < KC; ++c) gw[c] = 0.f;
    float gb = 0.f;
    for (int t0 = 0; t0 < a.To; t0 += kHeadRows) {
        const int n_rows = min(kHeadRows, a.To - t0);
        __syncthreads();
        stage_rows(a, b, t0, n_rows, seed, step, xm, keep_s);
        __syncthreads();
        float acc[KC], ar[KC];
        const bool live = r < n_rows;
        const long long row = (long long)b * a.To + t0 + r;
        if (live) {
            head_logits<KC>(xm + r * kXPitch, head ? Ws : Wd, head ? bs : bd, acc);
            if (head == 0) {
#pragma unroll
                for (int c = 0; c < KC; ++c) sig_s[r][c] = sigmoid_fast(acc[c]);
            } else {
                softmax_classes<KC>(acc, a.NC, ar);
#pragma unroll
                for (int c = 0; c < KC; ++c) dl[r * 2 * kMaxC + kMaxC + c] = ar[c];       // parked: head 0 needs `at`
            }
        }
        __syncthreads();
        if (live) {
            if (head == 0) {
#pragma unroll
                for (int c = 0; c < KC; ++c) {
                    float dld = 0.f;
                    if (c < a.NC) {
                        const float sg = sig_s[r][c];
                        const float at = fminf(fmaxf(dl[r * 2 * kMaxC + kMaxC + c], 1e-7f), 1.f);
                        const float dwk = a.d_weak[b * a.NC + c];
                        const float inv_den = 1.f / __ldg(a.den + b * a.NC + c);
                        const float ds = a.d_strong[row * a.NC + c] + dwk * at * inv_den;
                        dld = ds * sg * (1.f - sg);
                    }
                    acc[c] = dld;
                }
            } else {
                float da[KC];
                float dot = 0.f;
#pragma unroll
                for (int c = 0; c < KC; ++c) {
                    da[c] = 0.f;
                    if (c < a.NC) {
                        const float dwk = a.d_weak[b * a.NC + c];
                        const float inv_den = 1.f / __ldg(a.den + b * a.NC + c);
                        const float wk = __ldg(a.weak + b * a.NC + c);
                        const bool pass = ar[c] >= 1e-7f && ar[c] <= 1.f;   // clamp passes gradient inside [min, max]
                        da[c] = pass ? dwk * (sig_s[r][c] - wk) * inv_den : 0.f;
                        dot = fmaf(da[c], ar[c], dot);
                    }
                }
#pragma unroll
                for (int c = 0; c < KC; ++c) acc[c] = c < a.NC ? ar[c] * (da[c] - dot) : 0.f;
            }
        }
        __syncthreads();                               // every head-0 thread has read its `at` values
#pragma unroll
        for (int c = 0; c < KC; ++c) dl[r * 2 * kMaxC + head * kMaxC + c] = live ? acc[c] : 0.f;
        __syncthreads();
        // d x = mask * 2 * (Wd^T dl_d + Ws^T dl_s): thread (row, head) writes channels 64 head .. 64 head + 63
        if (live) {
            const uint4 kw4 = a.drop ? keep_s[r] : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            const float* dl_row = dl + r * 2 * kMaxC;
            float4* dx = reinterpret_cast<float4*>(a.d_x + row * kD);
#pragma unroll 2
            for (int k4 = 16 * head; k4 < 16 * head + 16; ++k4) {
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int c = 0; c < KC; ++c) {
                    const float4 wd = *reinterpret_cast<const float4*>(Wd + c * kD + 4 * k4);
                    const float4 ws = *reinterpret_cast<const float4*>(Ws + c * kD + 4 * k4);
                    const float gd = dl_row[c], gs = dl_row[kMaxC + c];
                    o.x = fmaf(gd, wd.x, fmaf(gs, ws.x, o.x));
                    o.y = fmaf(gd, wd.y, fmaf(gs, ws.y, o.y));
                    o.z = fmaf(gd, wd.z, fmaf(gs, ws.z, o.z));
                    o.w = fmaf(gd, wd.w, fmaf(gs, ws.w, o.w));
                }
                if (a.drop) {
                    const uint32_t kw = k4 < 8 ? kw4.x : (k4 < 16 ? kw4.y : (k4 < 24 ? kw4.z : kw4.w));
                    const uint32_t bits = kw >> ((4 * k4) & 31);
                    o.x = (bits & 1u) ? 2.f * o.x : 0.f;
                    o.y = (bits & 2u) ? 2.f * o.y : 0.f;
                    o.z = (bits & 4u) ? 2.f * o.z : 0.f;
                    o.w = (bits & 8u) ? 2.f * o.w : 0.f;
                }
                dx[k4] = o;
            }
        }
        // weight gradients of the chunk: column k of head (tid >> 7)
        {
            const int k = tid & 127;
            for (int q = 0; q < n_rows; ++q) {
                const float xv = xm[q * kXPitch + k];
                const float* dq = dl + q * 2 * kMaxC + head * kMaxC;
#pragma unroll
                for (int c = 0; c < KC; ++c) gw[c] = fmaf(dq[c], xv, gw[c]);
            }
            if (k < KC)
                for (int q = 0; q < n_rows; ++q) gb += dl[q * 2 * kMaxC + head * kMaxC + k];
        }
    }
    {
        const int k = tid & 127;
        float* g_w = head ? a.g_w_soft : a.g_w_dense;
        float* g_b = head ? a.g_b_soft : a.g_b_dense;
#pragma unroll
        for (int c = 0; c < KC; ++c)
            if (c < a.NC) atomicAdd(g_w + c * kD + k, gw[c]);
        if (k < a.NC) atomicAdd(g_b + k, gb);
    }
}

__global__ void __launch_bounds__(256)
adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                float* __restrict__ p_ema, long long n, float lr, float beta1, float beta2, float eps, float bc1,
                float bc2, float alpha, float grad_scale, const DcaseStepScalars* __restrict__ sc) {
    if (sc) { lr = sc->lr; bc1 = sc->bias_corr1; bc2 = sc->bias_corr2; alpha = sc->ema_alpha; grad_scale = sc->grad_scale; }
    const float step_size = lr / bc1;
    const float inv_sqrt_bc2 = rsqrtf(bc2);
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float gi = g[i] * grad_scale;
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
        const float pi = p[i] - step_size * (mi / denom);
        p[i] = pi;
        if (p_ema) p_ema[i] = alpha * p_ema[i] + (1.f - alpha) * pi;
    }
}


}  // namespace

int head_kernels_init() {
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(head_fwd_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHeadSmem));
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(head_bwd_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHeadSmem));
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(head_fwd_kernel<kMaxC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHeadSmem));
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(head_bwd_kernel<kMaxC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHeadSmem));
    return DCASE_OK;
}

int launch_head_fwd(const HeadArgs& a, cudaStream_t s) {
    DCASE_PROF("head_fwd", s);
    if (a.NC <= 10) head_fwd_kernel<10><<<a.B, kHeadThreads, kHeadSmem, s>>>(a);      // 10 classes: 37 % fewer FMAs / loads than 16 slots
    else head_fwd_kernel<kMaxC><<<a.B, kHeadThreads, kHeadSmem, s>>>(a);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int launch_head_bwd(const HeadArgs& a, cudaStream_t s) {
    DCASE_PROF("head_bwd", s);
    if (a.NC <= 10) head_bwd_kernel<10><<<a.B, kHeadThreads, kHeadSmem, s>>>(a);
    else head_bwd_kernel<kMaxC><<<a.B, kHeadThreads, kHeadSmem, s>>>(a);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int launch_mt_loss(const LossArgs& a, cudaStream_t s) {
    DCASE_PROF("mt_loss", s);
    const int grid = a.B < kLossMaxCtas ? a.B : kLossMaxCtas;
    mt_loss_kernel<<<grid, 256, 0, s>>>(a);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int launch_adam_ema(float* p, const float* g, float* m, float* v, float* p_ema, long long n, float lr, float beta1,
                    float beta2, float eps, float bc1, float bc2, float alpha, float grad_scale,
                    const DcaseStepScalars* sc, int num_sms, cudaStream_t s) {
    DCASE_PROF("adam_ema", s);
    long long blocks = (n + 255) / 256;
    if (blocks > num_sms * 8) blocks = num_sms * 8;
    adam_ema_kernel<<<(int)blocks, 256, 0, s>>>(p, g, m, v, p_ema, n, lr, beta1, beta2, eps, bc1, bc2, alpha,
                                               grad_scale, sc);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}
