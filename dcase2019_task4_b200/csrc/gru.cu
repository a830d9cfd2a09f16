// Bidirectional GRU (nn.GRU gate order r,z,n; h0 = 0) forward and BPTT, plus the small batched SGEMM and
// column-sum helpers used for the input projections and the weight gradients.
//
// Replaces (reference file:line): baseline/models/RNN.py:7-16 (BidirectionalGRU -> nn.GRU) and its backward.
//
// The recurrence is latency bound (T = 108 strictly sequential [B,64]x[64,192] products): one CTA per
// (clip, direction).  Forward: 192 threads own one gate row of W_hh each (registers, packed FFMA2), 64 threads do
// the gate math; h is double-buffered in shared memory.  Backward: 64 threads do the gate derivatives, all 256
// the W_hh^T product; every per-step operand is prefetched one step ahead so no global-load latency sits on the
// sequential chain.
#include "gru.cuh"

namespace {

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

// cp.async ring: per-step operands are staged several steps ahead (no register holds an in-flight load, so the
// sequential chain never waits on global-memory latency)
constexpr int kRing = 8;
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_ring() { asm volatile("cp.async.wait_group %0;" ::"n"(kRing - 1) : "memory"); }

// Up to 8 independent GEMMs per launch (blockIdx.z = problem * split + k-slice):
//   C[m][n] (ldc) (+)= sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] (+ bias[n])
// mode 0: C = result; mode 1: atomicAdd into C (split-K and/or several problems sharing one C).
__global__ void __launch_bounds__(256)
sgemm_batch_kernel(GemmBatch g) {
    __shared__ float As[16][68];
    __shared__ float Bs[16][68];
    const GemmProblem& pr = g.p[blockIdx.z / g.split];
    const int ks = blockIdx.z % g.split;
    const int M = pr.M, N = pr.N, K = pr.K;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    if (m0 >= M || n0 >= N) return;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int kchunk = ((K + g.split - 1) / g.split + 15) / 16 * 16;
    const int kbeg = ks * kchunk;
    const int kend = min(K, kbeg + kchunk);
    const float* __restrict__ A = pr.A;
    const float* __restrict__ B = pr.B;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + i * 256;
            {   // A tile: pick the index order that makes the global reads contiguous
                const int mm = pr.sak == 1 ? idx >> 4 : idx & 63, kk = pr.sak == 1 ? idx & 15 : idx >> 6;
                const int m = m0 + mm, k = k0 + kk;
                As[kk][mm] = (m < M && k < kend) ? __ldg(A + m * pr.sam + k * pr.sak) : 0.f;
            }
            {
                const int kk = pr.sbn == 1 ? idx >> 6 : idx & 15, nn = pr.sbn == 1 ? idx & 63 : idx >> 4;
                const int n = n0 + nn, k = k0 + kk;
                Bs[kk][nn] = (n < N && k < kend) ? __ldg(B + k * pr.sbk + n * pr.sbn) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
            const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (pr.bias && ks == 0) v += __ldg(pr.bias + n);
            float* dst = pr.C + (long long)m * pr.ldc + n;
            if (g.mode == 1) atomicAdd(dst, v); else *dst = v;
        }
    }
}

// out_q[n] += sum_m A_q[m][n] for up to 4 matrices per launch; grid = (ceil(N/64), row slices, n_matrices)
__global__ void __launch_bounds__(256)
colsum_batch_kernel(ColsumBatch c) {
    __shared__ float red[4][64];
    const float* __restrict__ A = c.A[blockIdx.z];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int n = blockIdx.x * 64 + tx;
    const int rows_per = (c.M + gridDim.y - 1) / gridDim.y;
    const int mbeg = blockIdx.y * rows_per, mend = min(c.M, mbeg + rows_per);
    float s = 0.f;
    if (n < c.N)
        for (int m = mbeg + ty; m < mend; m += 4) s += __ldg(A + (long long)m * c.N + n);
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && n < c.N) atomicAdd(c.out[blockIdx.z] + n, red[0][tx] + red[1][tx] + red[2][tx] + red[3][tx]);
}

constexpr int H = 64;

__global__ void __launch_bounds__(256)
gru_fwd_kernel(GruFwdArgs a) {
    __shared__ __align__(16) float hs[2][H];
    __shared__ float pre[2 * H];        // gi + gh of the r and z rows
    __shared__ float ghn_s[H], gin_s[H];
    __shared__ float gi_ring[kRing][3 * H];
    const int tid = threadIdx.x;
    const int b = blockIdx.x, dir = blockIdx.y;
    const int BT = a.B * a.T;
    const bool mat = tid < 3 * H;       // warps 0..5: one gate row each; warps 6..7: gate math for one unit each
    const int i = tid - 3 * H;
    float2 w2[H / 2];
    float bh = 0.f;
    if (mat) {
        const float4* wr = reinterpret_cast<const float4*>(a.w_hh[dir] + tid * H);
#pragma unroll
        for (int k4 = 0; k4 < H / 4; ++k4) {
            const float4 v = __ldg(wr + k4);
            w2[2 * k4] = make_float2(v.x, v.y);
            w2[2 * k4 + 1] = make_float2(v.z, v.w);
        }
        bh = __ldg(a.b_hh[dir] + tid);
    }
    if (tid < H) hs[0][tid] = 0.f;
    __syncthreads();
    const float* gi = a.gi + ((long long)dir * BT + (long long)b * a.T) * 3 * H;
    int cur = 0;
    int t = dir ? a.T - 1 : 0;
    if (mat) {
        for (int d = 0; d < kRing; ++d) {
            if (d < a.T) cp_async4(&gi_ring[d][tid], gi + (long long)(dir ? a.T - 1 - d : d) * 3 * H + tid);
            cp_async_commit();
        }
    }
    for (int s = 0; s < a.T; ++s) {
        const long long row = (long long)b * a.T + t;
        const int t_next = dir ? t - 1 : t + 1;
        if (mat) {
            cp_async_wait_ring();
            const float gi_v = gi_ring[s % kRing][tid];
            {
                const int sn = s + kRing;
                if (sn < a.T) cp_async4(&gi_ring[s % kRing][tid], gi + (long long)(dir ? a.T - 1 - sn : sn) * 3 * H + tid);
                cp_async_commit();
            }
            float2 s0 = make_float2(0.f, 0.f), s1 = s0;
            const float4* hv = reinterpret_cast<const float4*>(hs[cur]);
#pragma unroll
            for (int k4 = 0; k4 < H / 4; ++k4) {
                const float4 h4 = hv[k4];
                s0 = ffma2(w2[2 * k4], make_float2(h4.x, h4.y), s0);
                s1 = ffma2(w2[2 * k4 + 1], make_float2(h4.z, h4.w), s1);
            }
            const float gh = bh + ((s0.x + s0.y) + (s1.x + s1.y));
            if (tid < 2 * H) pre[tid] = gi_v + gh;
            else { ghn_s[tid - 2 * H] = gh; gin_s[tid - 2 * H] = gi_v; }
        }
        __syncthreads();
        if (!mat) {
            const float r = sigmoid_fast(pre[i]), z = sigmoid_fast(pre[H + i]);
            const float ghn = ghn_s[i];
            const float n = tanh_fast(gin_s[i] + r * ghn);
            const float hp = hs[cur][i];
            const float hn = (1.f - z) * n + z * hp;
            hs[cur ^ 1][i] = hn;
            a.out[row * 2 * H + dir * H + i] = hn;
            if (a.save_r) {
                const long long o = ((long long)dir * BT + row) * H + i;
                a.save_r[o] = r; a.save_z[o] = z; a.save_n[o] = n; a.save_hn[o] = ghn; a.save_hp[o] = hp;
            }
        }
        __syncthreads();
        cur ^= 1;
        t = t_next;
    }
}

__global__ void __launch_bounds__(256)
gru_bwd_kernel(GruBwdArgs a) {
    __shared__ __align__(16) float dgs[4 * 56];     // dgh of this step, j -> (j / 48) * 56 + j % 48 (bank spread)
    __shared__ float dhc[H];                         // W_hh^T dgh: recurrent part of dh for the next step
    __shared__ float ring[kRing][6][H];              // d_out, r, z, n, hn, hp of the coming steps
    const int tid = threadIdx.x;
    const int b = blockIdx.x, dir = blockIdx.y;
    const int i = tid >> 2, q = tid & 3;
    const int BT = a.B * a.T;
    float2 w2[24];
#pragma unroll
    for (int jj = 0; jj < 24; ++jj)
        w2[jj] = make_float2(__ldg(a.w_hh[dir] + (48 * q + 2 * jj) * H + i), __ldg(a.w_hh[dir] + (48 * q + 2 * jj + 1) * H + i));
    if (tid < H) dhc[tid] = 0.f;
    const long long sbase = (long long)dir * BT;
    const bool gate = tid < H;                       // warps 0..1: gate derivatives of unit `tid`
    auto stage = [&](int s_idx) {                    // gate thread `tid` stages its six operands of step s_idx
        const int t = dir ? s_idx : a.T - 1 - s_idx;
        const long long row = (long long)b * a.T + t;
        const long long o = (sbase + row) * H + tid;
        float (*slot)[H] = ring[s_idx % kRing];
        cp_async4(&slot[0][tid], a.d_out + row * 2 * H + dir * H + tid);
        cp_async4(&slot[1][tid], a.save_r + o);
        cp_async4(&slot[2][tid], a.save_z + o);
        cp_async4(&slot[3][tid], a.save_n + o);
        cp_async4(&slot[4][tid], a.save_hn + o);
        cp_async4(&slot[5][tid], a.save_hp + o);
    };
    if (gate) {
        for (int d = 0; d < kRing; ++d) {
            if (d < a.T) stage(d);
            cp_async_commit();
        }
    }
    __syncthreads();
    for (int s = 0; s < a.T; ++s) {
        const int t = dir ? s : a.T - 1 - s;
        if (gate) {
            cp_async_wait_ring();
            float (*slot)[H] = ring[s % kRing];
            const float c_do = slot[0][tid], c_r = slot[1][tid], c_z = slot[2][tid], c_n = slot[3][tid],
                        c_hn = slot[4][tid], c_hp = slot[5][tid];
            if (s + kRing < a.T) stage(s + kRing);
            cp_async_commit();
            const long long row = (long long)b * a.T + t;
            const float dh = c_do + dhc[tid];
            const float dn = dh * (1.f - c_z) * (1.f - c_n * c_n);
            const float dzp = dh * (c_hp - c_n) * c_z * (1.f - c_z);
            const float drp = dn * c_hn * c_r * (1.f - c_r);
            const float dghn = dn * c_r;
            float* gi = a.dgi + (sbase + row) * 3 * H;
            float* gh = a.dgh + (sbase + row) * 3 * H;
            gi[tid] = drp; gi[H + tid] = dzp; gi[2 * H + tid] = dn;
            gh[tid] = drp; gh[H + tid] = dzp; gh[2 * H + tid] = dghn;
            const int j0 = tid, j1 = H + tid, j2 = 2 * H + tid;
            dgs[(j0 / 48) * 56 + j0 % 48] = drp;
            dgs[(j1 / 48) * 56 + j1 % 48] = dzp;
            dgs[(j2 / 48) * 56 + j2 % 48] = dghn;
            dhc[tid] = dh * c_z;                     // direct path; the recurrent part is added below
        }
        __syncthreads();
        const float4* dv = reinterpret_cast<const float4*>(&dgs[q * 56]);
        float2 s0 = make_float2(0.f, 0.f), s1 = s0;
#pragma unroll
        for (int j4 = 0; j4 < 12; ++j4) {
            const float4 d4 = dv[j4];
            s0 = ffma2(w2[2 * j4], make_float2(d4.x, d4.y), s0);
            s1 = ffma2(w2[2 * j4 + 1], make_float2(d4.z, d4.w), s1);
        }
        float part = (s0.x + s0.y) + (s1.x + s1.y);
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        if (q == 0) dhc[i] += part;
        __syncthreads();
    }
}

}  // namespace

int launch_sgemm_batch(const GemmBatch& g, cudaStream_t s) {
    DCASE_PROF("sgemm", s);
    int maxM = 0, maxN = 0;
    for (int i = 0; i < g.n; ++i) { maxM = g.p[i].M > maxM ? g.p[i].M : maxM; maxN = g.p[i].N > maxN ? g.p[i].N : maxN; }
    if (g.n <= 0 || maxM <= 0 || maxN <= 0) return DCASE_OK;
    dim3 grid((maxN + 63) / 64, (maxM + 63) / 64, g.n * g.split);
    sgemm_batch_kernel<<<grid, 256, 0, s>>>(g);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int launch_colsum_batch(const ColsumBatch& c, cudaStream_t s) {
    DCASE_PROF("colsum", s);
    int slices = (c.M + 127) / 128;
    if (slices > 32) slices = 32;
    if (slices < 1) slices = 1;
    colsum_batch_kernel<<<dim3((c.N + 63) / 64, slices, c.n), 256, 0, s>>>(c);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int launch_gru_fwd(const GruFwdArgs& a, cudaStream_t s) {
    DCASE_PROF("gru_fwd", s);
    gru_fwd_kernel<<<dim3(a.B, 2), 256, 0, s>>>(a);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int launch_gru_bwd(const GruBwdArgs& a, cudaStream_t s) {
    DCASE_PROF("gru_bwd", s);
    gru_bwd_kernel<<<dim3(a.B, 2), 256, 0, s>>>(a);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}
