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

// Up to 8 independent GEMMs per launch (blockIdx.z = problem * split + k-slice):
//   C[m][n] (ldc) (+)= sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] (+ bias[n])
// mode 0: C = result; mode 1: atomicAdd into C (split-K and/or several problems sharing one C).
__global__ void __launch_bounds__(256)
sgemm_batch_kernel(GemmBatch g) {
    __shared__ float As[16][68];
    __shared__ float Bs[16][68];
    int pi = 0;
#pragma unroll
    for (int i = 1; i < 8; ++i) pi += (i < g.n && (int)blockIdx.z >= g.zoff[i]) ? 1 : 0;
    const GemmProblem& pr = g.p[pi];
    const int ks = blockIdx.z - g.zoff[pi];
    const int nsplit = g.zoff[pi + 1] - g.zoff[pi];
    const int M = pr.M, N = pr.N, K = pr.K;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    if (m0 >= M || n0 >= N) return;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int kchunk = ((K + nsplit - 1) / nsplit + 15) / 16 * 16;
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
__device__ __forceinline__ float sigmoid_ftz(float x) { return rcp_ftz(1.f + ex2_ftz(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_ftz(float x) { return fmaf(2.f, rcp_ftz(1.f + ex2_ftz(-2.8853900817779268f * x)), -1.f); }
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}
__device__ __forceinline__ void cp_async_wait_ahead() { asm volatile("cp.async.wait_group %0;" ::"n"(kRing - 2) : "memory"); }

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float hsum2(u64 v) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return lo + hi;
}
__device__ __forceinline__ u64 ffma2u(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ void cp_async4s(uint32_t dst_smem, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ ulonglong2 lds128u(uint32_t addr) {
    ulonglong2 v;
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }

__device__ __forceinline__ void cp_async16s(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Forward recurrence.  One CTA per (clip, direction); 256 threads = 64 hidden units x 4 lanes.  Lane q of unit i
// holds W_hh[g*64 + i][16q .. 16q+15] for the three gates g (48 registers) and reduces its partial dot products
// over the quad with two shuffles, so every lane of the quad has the three gate pre-activations and the gate math
// runs redundantly in registers: ONE block barrier per time step (h is double-buffered in shared memory).
// The whole [T][192] input-projection slab of the sequence (81 KB at T = 108) is copied into shared memory before
// the first step, so the 108-step chain contains no global load; everything in the loop is strength-reduced
// (32-bit shared addresses, element offsets advanced by a signed stride, branch-free lane roles).
__global__ void __launch_bounds__(256)
gru_fwd_kernel(GruFwdArgs a) {
    extern __shared__ __align__(16) float gru_smem[];
    float* hs = gru_smem;                    // [2][H]
    float* gis = gru_smem + 2 * H;           // [T][3H]
    const int tid = threadIdx.x;
    const int b = blockIdx.x, dir = blockIdx.y;
    const int T = a.T;
    const int BT = a.B * T;
    const int i = tid >> 2, q = tid & 3;
    const uint32_t hs_a = (uint32_t)__cvta_generic_to_shared(hs);
    const uint32_t gi_a = (uint32_t)__cvta_generic_to_shared(gis);
    {
        const float4* src = reinterpret_cast<const float4*>(a.gi + ((long long)dir * BT + (long long)b * T) * 3 * H);
        for (int k = tid; k < T * (3 * H / 4); k += 256) cp_async16s(gi_a + 16 * k, src + k);
    }
    const float* w_hh = dir ? a.w_hh[1] : a.w_hh[0];
    const float* b_hh = dir ? a.b_hh[1] : a.b_hh[0];
    u64 w2[3][8];
    float bh[3];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        const float4* wr = reinterpret_cast<const float4*>(w_hh + (g * H + i) * H + 16 * q);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
            const float4 v = __ldg(wr + k4);
            w2[g][2 * k4] = pack2(v.x, v.y);
            w2[g][2 * k4 + 1] = pack2(v.z, v.w);
        }
        bh[g] = __ldg(b_hh + g * H + i);
    }
    if (tid < H) hs[tid] = 0.f;
    const int dstep = dir ? -1 : 1;
    const int t_first = dir ? T - 1 : 0;
    // stores: lane 0: out + save_r; lane 1: save_z + save_n; lane 2: save_hn; lane 3: save_hp (element offsets)
    const bool saving = a.save_r != nullptr;
    float* pa = q == 0 ? a.out : (q == 1 ? a.save_z : (q == 2 ? a.save_hn : a.save_hp));
    float* pb = q == 0 ? a.save_r : a.save_n;
    const int row0 = b * T + t_first;
    int off_s = (dir * BT + row0) * H + i;             // offset into the save arrays
    int off_a = q == 0 ? row0 * 2 * H + dir * H + i : off_s;
    const int inc_s = dstep * H, inc_a = q == 0 ? dstep * 2 * H : dstep * H;
    const bool store_a = q == 0 || saving, store_b = saving && q < 2;
    // lane-role selectors as multipliers: va = hn, z, ghn or h_prev; vb = r or n
    const float ma0 = q == 0 ? 1.f : 0.f, ma1 = q == 1 ? 1.f : 0.f, ma2 = q == 2 ? 1.f : 0.f, ma3 = q == 3 ? 1.f : 0.f;
    cp_async_wait_all();
    __syncthreads();
    float h_own = 0.f;                      // h[i] of the previous step (every lane of the quad keeps it)
    uint32_t gaddr = gi_a + (uint32_t)(t_first * 3 * H + i) * 4;
    const int ginc = dstep * 3 * H * 4;
    uint32_t h_cur = hs_a + 64 * q, h_nxt = hs_a + H * 4 + 4 * i;
    const uint32_t h_flip = (hs_a + 64 * q) ^ (hs_a + H * 4 + 64 * q);
    const uint32_t n_flip = (hs_a + 4 * i) ^ (hs_a + H * 4 + 4 * i);
    for (int s = 0; s < T; ++s) {
        const float gi_r = lds32(gaddr), gi_z = lds32(gaddr + H * 4), gi_n = lds32(gaddr + 2 * H * 4);
        gaddr += ginc;
        u64 acc[3] = {0ull, 0ull, 0ull};
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
            const ulonglong2 h4 = lds128u(h_cur + 16 * k4);
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                acc[g] = ffma2u(w2[g][2 * k4], h4.x, acc[g]);
                acc[g] = ffma2u(w2[g][2 * k4 + 1], h4.y, acc[g]);
            }
        }
        const float gh_r = quad_sum(hsum2(acc[0])) + bh[0];
        const float gh_z = quad_sum(hsum2(acc[1])) + bh[1];
        const float ghn = quad_sum(hsum2(acc[2])) + bh[2];
        const float r = sigmoid_ftz(gi_r + gh_r), z = sigmoid_ftz(gi_z + gh_z);
        const float n = tanh_ftz(fmaf(r, ghn, gi_n));
        const float hn = fmaf(z, h_own - n, n);          // (1 - z) n + z h
        if (q == 0) sts32(h_nxt, hn);                    // every lane of the quad holds the same value: one writes (racecheck-clean)
        const float va = fmaf(ma0, hn, fmaf(ma1, z, fmaf(ma2, ghn, ma3 * h_own)));
        const float vb = q == 0 ? r : n;
        if (store_a) pa[off_a] = va;
        if (store_b) pb[off_s] = vb;
        off_a += inc_a; off_s += inc_s;
        h_own = hn;
        __syncthreads();                    // h of step s is visible
        h_cur ^= h_flip; h_nxt ^= n_flip;
    }
}

// Backward recurrence (BPTT), same quad layout: lane q of unit i holds W_hh[48q .. 48q+47][i] and reduces
// dh_prev[i] = sum_j W_hh[j][i] dgh[j] over the quad; the gate derivatives run redundantly in the quad.  dgh is
// double-buffered in shared memory: one block barrier per step.  The six [T][64] operand slabs of the sequence
// (d_out, r, z, n, hn, hp: 162 KB at T = 108) are copied into shared memory before the first step.
__global__ void __launch_bounds__(256)
gru_bwd_kernel(GruBwdArgs a) {
    extern __shared__ __align__(16) float gru_smem[];
    float* dgs = gru_smem;                   // [2][4 * 56]: dgh of a step, j -> (j / 48) * 56 + j % 48 (bank spread)
    float* ops = gru_smem + 2 * 4 * 56;      // [6][T][H]
    const int tid = threadIdx.x;
    const int b = blockIdx.x, dir = blockIdx.y;
    const int i = tid >> 2, q = tid & 3;
    const int T = a.T;
    const int BT = a.B * T;
    const uint32_t dgs_a = (uint32_t)__cvta_generic_to_shared(dgs);
    const uint32_t ops_a = (uint32_t)__cvta_generic_to_shared(ops);
    const uint32_t slab = (uint32_t)T * H * 4;
    {
        const long long so = ((long long)dir * BT + (long long)b * T) * H;
        const float* srcs[5] = {a.save_r + so, a.save_z + so, a.save_n + so, a.save_hn + so, a.save_hp + so};
        const int n16 = T * (H / 4);
        for (int k = tid; k < n16; k += 256) {
            const int t = k >> 4, c = k & 15;
            cp_async16s(ops_a + 16 * k, a.d_out + ((long long)(b * T + t) * 2 * H + dir * H) + 4 * c);
#pragma unroll
            for (int o = 0; o < 5; ++o) cp_async16s(ops_a + (o + 1) * slab + 16 * k, reinterpret_cast<const float4*>(srcs[o]) + k);
        }
    }
    const float* w_hh = dir ? a.w_hh[1] : a.w_hh[0];
    u64 w2[24];
#pragma unroll
    for (int jj = 0; jj < 24; ++jj)
        w2[jj] = pack2(__ldg(w_hh + (48 * q + 2 * jj) * H + i), __ldg(w_hh + (48 * q + 2 * jj + 1) * H + i));
    const int dstep = dir ? 1 : -1;                  // BPTT walks the sequence against the forward direction
    const int t_first = dir ? 0 : T - 1;
    const int row0 = b * T + t_first;
    // outputs of lane q: q0 -> dgi[i], dgh[i] = drp; q1 -> dgi[H+i], dgh[H+i] = dzp; q2 -> dgh[2H+i] = dghn;
    // q3 -> dgi[2H+i] = dn.  Lanes 0..2 also publish their value in dgs for the W_hh^T product.
    const int jg = q == 3 ? 2 * H + i : q * H + i;             // gate row written by this lane
    int off_g = (dir * BT + row0) * 3 * H + jg;
    const int inc_g = dstep * 3 * H;
    float* dgi_p = a.dgi; float* dgh_p = a.dgh;
    const uint32_t dg_slot = q == 3 ? (uint32_t)((3 * 56 + 48 + (i & 7)) * 4) : (uint32_t)(((jg / 48) * 56 + jg % 48) * 4);
    const float m0 = q == 0 ? 1.f : 0.f, m1 = q == 1 ? 1.f : 0.f, m2 = q == 2 ? 1.f : 0.f, m3 = q == 3 ? 1.f : 0.f;
    cp_async_wait_all();
    __syncthreads();
    float dh_rec = 0.f;                               // W_hh^T dgh of the previous step (all lanes of the quad)
    uint32_t oaddr = ops_a + (uint32_t)(t_first * H + i) * 4;
    const int oinc = dstep * H * 4;
    uint32_t dg_w = dgs_a + dg_slot, dg_r = dgs_a + (uint32_t)(q * 56 * 4);
    constexpr uint32_t dg_flip = 4 * 56 * 4;          // buffers at dgs_a and dgs_a + dg_flip
    uint32_t which = 0;
    for (int s = 0; s < T; ++s) {
        const float c_do = lds32(oaddr), c_r = lds32(oaddr + slab), c_z = lds32(oaddr + 2 * slab), c_n = lds32(oaddr + 3 * slab),
                    c_hn = lds32(oaddr + 4 * slab), c_hp = lds32(oaddr + 5 * slab);
        oaddr += oinc;
        const float dh = c_do + dh_rec;
        const float dn = dh * (1.f - c_z) * (1.f - c_n * c_n);
        const float dzp = dh * (c_hp - c_n) * c_z * (1.f - c_z);
        const float drp = dn * c_hn * c_r * (1.f - c_r);
        const float dghn = dn * c_r;
        const float v = fmaf(m0, drp, fmaf(m1, dzp, fmaf(m2, dghn, m3 * dn)));
        if (q != 2) dgi_p[off_g] = v;
        if (q != 3) dgh_p[off_g] = v;
        if (q != 3) sts32(dg_w + which, v);          // lane 3's value (dn) is not part of dgh
        off_g += inc_g;
        __syncthreads();
        u64 s0 = 0ull, s1 = 0ull;
#pragma unroll
        for (int j4 = 0; j4 < 12; ++j4) {
            const ulonglong2 d4 = lds128u(dg_r + which + 16 * j4);
            s0 = ffma2u(w2[2 * j4], d4.x, s0);
            s1 = ffma2u(w2[2 * j4 + 1], d4.y, s1);
        }
        dh_rec = fmaf(dh, c_z, quad_sum(hsum2(s0) + hsum2(s1)));   // direct path + recurrent part
        which ^= dg_flip;
    }
}

}  // namespace

int launch_sgemm_batch(const GemmBatch& g_in, cudaStream_t s) {
    DCASE_PROF("sgemm", s);
    GemmBatch g = g_in;
    int maxM = 0, maxN = 0;
    for (int i = 0; i < g.n; ++i) { maxM = g.p[i].M > maxM ? g.p[i].M : maxM; maxN = g.p[i].N > maxN ? g.p[i].N : maxN; }
    if (g.n <= 0 || maxM <= 0 || maxN <= 0) return DCASE_OK;
    g.zoff[0] = 0;
    for (int i = 0; i < 8; ++i) {
        const int sp = i < g.n ? (g.psplit[i] > 0 ? g.psplit[i] : g.split) : 0;
        g.zoff[i + 1] = g.zoff[i] + sp;
    }
    dim3 grid((maxN + 63) / 64, (maxM + 63) / 64, g.zoff[g.n]);
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

static size_t gru_fwd_smem(int T) { return (size_t)(2 * H + T * 3 * H) * sizeof(float); }
static size_t gru_bwd_smem(int T) { return (size_t)(2 * 4 * 56 + 6 * T * H) * sizeof(float); }
constexpr int kGruMaxT = 136;     // 6 * T * 64 * 4 B of operands must fit the 227 KB shared memory of one SM

int gru_kernels_init() {
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(gru_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gru_fwd_smem(kGruMaxT)));
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(gru_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gru_bwd_smem(kGruMaxT)));
    return DCASE_OK;
}

int launch_gru_fwd(const GruFwdArgs& a, cudaStream_t s) {
    DCASE_PROF("gru_fwd", s);
    DCASE_REQUIRE(a.T <= kGruMaxT, "sequence too long for the shared-memory resident GRU (T <= 136 output frames)");
    gru_fwd_kernel<<<dim3(a.B, 2), 256, gru_fwd_smem(a.T), s>>>(a);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int launch_gru_bwd(const GruBwdArgs& a, cudaStream_t s) {
    DCASE_PROF("gru_bwd", s);
    DCASE_REQUIRE(a.T <= kGruMaxT, "sequence too long for the shared-memory resident GRU (T <= 136 output frames)");
    gru_bwd_kernel<<<dim3(a.B, 2), 256, gru_bwd_smem(a.T), s>>>(a);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}
