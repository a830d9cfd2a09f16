// Bidirectional GRU (nn.GRU gate order r,z,n; h0 = 0) forward and BPTT, plus the small strided SGEMM and
// column-sum helpers used for the input projections and the weight gradients.
//
// Replaces (reference file:line): baseline/models/RNN.py:7-16 (BidirectionalGRU -> nn.GRU) and its backward.
//
// The recurrence is latency bound (T = 108 strictly sequential [B,64]x[64,192] products): one CTA per
// (clip, direction), W_hh resident in registers (one gate row per thread), h double-buffered in shared
// memory, one __syncthreads per time step.
#include "gru.cuh"

namespace {

// C[m][n] (ldc) = beta*C + sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] (+ bias[n]);  split-K over gridDim.z
// accumulates with atomics (then C must hold the initial value and beta is ignored).
__global__ void __launch_bounds__(256)
sgemm_kernel(int M, int N, int K, const float* __restrict__ A, long long sam, long long sak,
             const float* __restrict__ B, long long sbk, long long sbn, float* __restrict__ C, int ldc,
             const float* __restrict__ bias, int beta) {
    __shared__ float As[16][68];
    __shared__ float Bs[16][68];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int kchunk = ((K + gridDim.z - 1) / gridDim.z + 15) / 16 * 16;
    const int kbeg = blockIdx.z * kchunk;
    const int kend = min(K, kbeg + kchunk);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + i * 256;
            {
                const int mm = idx >> 4, kk = idx & 15;
                const int m = m0 + mm, k = k0 + kk;
                As[kk][mm] = (m < M && k < kend) ? __ldg(A + m * sam + k * sak) : 0.f;
            }
            {
                const int kk = idx >> 6, nn = idx & 63;
                const int n = n0 + nn, k = k0 + kk;
                Bs[kk][nn] = (n < N && k < kend) ? __ldg(B + k * sbk + n * sbn) : 0.f;
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
            if (bias && blockIdx.z == 0) v += __ldg(bias + n);
            float* dst = C + (long long)m * ldc + n;
            if (gridDim.z > 1) atomicAdd(dst, v);
            else *dst = beta ? *dst + v : v;
        }
    }
}

// out[n] += sum_m A[m][n]
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ A, int M, int N, float* __restrict__ out) {
    __shared__ float red[4][64];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int n = blockIdx.x * 64 + tx;
    float s = 0.f;
    if (n < N)
        for (int m = ty; m < M; m += 4) s += __ldg(A + (long long)m * N + n);
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && n < N) atomicAdd(out + n, red[0][tx] + red[1][tx] + red[2][tx] + red[3][tx]);
}

constexpr int H = 64;

__global__ void __launch_bounds__(4 * H)
gru_fwd_kernel(GruFwdArgs a) {
    __shared__ __align__(16) float hs[2][H];
    const int tid = threadIdx.x;
    const int b = blockIdx.x, dir = blockIdx.y;
    const int i = tid >> 2, g = tid & 3;
    const int lane = tid & 31, base = lane & ~3;
    const int BT = a.B * a.T;
    float w[H];
    float bh = 0.f;
    if (g < 3) {
        const float4* wr = reinterpret_cast<const float4*>(a.w_hh[dir] + (g * H + i) * H);
#pragma unroll
        for (int k4 = 0; k4 < H / 4; ++k4) {
            const float4 v = __ldg(wr + k4);
            w[4 * k4] = v.x; w[4 * k4 + 1] = v.y; w[4 * k4 + 2] = v.z; w[4 * k4 + 3] = v.w;
        }
        bh = __ldg(a.b_hh[dir] + g * H + i);
    } else {
#pragma unroll
        for (int k = 0; k < H; ++k) w[k] = 0.f;
    }
    if (tid < H) hs[0][tid] = 0.f;
    __syncthreads();
    const float* gi = a.gi + (long long)dir * BT * 3 * H;
    int cur = 0;
    const int gcol = (g < 3 ? g : 0) * H + i;
    int t = dir ? a.T - 1 : 0;
    float gi_v = __ldg(gi + ((long long)b * a.T + t) * 3 * H + gcol);
    for (int s = 0; s < a.T; ++s) {
        const long long row = (long long)b * a.T + t;
        const int t_next = dir ? t - 1 : t + 1;
        float gi_next = 0.f;
        if (s + 1 < a.T) gi_next = __ldg(gi + ((long long)b * a.T + t_next) * 3 * H + gcol);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        const float4* hv = reinterpret_cast<const float4*>(hs[cur]);
#pragma unroll
        for (int k4 = 0; k4 < H / 4; ++k4) {
            const float4 h4 = hv[k4];
            s0 = fmaf(w[4 * k4], h4.x, s0);
            s1 = fmaf(w[4 * k4 + 1], h4.y, s1);
            s2 = fmaf(w[4 * k4 + 2], h4.z, s2);
            s3 = fmaf(w[4 * k4 + 3], h4.w, s3);
        }
        const float gh = bh + ((s0 + s1) + (s2 + s3));
        const float v = gi_v + gh;
        const float rp = __shfl_sync(0xffffffffu, v, base);
        const float zp = __shfl_sync(0xffffffffu, v, base + 1);
        const float ghn = __shfl_sync(0xffffffffu, gh, base + 2);
        const float gin = __shfl_sync(0xffffffffu, gi_v, base + 2);
        if (g == 0) {
            const float r = sigmoid_fast(rp), z = sigmoid_fast(zp);
            const float n = tanh_fast(gin + r * ghn);
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
        gi_v = gi_next;
        t = t_next;
    }
}

__global__ void __launch_bounds__(4 * H)
gru_bwd_kernel(GruBwdArgs a) {
    __shared__ __align__(16) float dgs[2][4 * 56];
    const int tid = threadIdx.x;
    const int b = blockIdx.x, dir = blockIdx.y;
    const int i = tid >> 2, q = tid & 3;
    const int BT = a.B * a.T;
    float w[48];
#pragma unroll
    for (int jj = 0; jj < 48; ++jj) w[jj] = __ldg(a.w_hh[dir] + (48 * q + jj) * H + i);
    float dh_carry = 0.f;
    int buf = 0;
    const long long sbase = (long long)dir * BT;
    for (int s = 0; s < a.T; ++s) {
        const int t = dir ? s : a.T - 1 - s;
        const long long row = (long long)b * a.T + t;
        float dhz = 0.f;
        if (q == 0) {
            const long long o = (sbase + row) * H + i;
            const float dh = __ldg(a.d_out + row * 2 * H + dir * H + i) + dh_carry;
            const float r = a.save_r[o], z = a.save_z[o], n = a.save_n[o], hn = a.save_hn[o], hp = a.save_hp[o];
            const float dn = dh * (1.f - z) * (1.f - n * n);
            const float dzp = dh * (hp - n) * z * (1.f - z);
            const float drp = dn * hn * r * (1.f - r);
            const float dghn = dn * r;
            float* gi = a.dgi + (sbase + row) * 3 * H;
            float* gh = a.dgh + (sbase + row) * 3 * H;
            gi[i] = drp; gi[H + i] = dzp; gi[2 * H + i] = dn;
            gh[i] = drp; gh[H + i] = dzp; gh[2 * H + i] = dghn;
            // slot(j) = (j / 48) * 56 + j % 48
            const int j0 = i, j1 = H + i, j2 = 2 * H + i;
            dgs[buf][(j0 / 48) * 56 + j0 % 48] = drp;
            dgs[buf][(j1 / 48) * 56 + j1 % 48] = dzp;
            dgs[buf][(j2 / 48) * 56 + j2 % 48] = dghn;
            dhz = dh * z;
        }
        __syncthreads();
        const float4* dv = reinterpret_cast<const float4*>(&dgs[buf][q * 56]);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int j4 = 0; j4 < 12; ++j4) {
            const float4 d4 = dv[j4];
            s0 = fmaf(w[4 * j4], d4.x, s0);
            s1 = fmaf(w[4 * j4 + 1], d4.y, s1);
            s2 = fmaf(w[4 * j4 + 2], d4.z, s2);
            s3 = fmaf(w[4 * j4 + 3], d4.w, s3);
        }
        float part = (s0 + s1) + (s2 + s3);
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        if (q == 0) dh_carry = dhz + part;
        buf ^= 1;
    }
}

}  // namespace

int launch_sgemm(int M, int N, int K, const float* A, long long sam, long long sak, const float* B, long long sbk,
                 long long sbn, float* C, int ldc, const float* bias, int beta, int split_k, cudaStream_t s) {
    DCASE_PROF("sgemm", s);
    if (M <= 0 || N <= 0 || K <= 0) return DCASE_OK;
    dim3 grid((N + 63) / 64, (M + 63) / 64, split_k < 1 ? 1 : split_k);
    sgemm_kernel<<<grid, 256, 0, s>>>(M, N, K, A, sam, sak, B, sbk, sbn, C, ldc, bias, beta);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int launch_colsum(const float* A, int M, int N, float* out, cudaStream_t s) {
    DCASE_PROF("colsum", s);
    colsum_kernel<<<(N + 63) / 64, 256, 0, s>>>(A, M, N, out);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int launch_gru_fwd(const GruFwdArgs& a, cudaStream_t s) {
    DCASE_PROF("gru_fwd", s);
    gru_fwd_kernel<<<dim3(a.B, 2), 4 * H, 0, s>>>(a);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int launch_gru_bwd(const GruBwdArgs& a, cudaStream_t s) {
    DCASE_PROF("gru_bwd", s);
    gru_bwd_kernel<<<dim3(a.B, 2), 4 * H, 0, s>>>(a);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}
