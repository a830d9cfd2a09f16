// BiGRU forward for hidden sizes 128 / 256 (BASELINE.json configs[4]: the BiGRU sweep) -- EXPERIMENTAL.
//
// Replaces (reference file:line): baseline/models/RNN.py:7-16 (nn.GRU(n_in, H, num_layers=2, bidirectional=True,
// batch_first=True)) for the hidden sizes cfg.crnn_kwargs does not select.  H = 64 stays with gru.cu.
//
// W_hh (3H x H fp32: 196 KB at H = 128, 786 KB at H = 256) no longer fits one CTA, so the recurrence of one
// (direction, group of 4 clips) runs on a THREAD-BLOCK CLUSTER of H / 32 CTAs (4 or 8).  CTA `rank` owns hidden units
// [32 rank, 32 rank + 32): thread (unit, k-lane) keeps the 3 x H/8 weights of its unit's r / z / n rows for its k-lane in
// REGISTERS (96 at H = 256; lane kl holds columns {32 j + 4 kl + c}, so the 8 lanes of a unit read one contiguous,
// conflict-free 128-byte line of h per j), multiplies them with the previous h of the 4 clips from shared memory, and
// reduces over the 8 lanes with three shuffles.  Every lane then has the three gate pre-activations, the gate math runs
// redundantly, and lane kl stores the new h[unit] of every clip straight into the h buffer of CTA kl of the cluster
// (distributed shared memory): after ONE cluster barrier per step every CTA holds the complete new h.  h is double
// buffered, the input projections of the next step are prefetched into registers.
//
// STATUS (round 1): compiles for sm_100a; written after the round's GPU budget was spent, so it has not run on
// hardware.  tests/test_gpu_bigru.py covers it only under DCASE_EXPERIMENTAL=1.
#include <cooperative_groups.h>

#include "../../include/dcase_b200.h"
#include "ctx.h"
#include "gru.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kUnits = 32;   // hidden units per CTA (256 threads = 32 units x 8 k-lanes)
constexpr int kBG = 4;       // clips per cluster

__device__ __forceinline__ float ex2c(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcpc(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sigmoid_c(float x) { return rcpc(1.f + ex2c(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_c(float x) { return fmaf(2.f, rcpc(1.f + ex2c(-2.8853900817779268f * x)), -1.f); }

template <int H>
__global__ void __launch_bounds__(256)
gru_cluster_fwd_kernel(const float* __restrict__ gi,        // [2][B*T][3H] input projections incl. b_ih
                       const float* __restrict__ w_hh0, const float* __restrict__ w_hh1,   // [3H][H] per direction
                       const float* __restrict__ b_hh0, const float* __restrict__ b_hh1,   // [3H]
                       float* __restrict__ out,             // [B*T][2H]
                       int B, int T) {
    constexpr int CL = H / kUnits;       // CTAs per cluster
    constexpr int NJ = H / 32;           // 32-column blocks of a weight row; a lane holds 4 columns of each
    static_assert(CL >= 1 && CL <= 8, "portable cluster size");
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    __shared__ __align__(16) float hs[2][kBG][H];
    const int tid = threadIdx.x, kl = tid & 7, ul = tid >> 3;
    const int u = rank * kUnits + ul;
    const int dir = blockIdx.z;
    const int b0 = blockIdx.y * kBG;
    const float* w_hh = dir ? w_hh1 : w_hh0;
    const float* b_hh = dir ? b_hh1 : b_hh0;

    float w[3][NJ][4];
    float bh[3];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(w_hh + (size_t)(g * H + u) * H + 32 * j + 4 * kl));
            w[g][j][0] = v.x; w[g][j][1] = v.y; w[g][j][2] = v.z; w[g][j][3] = v.w;
        }
        bh[g] = __ldg(b_hh + g * H + u);
    }
    for (int i = tid; i < 2 * kBG * H; i += 256) (&hs[0][0][0])[i] = 0.f;
    cluster.sync();                      // every CTA's h buffers are zero before the first remote store lands

    // lane kl of every unit forwards the unit's new h to CTA kl of the cluster (lanes >= CL have no target)
    float* remote = kl < CL ? cluster.map_shared_rank(&hs[0][0][0], kl) : nullptr;

    const size_t BT = (size_t)B * T;
    const int dstep = dir ? -1 : 1;
    int t = dir ? T - 1 : 0;
    float h_own[kBG], gcur[kBG][3];
#pragma unroll
    for (int bb = 0; bb < kBG; ++bb) {
        h_own[bb] = 0.f;
#pragma unroll
        for (int g = 0; g < 3; ++g)
            gcur[bb][g] = b0 + bb < B ? __ldg(gi + ((size_t)dir * BT + (size_t)(b0 + bb) * T + t) * 3 * H + g * H + u) : 0.f;
    }

    for (int s = 0; s < T; ++s, t += dstep) {
        const int cur = s & 1, nxt = cur ^ 1;
        float gnext[kBG][3];
#pragma unroll
        for (int bb = 0; bb < kBG; ++bb)
#pragma unroll
            for (int g = 0; g < 3; ++g)
                gnext[bb][g] = (s + 1 < T && b0 + bb < B)
                    ? __ldg(gi + ((size_t)dir * BT + (size_t)(b0 + bb) * T + (t + dstep)) * 3 * H + g * H + u) : 0.f;

        float acc[kBG][3];
#pragma unroll
        for (int bb = 0; bb < kBG; ++bb) { acc[bb][0] = 0.f; acc[bb][1] = 0.f; acc[bb][2] = 0.f; }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
#pragma unroll
            for (int bb = 0; bb < kBG; ++bb) {
                const float4 h4 = *reinterpret_cast<const float4*>(&hs[cur][bb][32 * j + 4 * kl]);
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    acc[bb][g] = fmaf(w[g][j][0], h4.x, acc[bb][g]);
                    acc[bb][g] = fmaf(w[g][j][1], h4.y, acc[bb][g]);
                    acc[bb][g] = fmaf(w[g][j][2], h4.z, acc[bb][g]);
                    acc[bb][g] = fmaf(w[g][j][3], h4.w, acc[bb][g]);
                }
            }
        }
#pragma unroll
        for (int bb = 0; bb < kBG; ++bb) {
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                float v = acc[bb][g];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                acc[bb][g] = v + bh[g];
            }
            const float r = sigmoid_c(gcur[bb][0] + acc[bb][0]);
            const float z = sigmoid_c(gcur[bb][1] + acc[bb][1]);
            const float n = tanh_c(fmaf(r, acc[bb][2], gcur[bb][2]));
            const float hn = fmaf(z, h_own[bb] - n, n);          // (1 - z) n + z h
            h_own[bb] = hn;
            if (remote) remote[(nxt * kBG + bb) * H + u] = hn;
            if (kl == 0 && b0 + bb < B) out[((size_t)(b0 + bb) * T + t) * 2 * H + dir * H + u] = hn;
#pragma unroll
            for (int g = 0; g < 3; ++g) gcur[bb][g] = gnext[bb][g];
        }
        cluster.sync();                  // the new h is complete in every CTA; the old buffer may be overwritten
    }
}

template <int H>
int launch_cluster_fwd(const float* gi, const float* w_hh0, const float* w_hh1, const float* b_hh0, const float* b_hh1,
                       float* out, int B, int T, cudaStream_t s) {
    constexpr int CL = H / kUnits;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL, (B + kBG - 1) / kBG, 2);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DCASE_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gru_cluster_fwd_kernel<H>, gi, w_hh0, w_hh1, b_hh0, b_hh1, out, B, T));
    ++g_dcase_launches;
    return DCASE_OK;
}

// element offsets of nn.GRU(n_in, H, num_layers=2, bidirectional=True).named_parameters() order
struct GruOff { long long w_ih[2][2], w_hh[2][2], b_ih[2][2], b_hh[2][2], total; };
GruOff gru_offsets(int n_in, int H) {
    GruOff o{};
    long long off = 0;
    for (int l = 0; l < 2; ++l) {
        const int nin = l == 0 ? n_in : 2 * H;
        for (int d = 0; d < 2; ++d) {
            o.w_ih[l][d] = off; off += 3LL * H * nin;
            o.w_hh[l][d] = off; off += 3LL * H * H;
            o.b_ih[l][d] = off; off += 3LL * H;
            o.b_hh[l][d] = off; off += 3LL * H;
        }
    }
    o.total = off;
    return o;
}

}  // namespace

extern "C" {

size_t dcase_bigru_param_count_h(int n_in, int H) { return (size_t)gru_offsets(n_in, H).total; }

size_t dcase_bigru_workspace_bytes_h(int B, int To, int H) {
    if (B < 1 || To < 1 || H < 1) return 0;
    const size_t BT = (size_t)B * To;
    return (BT * 2 * 3 * H + BT * 2 * H) * sizeof(float);
}

int dcase_bigru_forward_h(dcase_ctx* ctx, const float* x, int B, int To, int n_in, int H, const float* rnn_params,
                          float* out, void* ws, void* stream_) {
    cudaStream_t s = (cudaStream_t)stream_;
    DCASE_REQUIRE(ctx && x && rnn_params && out && ws, "null argument");
    DCASE_REQUIRE(B >= 1 && To >= 1 && n_in >= 1, "bad shape");
    DCASE_REQUIRE(H == 128 || H == 256, "the cluster BiGRU is built for hidden 128 and 256 (64: dcase_bigru_forward)");
    const GruOff o = gru_offsets(n_in, H);
    const int BT = B * To;
    float* gi = (float*)ws;
    float* mid = gi + (size_t)BT * 2 * 3 * H;
    const float* rin = x;
    for (int l = 0; l < 2; ++l) {
        const int nin = l == 0 ? n_in : 2 * H;
        float* rout = l == 0 ? mid : out;
        GemmBatch gb{};
        for (int d = 0; d < 2; ++d)
            gb.p[d] = GemmProblem{BT, 3 * H, nin, rin, nin, 1, rnn_params + o.w_ih[l][d], 1, nin,
                                  gi + (size_t)d * BT * 3 * H, 3 * H, rnn_params + o.b_ih[l][d]};
        gb.n = 2; gb.split = 1; gb.mode = 0;
        int rc = launch_sgemm_batch(gb, s);
        if (rc != DCASE_OK) return rc;
        DCASE_PROF("gru_cluster_fwd", s);
        if (H == 128)
            rc = launch_cluster_fwd<128>(gi, rnn_params + o.w_hh[l][0], rnn_params + o.w_hh[l][1],
                                         rnn_params + o.b_hh[l][0], rnn_params + o.b_hh[l][1], rout, B, To, s);
        else
            rc = launch_cluster_fwd<256>(gi, rnn_params + o.w_hh[l][0], rnn_params + o.w_hh[l][1],
                                         rnn_params + o.b_hh[l][0], rnn_params + o.b_hh[l][1], rout, B, To, s);
        if (rc != DCASE_OK) return rc;
        rin = rout;
    }
    return DCASE_OK;
}

}  // extern "C"
