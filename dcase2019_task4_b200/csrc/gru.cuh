// BiGRU + small GEMM helpers (host launchers) -- see gru.cu.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"

struct GruFwdArgs {
    const float* gi;         // [2][B*T][3H] input projections incl. b_ih
    const float* w_hh[2];    // [3H][H]
    const float* b_hh[2];    // [3H]
    float* out;              // [B*T][2H]
    float* save_r;           // [2][B*T][H] each, nullable (no backward)
    float* save_z;
    float* save_n;
    float* save_hn;          // W_hn h + b_hn
    float* save_hp;          // h_{prev}
    int B, T;
};

struct GruBwdArgs {
    const float* d_out;      // [B*T][2H]
    const float* w_hh[2];
    const float* save_r;
    const float* save_z;
    const float* save_n;
    const float* save_hn;
    const float* save_hp;
    float* dgi;              // [2][B*T][3H]
    float* dgh;              // [2][B*T][3H]
    int B, T;
};

struct GemmProblem {
    int M, N, K;
    const float* A; long long sam, sak;     // A(m,k) = A[m*sam + k*sak]
    const float* B; long long sbk, sbn;     // B(k,n) = B[k*sbk + n*sbn]
    float* C; int ldc;
    const float* bias;                       // nullable, added once
};
struct GemmBatch {
    GemmProblem p[8];
    int n;          // problems in this launch
    int split;      // default K slices per problem
    int mode;       // 0: C = result (split must be 1); 1: atomicAdd into C
    int psplit[8];  // per-problem K slices (0 = use `split`); problems with a short K need none
    int zoff[9];    // filled by launch_sgemm_batch: first blockIdx.z of every problem
};
struct ColsumBatch {
    const float* A[4];   // [M][N] each
    float* out[4];       // [N] each, accumulated with atomics
    int n, M, N;
};

int launch_sgemm_batch(const GemmBatch& g, cudaStream_t s);
int launch_colsum_batch(const ColsumBatch& c, cudaStream_t s);
int gru_kernels_init();
int launch_gru_fwd(const GruFwdArgs& a, cudaStream_t s);
int launch_gru_bwd(const GruBwdArgs& a, cudaStream_t s);
