// tcgen05 GEMM for the GRU input projections and their data gradients (host launcher) -- see gemm_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"

// `problems` independent outputs C[i] (blockIdx.z), each the sum over `parts` operand pairs; pair (i, j) is
// A[i * parts + j] ([M][K] row-major, row stride lda) times B[i * parts + j]:
//   b_mn_major = 0:  B = [N][K] row-major (row stride ldb):  C = A B^T      (x @ W_ih^T)
//   b_mn_major = 1:  B = [K][N] row-major (row stride ldb):  C = A B        (dGi @ W_ih)
// problems * parts <= 2; N % 64 == 0, K % 64 == 0; bias[i] (length N) may be null.
struct GemmTcBatch {
    int problems, parts, M, N, K;
    const float* A[2]; long long lda;
    const float* B[2]; long long ldb; int b_mn_major;
    float* C[2]; int ldc;
    const float* bias[2];
};

int gemm_tc_init();
int launch_gemm_tc(const GemmTcBatch& p, cudaStream_t s);
