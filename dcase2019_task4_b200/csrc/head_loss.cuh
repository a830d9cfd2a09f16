// Head / loss / optimizer kernels (host launchers) -- see head_loss.cu.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"

struct LossArgs {
    const float* strong_s;
    const float* weak_s;
    const float* strong_t;   // nullable: no teacher (main_simple_CRNN.py)
    const float* weak_t;
    const float* target;     // [B][To][NC]
    int B, To, NC;
    int weak_lo, weak_hi;    // weak_mask = slice(weak_lo, weak_hi); empty = None
    int strong_lo, strong_hi;
    float cons_weight;
    const DcaseStepScalars* sc;
    float* meters;           // [8]
    float* partials;         // [kLossMaxCtas][8] per-CTA partial sums (context scratch)
    unsigned int* ticket;    // completion counter (context scratch, zero between launches)
    float* d_strong;
    float* d_weak;
};

struct HeadArgs {
    const float* x;          // [B*To][128] BiGRU output
    const float* w_dense;    // [NC][128]
    const float* b_dense;    // [NC]
    const float* w_soft;     // [NC][128]
    const float* b_soft;     // [NC]
    int B, To, NC;
    int drop;                // 1: Dropout(0.5) on x
    uint64_t seed;
    uint32_t step;
    uint32_t stream;
    const DcaseStepScalars* sc;
    float* strong;           // [B*To][NC]   (fwd out)
    float* weak;             // [B][NC]      (fwd out, bwd in)
    float* den;              // [B][NC]      sum_t sof (fwd out, bwd in)
    // backward
    const float* d_strong;   // [B*To][NC]
    const float* d_weak;     // [B][NC]
    float* d_x;              // [B*To][128]
    float* g_w_dense;
    float* g_b_dense;
    float* g_w_soft;
    float* g_b_soft;
    // backward, optional: the loss kernel's work as the first phase of head_bwd (grid = B <= kLossMaxCtas CTAs; CTA b then
    // consumes the gradients of clip b it has just written: d_strong / d_weak above must equal loss.d_strong / loss.d_weak)
    int fused_loss;
    LossArgs loss;
};

constexpr int kLossMaxCtas = 1024;
constexpr size_t kLossScratchBytes = (kLossMaxCtas * 8 + 8) * sizeof(float);

int head_kernels_init();
int launch_head_fwd(const HeadArgs& a, cudaStream_t s);
int launch_head_bwd(const HeadArgs& a, cudaStream_t s);
int launch_mt_loss(const LossArgs& a, cudaStream_t s);
int launch_adam_ema(float* p, const float* g, float* m, float* v, float* p_ema, long long n, float lr, float beta1,
                    float beta2, float eps, float bc1, float bc2, float alpha, float grad_scale,
                    const DcaseStepScalars* sc, int num_sms, cudaStream_t s);
