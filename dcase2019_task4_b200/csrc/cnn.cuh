// CNN stack kernels (host launchers) -- see cnn.cu for the reference citations.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"

// Folded layer-0 parameters produced by bn0_finalize (floats):
//   wf[9][64] (tap-major), bf[64], mean[64], invstd[64], a[64]
constexpr int kFold0Wf = 0;
constexpr int kFold0Bf = 576;
constexpr int kFold0Mean = 640;
constexpr int kFold0Invstd = 704;
constexpr int kFold0A = 768;
constexpr int kFold0Size = 832;

// BN affine of layers 1,2 produced by bn_finalize (floats): scale[64], shift[64], mean[64], invstd[64]
constexpr int kBnScale = 0;
constexpr int kBnShift = 64;
constexpr int kBnMean = 128;
constexpr int kBnInvstd = 192;
constexpr int kBnSize = 256;

struct DropoutCfg {
    int enabled;            // 0: identity; 1: p = 0.5 (keep-bit from Philox, scale 2)
    uint64_t seed;
    uint32_t step;
    uint32_t stream;        // model_id * 8 + layer
    const DcaseStepScalars* sc;  // if non-null overrides seed/step (CUDA-graph replay)
};

// training: tap moments of the input AND (last block, fold0 != NULL) the BN fold of block 0; eval: launch_bn0_finalize alone
int launch_cnn0_moments(const float* x, int B, int T, double* mom, const float* conv_w, const float* conv_b,
                        const float* gamma, const float* beta, float* running, float* fold0, int num_sms, cudaStream_t s);
int launch_bn0_finalize(const double* mom, long long n_pix, const float* conv_w, const float* conv_b,
                        const float* gamma, const float* beta, float* running /*[2][64]*/, int training,
                        float* fold0, double* mom_copy /*nullable: also copy the 54 moments there*/, cudaStream_t s);
// cnn0.cu: fused block 0 (conv0 + BN + GLU + dropout + pool), forward / backward / parameter gradients
// out (fp32, tf32-rounded: operand of the backward's weight-gradient MMAs) and out_h (fp16: operand of the next block's
// forward conv) are both optional
int launch_cnn0_fwd(const float* x, int B, int T, const float* fold0, const float* glu_w, const float* glu_b,
                    DropoutCfg drop, float* out, void* out_h, unsigned int* tile_ctr /*4 bytes of scratch*/, int num_sms,
                    cudaStream_t s);
constexpr int kCnn0AccFloats = 128 * 16;   // {U[64][16], S2[64][16]}, zeroed before launch_cnn0_bwd
// what turns block 0's {U | S2} accumulator into its parameter gradients (n_pix = pixels behind the statistics; the
// gradients are written x param_grad_scale, see launch_bn_bwd_apply)
struct Cnn0BwdFinalize {
    const double* mom;
    long long n_pix;
    const float *conv_w, *conv_b;
    float param_grad_scale;
    float *g_conv_w, *g_conv_b, *g_gamma, *g_beta, *g_glu_w, *g_glu_b;
};
// fin != NULL: the last CTA to finish runs the finalize pass itself (one launch less at the end of the backward chain)
int launch_cnn0_bwd(const float* x, int B, int T, const float* fold0, const float* glu_w, const float* glu_b,
                    DropoutCfg drop, const float* d_out, float* us, const Cnn0BwdFinalize* fin, int num_sms, cudaStream_t s);
int launch_cnn0_bwd_finalize(const Cnn0BwdFinalize& fin, const float* fold0, const float* glu_w, const float* us, cudaStream_t s);
int cnn0_kernels_init();
int launch_glu_pool_fwd(const float* ypre, long long n_pix, int F, const float* glu_img, DropoutCfg drop, float* out,
                        void* out_h /*nullable fp16 copy*/, int num_sms, cudaStream_t s);
// conv_tc.cu: weight images are the swizzled shared-memory layout of the tcgen05 B operand (36864 floats each)
// operand images (forward + mirrored / transposed for the data gradient) of BOTH 64 -> 64 conv layers, one launch
// (w_fwd: tf32 forward image, nullable; w_fwd_h: fp16 forward image [9][64][64] = 73,728 bytes, nullable)
int launch_conv_w_prep(const float* w1, float* w_fwd1, float* w_dgrad1, void* w_fwd1_h, const float* w2, float* w_fwd2,
                       float* w_dgrad2, void* w_fwd2_h, cudaStream_t s);
int launch_conv3x3_h(const void* in_h, int B, int T_l, int F, const void* w_img_h, const float* bias, float* out,
                     double* stats /*nullable [2][64]*/, int num_sms, cudaStream_t s);
int launch_conv3x3(const float* in, int B, int T_l, int F, const float* w_img, const float* bias,
                   float* out, double* stats /*nullable [2][64]*/, int num_sms, cudaStream_t s);
// GLU operand image written by bn_finalize (bytes): W' 16 KB | P 8 KB | {bias', exp scale, exp shift} 768 B |
// Wm 16 KB (the un-folded Wg as MN-major operand of the backward's dY = DL Wg)
constexpr int kGluImgP = 16384;
constexpr int kGluImgMisc = 16384 + 8192;
constexpr int kGluImgWm = 16384 + 8192 + 768;
constexpr int kGluImgBytes = kGluImgWm + 16384;
int launch_bn_finalize(const double* stats, long long n_pix, const float* gamma, const float* beta,
                       float* running, int training, float* bn, const float* glu_w, const float* glu_b, int F,
                       float* glu_img /*nullable, kGluImgBytes*/, cudaStream_t s);
int launch_glu_pool_bwd(const float* ypre, long long n_pix, int F, const float* bn, const float* glu_img, DropoutCfg drop,
                        const float* d_out, float* d_y, float* s12 /*[2][64]*/, float* g_glu_w, float* g_glu_b, int num_sms,
                        cudaStream_t s);
// s12 = {sum dy, sum dy xhat} over the n_stat pixels the statistics were taken over (n_pix local ones are rewritten);
// the BatchNorm parameter gradients
// are written as param_grad_scale x those sums (SyncBN: global sums, 1 / world_size so the gradient exchange restores them)
int launch_bn_bwd_apply(float* d_y, const float* ypre, long long n_pix, long long n_stat, const float* bn, const float* gamma,
                        const float* s12, float param_grad_scale, float* g_gamma, float* g_beta, float* g_conv_b,
                        int num_sms, cudaStream_t s);
int launch_conv_wgrad(const float* d_pre, const float* in, int B, int T_l, int F, float* g_w, int num_sms,
                      cudaStream_t s);
int cnn_kernels_init();
int glu_tma_kernels_init();
int conv_tc_kernels_init();
