// Internal definition of the opaque dcase_ctx handle.
#pragma once
#include <cuda_runtime.h>

struct dcase_syncbn;
// p2p.cu: sum `vals` over the ranks of the group in place (one single-CTA kernel, rank-ordered, replay-safe)
int syncbn_world(const dcase_syncbn* h);
int syncbn_allreduce_f64(const dcase_syncbn* h, double* vals, int n, int slot, cudaStream_t s);
int syncbn_allreduce_f32(const dcase_syncbn* h, float* vals, int n, int slot, cudaStream_t s);

struct dcase_ctx {
    int device;
    int num_sms;
    // log-mel constant tables (device)
    float2* d_window2;   // [1024] {w[2n], w[2n+1]} of the symmetric Hamming window
    float* d_mel_wt;     // [4][22][32] Slaney weights of the 128 work items, transposed, zero padded, halved
    int* d_mel_start;    // [128] first FFT bin of each work item
    int* d_mel_owner;    // [64][2] {first item, item count} of each band
    int mel_nnz;
    float* d_resample_win;    // [2][32769] kaiser_best half window | its forward differences (read_audio's resampler)
    // second stream + events: the teacher forward runs concurrently with the student forward (dcase_mt_fwd_bwd)
    cudaStream_t aux_stream;
    cudaEvent_t ev_fork, ev_join;
    // per model (0 student, 1 teacher): the conv weight images only depend on the parameters, so they are built on a side
    // stream beside block 0 instead of sitting between cnn0 and conv1 on the forward chain
    cudaStream_t prep_stream[2];
    cudaEvent_t ev_prep_fork[2], ev_prep_join[2];
    // backward: the GRU and conv weight-gradient kernels leave the critical path (dcase_crnn_backward)
    cudaEvent_t ev_bwd_fork[4], ev_bwd_join[4];   // [0,1] GRU layers, [2,3] conv blocks 1, 2
    // scratch of the loss kernel: per-CTA partial sums + completion ticket (head_loss.cuh)
    float* d_loss_scratch;
    // exact-global-batch BatchNorm under data parallelism (dcase_ctx_set_syncbn); NULL = per-replica statistics
    dcase_syncbn* syncbn;
    // host copy of the dense filterbank for dcase_mel_filterbank()
    float* h_mel_dense;  // [64 * 1025]
};
