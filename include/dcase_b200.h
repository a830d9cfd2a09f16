/* dcase_b200 -- C ABI of the B200-native DCASE2019-task4 hot path.
 *
 * One shared library (libdcase_b200.so, sm_100a) behind the reference's Python surface.  The reference
 * (turpaultn/DCASE2019_task4) is pure Python and has no FFI layer of its own; each entry point below names
 * the reference interface (file:line under /root/reference) whose work it replaces.  INTEGRATION.md shows
 * the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns 0 or a negative DCASE_ERR_* code; dcase_last_error() gives the text
 *     (thread-local);
 *   - all tensor pointers are DEVICE pointers owned by the caller (torch tensors: tensor.data_ptr()),
 *     including workspaces (sizes from the *_bytes / *_count helpers); the library never allocates
 *     per call, never synchronises, and launches on the given stream (cudaStream_t passed as void*);
 *   - activations are channels-last fp32; features are [B, T, 64] (the reference's [B,1,T,64] NCHW
 *     with C = 1 is the same memory);
 *   - parameters are ONE flat fp32 slab in the order of CRNN(**cfg.crnn_kwargs).named_parameters()
 *     (SURVEY.md section 10): cnn.cnn.{conv,batchnorm,glu}{0,1,2}, rnn.rnn.*_l{0,1}[_reverse], dense,
 *     dense_softmax.  Gradients use the same layout.  BN running stats are [3][2][64] =
 *     {layer}{running_mean, running_var}{channel}.
 *
 * RNG contract (restated in numpy by oracle/philox.py)
 *   Philox4x32-10, key = (seed_lo, seed_hi), counter = (row_lo, row_hi, stream, step).
 *   Dropout(0.5) keep bit of element (row, col) = bit (col & 31) of output word (col >> 5), where row is
 *   the channels-last pixel index ((b*T + t)*F + f) for CNN block i (stream = 8*model_id + i) and the
 *   frame index (b*To + t) for the head (stream = 8*model_id + 3).  Teacher noise: counter row =
 *   (b*T + t)*16 + q for mel bins 4q..4q+3, stream 4; u = (w + 0.5) * 2^-32; Box-Muller pairs (u0,u1),
 *   (u2,u3); noise = 0.25 * |n|  (AugmentGaussianNoise, DataLoad.py:285: std = 0.5 ** 2).
 */
#ifndef DCASE_B200_H_
#define DCASE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCASE_B200_VERSION 100

#define DCASE_FLAG_BN_BATCH_STATS 1   /* module.train(): BatchNorm uses + updates batch statistics */
#define DCASE_FLAG_DROPOUT 2          /* Dropout(0.5) active (CNN blocks and head) */

typedef struct dcase_ctx dcase_ctx;

int dcase_version(void);
const char* dcase_last_error(void);
int dcase_ctx_create(dcase_ctx** out, int device);
int dcase_ctx_destroy(dcase_ctx* ctx);

/* ---- measurement helpers (bench.py) -------------------------------------------------------------- */
/* Kernel launches issued by this library since load (bench.py's "gpu_launches"). */
unsigned long long dcase_launch_count(void);
/* Per-kernel CUDA-event timing on the launching stream: begin, run some steps, end -> "name,count,total_ms\n"
 * lines into buf (synchronises the device). */
int dcase_profile_begin(void);
int dcase_profile_end(char* buf, size_t cap);
/* timeline of one EAGER step with the streams overlapping as in production: "name,stream,start_us,end_us" per launch */
int dcase_profile_timeline_begin(void);
int dcase_profile_timeline_end(char* buf, size_t cap);

/* Self-test of the tcgen05 / TMEM primitives (csrc/tc.cuh).  mode 0: D[128][64] = A[128][64] * B[64][64]^T with
 * K-major operands; mode 1: raw TMEM dump [128 lanes][64 columns] of D[m][n] = sum_p A[p][m] * B[p][n], p < 128,
 * with MN-major operands and M = 64. */
int dcase_selftest_umma(dcase_ctx* ctx, int mode, const float* A, const float* B, float* D, void* stream);
/* D[128][64] = Arows * B^T where the K-major A operand starts `shift` rows into a 256-row swizzled buffer with its
 * 8-row groups `pitch` rows apart (how the conv kernels address the 9 taps of one staged halo). base_mode 0 is the
 * encoding the kernels use (descriptor base_offset = 0: the swizzle follows absolute shared-memory addresses). */
int dcase_selftest_umma_shift(dcase_ctx* ctx, int shift, int pitch, int base_mode, const float* A, const float* B,
                              float* D, void* stream);
/* Micro-benchmark of the tensor core issue path: n_ctas CTAs (one warp each) issue `reps` back-to-back
 * tcgen05.mma kind::tf32 of shape M x N x 8 with both operands in shared memory; cycles_per_mma[n_ctas] (device)
 * receives the measured SM cycles per MMA (a_sbo / a_shift reproduce the strided, shifted halo operand of the convs).  tools/umma_bench.py prints the table DESIGN.md quotes. */
int dcase_bench_umma(dcase_ctx* ctx, int M, int N, int a_mn_major, int b_mn_major, int reps, int n_ctas,
                     int a_sbo_bytes /* K-major A: byte pitch of the 8-row groups (1024 = dense) */,
                     int a_shift_bytes /* K-major A: start offset into the buffer (multiple of 128) */,
                     int commit_every /* > 0: a tcgen05.commit after every commit_every MMAs (multiple of 4) */,
                     float* cycles_per_mma, void* stream);

/* Per-step scalars in device memory (so a captured CUDA graph replays with new values).
 * Layout must match DcaseStepScalars in csrc/common.cuh. */
typedef struct dcase_step_scalars {
    uint64_t seed;
    uint32_t step;
    float cons_weight;
    float ema_alpha;
    float lr;
    float bias_corr1;
    float bias_corr2;
    float grad_scale;
    float pad_;
} dcase_step_scalars;

/* ---- log-mel -------------------------------------------------------------------------------------- */

/* librosa.stft(center=True) frame count: 1 + n_samples / 511.  DatasetDcase2019Task4.py:211-218 */
int dcase_logmel_num_frames(int n_samples);

/* The Slaney filterbank the kernel uses (librosa.filters.mel(44100, 2048, 64, 0, 22050, htk=False,
 * norm=None)), dense float32 [64][1025] into HOST memory.  DatasetDcase2019Task4.py:220-225 */
int dcase_mel_filterbank(dcase_ctx* ctx, float* out_host);

/* DatasetDcase2019Task4.calculate_mel_spec (DatasetDcase2019Task4.py:197-231, save_log_feature=False):
 * wave [B][L] -> amplitude mel [B][T][64], T = 1 + L/511.  Hamming(2048) STFT, hop 511, reflect pad. */
int dcase_logmel_fwd(dcase_ctx* ctx, const float* wave, int B, int L, float* mel_amp, void* stream);
/* Same from 16-bit PCM as stored in the wav files (soundfile scaling 1/32768, utils/utils.py:187). */
int dcase_logmel_fwd_pcm16(dcase_ctx* ctx, const int16_t* wave, int B, int L, float* mel_amp, void* stream);

/* read_audio's mono mix-down (utils/utils.py:187-189: np.mean(audio, axis=1)): interleaved frames
 * [n_frames][n_channels] of float32 or 16-bit PCM (scaled by 1 / 32768 like soundfile.read) -> mono float32
 * [n_frames], ready for dcase_logmel_fwd. */
int dcase_audio_mixdown(dcase_ctx* ctx, const void* interleaved, int is_pcm16, long long n_frames, int n_channels,
                        float* mono, void* stream);

/* read_audio's resampling step (utils/utils.py:190-192: librosa.resample(audio, orig_sr=fs, target_sr=target_fs), default
 * res_type 'kaiser_best' of librosa < 0.10 = resampy's band-limited sinc interpolation: 64 zero crossings, 512 table
 * samples per crossing, Kaiser beta 14.769656459379492, roll-off 0.9475937167399596; restated in oracle/resample.py,
 * parity unpinned against resampy / librosa, which are not installed).  mono float32 [n_in] at sr_in -> [n_out] at
 * sr_out with n_out = dcase_audio_resample_len(n_in, sr_in, sr_out) = ceil(n_in * sr_out / sr_in) (librosa's fix_length:
 * the samples past resampy's int(n_in * ratio) are zero). */
long long dcase_audio_resample_len(long long n_in, int sr_in, int sr_out);
int dcase_audio_resample(dcase_ctx* ctx, const float* mono, long long n_in, int sr_in, int sr_out, float* out,
                         void* stream);
/* test hook, host only: resampy's float64 clock (the running sum of 1 / ratio) as the kernel reconstructs it from its
 * per-binade segment table -- must equal the serial sum bit for bit */
int dcase_audio_resample_clock(long long n, int sr_in, int sr_out, double* out_host);

/* get_transforms(frames, scaler, augment_type="noise") (utils/utils.py:397-412) on a batch:
 * AugmentGaussianNoise (DataLoad.py:274-287) -> ApplyLog / librosa.amplitude_to_db (DataLoad.py:192-207)
 * -> PadOrTrunc (DataLoad.py:210-259) -> ToTensor -> Normalize / Scaler.normalize (Scaler.py:99-105).
 * mel_amp [B][T_in][64] -> clean [B][T_out][64] and, if noisy != NULL, noisy [B][T_out][64].
 * noise: explicit |N(0,0.25)| sample [B][T_in][64] or NULL (Philox, see RNG contract).  With scalars non-NULL the
 * seed comes from the device struct and `step` is an OFFSET added to its step (0 = this iteration's noise, 1 = the
 * noise of the next iteration, for features prepared one step ahead).  clip_max_ws: [2*B] floats of scratch. */
int dcase_logmel_finish(dcase_ctx* ctx, const float* mel_amp, int B, int T_in, int T_out, const float* mean,
                        const float* stdv, const float* noise, uint64_t seed, uint32_t step, const void* scalars,
                        float* clip_max_ws, float* clean, float* noisy, void* stream);

/* Scaler.means (utils/Scaler.py:34-87) over a batch of clips, as main.py:249-250 runs it over the training set
 * through get_transforms(frames) without a scaler: sums [2][64] float64 (zeroed by the caller before the first
 * batch) accumulates, per mel bin, the per-clip mean over the T_out frames of the feature and of its float32
 * square.  apply_log = 1: feats is the amplitude mel [B][T_in][64] and the kernel applies ApplyLog (dB, top_db
 * floor at the clip maximum) and PadOrTrunc(T_out) itself; clip_max_ws is [B] floats of scratch.
 * apply_log = 0: feats [B][T_in == T_out][64] are reduced as they are (clip_max_ws may be NULL). */
int dcase_scaler_accumulate(dcase_ctx* ctx, const float* feats, int B, int T_in, int T_out, int apply_log,
                            float* clip_max_ws, double* sums, void* stream);
/* Scaler.means' division by the sample count (Scaler.py:72-73) and Scaler.calculate_scaler's std
 * (Scaler.py:89-97): mean / mean_of_square [64] float64 (the state_dict wire format, Scaler.py:107-113) and the
 * float32 mean / std [64] that dcase_logmel_finish reads.  Any output may be NULL. */
int dcase_scaler_finalize(dcase_ctx* ctx, const double* sums, long long n_samples, double* mean,
                          double* mean_of_square, float* mean_f32, float* std_f32, void* stream);

/* ---- CRNN ------------------------------------------------------------------------------------------ */

/* Number of fp32 elements of the flat parameter slab (214,356 for n_class = 10). */
size_t dcase_crnn_param_count(int n_class);
/* Offset (in elements) of a named parameter inside the slab, or -1.  Names as in named_parameters(). */
long long dcase_crnn_param_offset(int n_class, const char* name);

size_t dcase_crnn_workspace_bytes(int B, int T, int n_class);
/* Location of a named intermediate inside the workspace (tests / debugging); returns 0 or DCASE_ERR_ARG.
 * Names: out0 ypre1 out1 ypre2 out2 rnn0 rnn1 fold0 bn1 bn2 d_out0 d_out1 d_out2 d_rnn0 d_rnn1 */
int dcase_crnn_ws_tensor(int B, int T, int n_class, const char* name, size_t* offset_bytes, size_t* n_elems);

/* CRNN.forward (models/CRNN.py:59-84; CNN.py:85-89; RNN.py:14-16).
 * x [B][T][64], T % 8 == 0 -> strong [B][T/8][n_class], weak [B][n_class].
 * flags: DCASE_FLAG_*; without BN_BATCH_STATS this is module.eval().  bn_running [3][2][64] is updated
 * in train mode.  model_id selects the dropout streams (0 student, 1 teacher). */
int dcase_crnn_forward(dcase_ctx* ctx, const float* x, int B, int T, int n_class, const float* params,
                       float* bn_running, int flags, uint64_t seed, uint32_t step, int model_id,
                       const void* scalars, float* strong, float* weak, void* workspace, void* stream);

/* BidirectionalGRU.forward (models/RNN.py:7-16) on its own: nn.GRU(64, 64, num_layers=2, bidirectional=True,
 * batch_first=True), gate order (r, z, n), zero initial state.  x [B][To][64] -> out [B][To][128] (forward | reverse).
 * rnn_params points at rnn.rnn.weight_ih_l0 inside the flat slab (or at a copy of the 16 GRU tensors in
 * named_parameters() order: weight_ih, weight_hh, bias_ih, bias_hh for l0, l0_reverse, l1, l1_reverse = 124,416
 * floats).  To <= 136 (the sequence is resident in shared memory).  BASELINE.json configs[4] times this entry. */
size_t dcase_bigru_workspace_bytes(int B, int To);
int dcase_bigru_forward(dcase_ctx* ctx, const float* x, int B, int To, const float* rnn_params, float* out,
                        void* workspace, void* stream);

/* EXPERIMENTAL (compiles, not yet run on hardware): the same operator for hidden sizes 128 / 256 (the other
 * points of BASELINE.json configs[4]), nn.GRU(n_in, H, num_layers=2, bidirectional=True, batch_first=True).
 * x [B][To][n_in] -> out [B][To][2H]; rnn_params = the 16 tensors in named_parameters() order
 * (dcase_bigru_param_count_h floats).  The recurrence of one (direction, 4 clips) runs on a thread-block cluster
 * of H / 32 CTAs that exchange h through distributed shared memory (csrc/gru_cluster.cu); no length limit. */
size_t dcase_bigru_param_count_h(int n_in, int H);
size_t dcase_bigru_workspace_bytes_h(int B, int To, int H);
int dcase_bigru_forward_h(dcase_ctx* ctx, const float* x, int B, int To, int n_in, int H, const float* rnn_params,
                          float* out, void* workspace, void* stream);

/* Backward of the forward pass that filled `workspace` (same x, params, flags, seed, step, model_id).
 * grads (param_count elements) is overwritten with d loss / d params  (loss.backward(), main.py:153). */
int dcase_crnn_backward(dcase_ctx* ctx, const float* x, int B, int T, int n_class, const float* params, int flags,
                        uint64_t seed, uint32_t step, int model_id, const void* scalars, const float* d_strong,
                        const float* d_weak, const float* weak, void* workspace, float* grads, void* stream);

/* ---- mean-teacher losses (main.py:95-145; main_simple_CRNN.py:46-66 when strong_t == NULL) ---------- */
/* meters[8] = {weak_class_loss, Weak EMA loss, Strong loss, Strong EMA loss, Consistency strong,
 *              Consistency weak, Loss, Consistency weight};  d_strong / d_weak = d Loss / d student outputs.
 * weak_mask = slice(weak_lo, weak_hi), strong_mask = slice(strong_lo, strong_hi); empty slice = None. */
int dcase_mt_loss(dcase_ctx* ctx, const float* strong_s, const float* weak_s, const float* strong_t,
                  const float* weak_t, const float* target, int B, int To, int n_class, int weak_lo, int weak_hi,
                  int strong_lo, int strong_hi, float cons_weight, const void* scalars, float* meters,
                  float* d_strong, float* d_weak, void* stream);

/* ---- optimizer.step() + update_ema_variables (main.py:154-157, :45-49) ------------------------------ */
/* torch.optim.Adam (no amsgrad / weight decay) on flat slabs, then p_ema = alpha*p_ema + (1-alpha)*p.
 * step_t is Adam's step count AFTER the increment; g is multiplied by grad_scale first (1/world_size). */
int dcase_adam_ema_step(dcase_ctx* ctx, float* p, const float* g, float* m, float* v, float* p_ema, size_t n,
                        float lr, float beta1, float beta2, float eps, int step_t, float ema_alpha,
                        float grad_scale, const void* scalars, void* stream);

/* ---- data-parallel gradient exchange fused with the optimizer (verified on 2 / 4 / 8 GPUs, tests/test_gpu_dp.py) ---- */
/* Under data parallelism (one process per GPU of one node) the pair "all-reduce of the gradient slab" +
 * dcase_adam_ema_step becomes ONE kernel that reads every rank's slab out of the peers' HBM over NVLink (CUDA IPC
 * mappings), sums in rank order, and applies Adam + EMA (main.py:152-157, :45-49 on N replicas).  csrc/p2p.cu describes
 * the flag protocol.  Set-up: every rank calls dcase_p2p_create (allocates its slab + flag block, returns
 * dcase_p2p_handle_bytes() bytes of IPC handles into HOST memory), the ranks exchange the handle blobs (any host
 * channel, e.g. torch.distributed.all_gather_object), then dcase_p2p_connect(world x blob, rank-major).  Per step:
 * dcase_p2p_begin_step before the backward writes dcase_p2p_grads(), dcase_p2p_adam_ema_step after it. */
typedef struct dcase_p2p dcase_p2p;
int dcase_p2p_handle_bytes(void);
int dcase_p2p_create(dcase_ctx* ctx, int world, int rank, size_t n_floats, dcase_p2p** out, void* handles_out_host);
int dcase_p2p_connect(dcase_p2p* h, const void* all_handles_host);
float* dcase_p2p_grads(dcase_p2p* h);
int dcase_p2p_begin_step(dcase_p2p* h, void* stream);
int dcase_p2p_adam_ema_step(dcase_ctx* ctx, dcase_p2p* h, float* p, float* m, float* v, float* p_ema, float lr,
                            float beta1, float beta2, float eps, int step_t, float ema_alpha, const void* scalars,
                            void* stream);
int dcase_p2p_destroy(dcase_p2p* h);

/* ---- SyncBN: exact-global-batch BatchNorm statistics under data parallelism ---------------------- */
/* The reference run at a global batch of N x 24 on one device normalises each BatchNorm2d over ALL clips
 * (models/CNN.py:49, train mode).  With a group attached to the context, every train-mode dcase_crnn_forward /
 * dcase_crnn_backward / dcase_mt_fwd_bwd sums the per-channel statistics (forward: sum x, sum x^2 -- block 0: the tap
 * moments of its input; backward: sum dy, sum dy xhat -- block 0: its {U | S2} accumulator) over the ranks with one
 * single-CTA kernel per BatchNorm (peer-memory mailboxes + epoch flags, csrc/p2p.cu; no NCCL, replays inside a CUDA
 * graph), so N replicas reproduce the one-device result.  BatchNorm-derived parameter gradients are written as
 * global / N: the gradient exchange (sum, then x 1/N) restores them.  Every rank must issue the same sequence of
 * train-mode calls; the teacher forward runs on the student's stream in this mode.  Default (no group): per-replica
 * statistics, the reference's semantics at its own batch of 24 per device.
 * Set-up mirrors dcase_p2p_*: create (returns dcase_syncbn_handle_bytes() bytes of IPC handle into HOST memory),
 * exchange the blobs, connect(world x blob, rank-major), dcase_ctx_set_syncbn(ctx, h) (NULL detaches). */
typedef struct dcase_syncbn dcase_syncbn;
int dcase_syncbn_handle_bytes(void);
int dcase_syncbn_create(dcase_ctx* ctx, int world, int rank, dcase_syncbn** out, void* handle_out_host);
int dcase_syncbn_connect(dcase_syncbn* h, const void* all_handles_host);
int dcase_ctx_set_syncbn(dcase_ctx* ctx, dcase_syncbn* h);
/* the collective itself (sum in place over the group, n * elem <= 8192 bytes, slot in [0,16): slots 0-8 are used by
 * the CRNN path): exposed for tests */
int dcase_syncbn_allreduce(dcase_syncbn* h, void* vals_dev, int n, int is_double, int slot, void* stream);
int dcase_syncbn_destroy(dcase_syncbn* h);

/* BatchNorm batch statistics of CNN block 0 (models/CNN.py:49) follow from 54 moments of the block's INPUT (conv0 is
 * linear in its 9 taps): 9 tap sums + 45 second moments, float64, mom[56] (the last two entries are scratch).  They
 * depend on x alone, so a pipelined caller computes them beside the previous iteration and hands them to
 * dcase_mt_fwd_bwd (mom_s / mom_t); otherwise the forward computes them itself. */
int dcase_cnn0_input_moments(dcase_ctx* ctx, const float* x, int B, int T, double* mom, void* stream);

/* ---- one mean-teacher iteration, main.py:84-153 (forward x2, losses, backward) ----------------------- */
typedef struct dcase_mt_args {
    const float* x_student;   /* [B][T][64] clean */
    const float* x_teacher;   /* [B][T][64] noisy; NULL = no teacher (main_simple_CRNN.py) */
    const float* target;      /* [B][T/8][n_class], -1 rows for unlabeled clips */
    int B, T, n_class;
    int weak_lo, weak_hi, strong_lo, strong_hi;
    const float* params_s;
    const float* params_t;
    float* bn_s;
    float* bn_t;
    int flags;
    uint64_t seed;
    uint32_t step;
    float cons_weight;
    const void* scalars;      /* device dcase_step_scalars or NULL */
    float* strong_s;          /* outputs */
    float* weak_s;
    float* strong_t;
    float* weak_t;
    float* meters;            /* [8] */
    float* d_strong;          /* scratch [B][T/8][n_class] */
    float* d_weak;            /* scratch [B][n_class] */
    void* ws_s;
    void* ws_t;
    float* grads;             /* [param_count] out */
    void* after_forward_event; /* optional cudaEvent_t recorded on `stream` once the student forward is enqueued (NULL: none):
                                 lets the caller start independent work -- the next batch's features -- on another stream
                                 alongside the backward instead of alongside the forward */
    const double* mom_s;      /* optional (NULL: computed inside): the tap moments of x_student / x_teacher from */
    const double* mom_t;      /* dcase_cnn0_input_moments, e.g. computed beside the previous iteration             */
} dcase_mt_args;

int dcase_mt_fwd_bwd(dcase_ctx* ctx, const dcase_mt_args* args, void* stream);

/* sizeof(dcase_mt_args) / sizeof(dcase_step_scalars) as compiled, so FFI struct mirrors can be checked at load. */
size_t dcase_sizeof_mt_args(void);
size_t dcase_sizeof_step_scalars(void);

#ifdef __cplusplus
}
#endif
#endif /* DCASE_B200_H_ */
