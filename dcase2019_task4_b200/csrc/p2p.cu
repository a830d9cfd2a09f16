// Data-parallel exchanges over NVLink peer memory: (1) the gradient exchange FUSED with the optimizer, (2) the small
// all-reduce of BatchNorm statistics behind the exact-global-batch mode (SyncBN, second half of this file).
//
// Replaces, under data parallelism, the pair  "NCCL all-reduce of the flat gradient slab" + "fused Adam + EMA kernel"
// (loss.backward() / optimizer.step() / update_ema_variables of baseline/main.py:152-157, :45-49 on N replicas) by ONE
// kernel: every rank reads the gradient slab of EVERY rank straight out of the peers' HBM (CUDA IPC mappings of
// cudaMalloc'd slabs, loads travel over NVLink / NVSwitch), sums them in a fixed rank order (so all replicas compute
// bit-identical sums and stay in lock-step without a broadcast), and applies Adam + the teacher EMA to its own replica.
// The exchange moves (N - 1) x 857 KB into each GPU, a few microseconds at NVLink 5 rates; what it removes is the
// collective's launch / protocol latency (~20-30 us per step against a 1.2 ms step).
//
// Synchronisation uses flags in each rank's own memory that the PEERS write remotely (so every spin is a local poll):
//   ready[r] = e   written by rank r after its backward of step e        -> "my gradients of step e are complete"
//   done[r]  = e   written by rank r after its optimizer kernel of step e -> "I no longer read anybody's step-e slab"
// Per step and rank, in stream order:
//   p2p_wait_done    (1 warp)  spin until done[r] >= epoch for all r: nobody still reads my slab -> backward may overwrite it
//   ... forward / backward write the local slab ...
//   p2p_signal_ready (1 warp)  epoch += 1; system fence; ready[rank] = epoch in every peer's flag block
//   adam_ema_p2p               spin until ready[r] >= epoch for all r; volatile (uncached) loads of all N slabs; Adam; EMA;
//                              the last CTA to finish writes done[rank] = epoch in every peer's flag block
// The epoch counter lives in device memory and is advanced by the signal kernel itself, so a captured CUDA graph replays
// without host-side values.  A rank cannot run ahead by more than one step: its optimizer kernel of step e + 1 waits for
// every peer's ready flag of step e + 1.
//
// STATUS: verified on 2 / 4 / 8 B200s in round 2 (tests/test_gpu_dp.py); the default exchange of MeanTeacherEngine for
// world_size > 1 (DCASE_DP_NCCL=1 selects NCCL all-reduce + the plain optimizer kernel).
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/dcase_b200.h"
#include "common.cuh"
#include "ctx.h"

namespace {
constexpr int kMaxWorld = 16;
}

struct dcase_p2p {
    int world, rank;
    size_t n;                      // floats per slab
    float* grads;                  // local slab (cudaMalloc, exported over CUDA IPC)
    uint32_t* sync;                // local flag block: ready[world], done[world], epoch, ticket
    float** d_peer_grads;          // device array [world] (entry `rank` is the local slab)
    uint32_t** d_peer_sync;        // device array [world]
    void* opened[2 * kMaxWorld];   // IPC mappings to close
    int n_opened;
};

namespace {

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// peer slabs are re-written every step at the same addresses: the loads must not be served by a stale local cache line
__device__ __forceinline__ float4 ld_volatile_f4(const float* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool reached(uint32_t flag, uint32_t epoch) { return (int32_t)(flag - epoch) >= 0; }

__global__ void p2p_wait_done_kernel(const uint32_t* sync, int world) {
    const uint32_t epoch = ld_volatile_u32(sync + 2 * world);
    if ((int)threadIdx.x < world)
        while (!reached(ld_volatile_u32(sync + world + threadIdx.x), epoch)) __nanosleep(64);
}

__global__ void p2p_signal_ready_kernel(uint32_t* sync, uint32_t* const* peer_sync, int world, int rank) {
    __shared__ uint32_t epoch_s;
    if (threadIdx.x == 0) {
        epoch_s = ld_volatile_u32(sync + 2 * world) + 1u;
        st_volatile_u32(sync + 2 * world, epoch_s);
        st_volatile_u32(sync + 2 * world + 1, 0u);          // completion ticket of this step's optimizer kernel
    }
    __syncthreads();
    __threadfence_system();                                 // the slab written by the preceding kernels is visible system-wide
    if ((int)threadIdx.x < world) st_volatile_u32(peer_sync[threadIdx.x] + rank, epoch_s);
}

__global__ void __launch_bounds__(256)
adam_ema_p2p_kernel(float* __restrict__ p, float* const* __restrict__ peer_grads, uint32_t* sync,
                    uint32_t* const* __restrict__ peer_sync, int world, int rank, float* __restrict__ m,
                    float* __restrict__ v, float* __restrict__ p_ema, long long n, float lr, float beta1, float beta2,
                    float eps, float bc1, float bc2, float alpha, const DcaseStepScalars* __restrict__ sc) {
    __shared__ uint32_t epoch_s;
    __shared__ int last_s;
    if (sc) { lr = sc->lr; bc1 = sc->bias_corr1; bc2 = sc->bias_corr2; alpha = sc->ema_alpha; }
    if (threadIdx.x == 0) epoch_s = ld_volatile_u32(sync + 2 * world);
    __syncthreads();
    const uint32_t epoch = epoch_s;
    if ((int)threadIdx.x < world)
        while (!reached(ld_volatile_u32(sync + threadIdx.x), epoch)) __nanosleep(32);
    __syncthreads();
    __threadfence_system();

    const float grad_scale = 1.f / (float)world;            // mean over the replicas (main.py's loss is a batch mean)
    const float step_size = lr / bc1;
    const float inv_sqrt_bc2 = rsqrtf(bc2);
    const long long n4 = n >> 2;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < world; ++r) {                   // fixed order: every replica computes the same sum
            const float4 t = ld_volatile_f4(peer_grads[r] + 4 * i);
            g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
        }
        const float gv[4] = {g.x * grad_scale, g.y * grad_scale, g.z * grad_scale, g.w * grad_scale};
        float4 pm = reinterpret_cast<float4*>(m)[i], pv = reinterpret_cast<float4*>(v)[i];
        float4 pp = reinterpret_cast<float4*>(p)[i];
        float mm[4] = {pm.x, pm.y, pm.z, pm.w}, vv[4] = {pv.x, pv.y, pv.z, pv.w}, ww[4] = {pp.x, pp.y, pp.z, pp.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            mm[k] = beta1 * mm[k] + (1.f - beta1) * gv[k];
            vv[k] = beta2 * vv[k] + (1.f - beta2) * gv[k] * gv[k];
            const float denom = sqrtf(vv[k]) * inv_sqrt_bc2 + eps;
            ww[k] = ww[k] - step_size * (mm[k] / denom);
        }
        reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
        reinterpret_cast<float4*>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
        reinterpret_cast<float4*>(p)[i] = make_float4(ww[0], ww[1], ww[2], ww[3]);
        if (p_ema) {
            float4 pe = reinterpret_cast<float4*>(p_ema)[i];
            pe.x = alpha * pe.x + (1.f - alpha) * ww[0]; pe.y = alpha * pe.y + (1.f - alpha) * ww[1];
            pe.z = alpha * pe.z + (1.f - alpha) * ww[2]; pe.w = alpha * pe.w + (1.f - alpha) * ww[3];
            reinterpret_cast<float4*>(p_ema)[i] = pe;
        }
    }
    // the last CTA to finish tells every peer that this rank no longer reads their slabs of this step
    __syncthreads();
    if (threadIdx.x == 0) last_s = atomicAdd(sync + 2 * world + 1, 1u) == gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    if (last_s && (int)threadIdx.x < world) {
        __threadfence_system();
        st_volatile_u32(peer_sync[threadIdx.x] + world + rank, epoch);
    }
}

}  // namespace

extern "C" {

int dcase_p2p_handle_bytes(void) { return 2 * (int)sizeof(cudaIpcMemHandle_t); }

int dcase_p2p_create(dcase_ctx* ctx, int world, int rank, size_t n_floats, dcase_p2p** out, void* handles_out) {
    DCASE_REQUIRE(ctx && out && handles_out, "null argument");
    DCASE_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "bad world / rank");
    DCASE_REQUIRE(n_floats > 0 && n_floats % 4 == 0, "the slab is read as float4 (214,356 = 4 x 53,589)");
    dcase_p2p* h = new dcase_p2p();
    h->world = world; h->rank = rank; h->n = n_floats; h->n_opened = 0;
    h->grads = nullptr; h->sync = nullptr; h->d_peer_grads = nullptr; h->d_peer_sync = nullptr;
    DCASE_CUDA_CHECK(cudaMalloc(&h->grads, n_floats * sizeof(float)));
    DCASE_CUDA_CHECK(cudaMemset(h->grads, 0, n_floats * sizeof(float)));
    DCASE_CUDA_CHECK(cudaMalloc(&h->sync, (2 * world + 2) * sizeof(uint32_t)));
    DCASE_CUDA_CHECK(cudaMemset(h->sync, 0, (2 * world + 2) * sizeof(uint32_t)));
    DCASE_CUDA_CHECK(cudaMalloc(&h->d_peer_grads, world * sizeof(float*)));
    DCASE_CUDA_CHECK(cudaMalloc(&h->d_peer_sync, world * sizeof(uint32_t*)));
    cudaIpcMemHandle_t* hs = (cudaIpcMemHandle_t*)handles_out;
    DCASE_CUDA_CHECK(cudaIpcGetMemHandle(&hs[0], h->grads));
    DCASE_CUDA_CHECK(cudaIpcGetMemHandle(&hs[1], h->sync));
    DCASE_CUDA_CHECK(cudaDeviceSynchronize());
    *out = h;
    return DCASE_OK;
}

// all_handles: world x dcase_p2p_handle_bytes() bytes, rank-major, as gathered from every rank's dcase_p2p_create
int dcase_p2p_connect(dcase_p2p* h, const void* all_handles) {
    DCASE_REQUIRE(h && all_handles, "null argument");
    float* pg[kMaxWorld];
    uint32_t* ps[kMaxWorld];
    const cudaIpcMemHandle_t* hs = (const cudaIpcMemHandle_t*)all_handles;
    for (int r = 0; r < h->world; ++r) {
        if (r == h->rank) { pg[r] = h->grads; ps[r] = h->sync; continue; }
        void* a = nullptr;
        void* b = nullptr;
        DCASE_CUDA_CHECK(cudaIpcOpenMemHandle(&a, hs[2 * r], cudaIpcMemLazyEnablePeerAccess));
        DCASE_CUDA_CHECK(cudaIpcOpenMemHandle(&b, hs[2 * r + 1], cudaIpcMemLazyEnablePeerAccess));
        h->opened[h->n_opened++] = a;
        h->opened[h->n_opened++] = b;
        pg[r] = (float*)a; ps[r] = (uint32_t*)b;
    }
    DCASE_CUDA_CHECK(cudaMemcpy(h->d_peer_grads, pg, h->world * sizeof(float*), cudaMemcpyHostToDevice));
    DCASE_CUDA_CHECK(cudaMemcpy(h->d_peer_sync, ps, h->world * sizeof(uint32_t*), cudaMemcpyHostToDevice));
    return DCASE_OK;
}

float* dcase_p2p_grads(dcase_p2p* h) { return h ? h->grads : nullptr; }

int dcase_p2p_begin_step(dcase_p2p* h, void* stream) {
    DCASE_REQUIRE(h && h->d_peer_sync, "not connected");
    DCASE_PROF("p2p_wait_done", (cudaStream_t)stream);
    p2p_wait_done_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(h->sync, h->world);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int dcase_p2p_adam_ema_step(dcase_ctx* ctx, dcase_p2p* h, float* p, float* m, float* v, float* p_ema, float lr,
                            float beta1, float beta2, float eps, int step_t, float ema_alpha, const void* scalars,
                            void* stream_) {
    cudaStream_t s = (cudaStream_t)stream_;
    DCASE_REQUIRE(ctx && h && h->d_peer_sync && p && m && v, "null argument / not connected");
    DCASE_REQUIRE(scalars || step_t >= 1, "Adam step count starts at 1");
    DCASE_PROF("adam_ema_p2p", s);
    p2p_signal_ready_kernel<<<1, 32, 0, s>>>(h->sync, h->d_peer_sync, h->world, h->rank);
    DCASE_LAUNCH_CHECK();
    float bc1 = 1.f, bc2 = 1.f;                             // same arithmetic as dcase_adam_ema_step (api.cu)
    if (step_t >= 1) {
        bc1 = (float)(1.0 - pow((double)beta1, (double)step_t));
        bc2 = (float)(1.0 - pow((double)beta2, (double)step_t));
    }
    long long blocks = ((long long)h->n / 4 + 255) / 256;
    if (blocks > ctx->num_sms * 2) blocks = ctx->num_sms * 2;     // every CTA must be co-resident: they all spin on the flags
    adam_ema_p2p_kernel<<<(int)blocks, 256, 0, s>>>(p, h->d_peer_grads, h->sync, h->d_peer_sync, h->world, h->rank, m, v,
                                                   p_ema, (long long)h->n, lr, beta1, beta2, eps, bc1, bc2, ema_alpha,
                                                   (const DcaseStepScalars*)scalars);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int dcase_p2p_destroy(dcase_p2p* h) {
    if (!h) return DCASE_OK;
    cudaDeviceSynchronize();
    for (int i = 0; i < h->n_opened; ++i) cudaIpcCloseMemHandle(h->opened[i]);
    cudaFree(h->d_peer_grads); cudaFree(h->d_peer_sync); cudaFree(h->sync); cudaFree(h->grads);
    delete h;
    return DCASE_OK;
}

}  // extern "C"

// =====================================================================================================================
// SyncBN: exact-global-batch BatchNorm statistics under data parallelism
// =====================================================================================================================
// The reference at a global batch of 192 on ONE device normalises every BatchNorm2d over all 192 clips
// (baseline/models/CNN.py:49 in train mode); per-replica statistics (the default here) normalise over the rank's 24.
// This mode makes the N replicas reproduce the one-device result: the per-channel sums every BatchNorm needs -- forward
// (sum x, sum x^2; block 0: the 54 tap moments of its input) and backward (sum dy, sum dy xhat; block 0: the {U | S2}
// accumulator) -- are summed over the ranks between the kernel that produces them and the kernel that consumes them.
//
// One all-reduce = ONE single-CTA kernel per rank, no NCCL: every rank PUSHES its values into a mailbox in every peer's
// memory (its own included), publishes an epoch flag next to it, spins on its LOCAL flags until every rank's values of
// this epoch have arrived, and sums the world's contributions in rank order (all replicas obtain bit-identical sums).
// Mailboxes are double buffered by epoch parity: a rank can be at most one epoch ahead of a peer on the same slot
// (finishing epoch e + 1 needs the peer's flag of e + 1, which the peer only raises after it has read epoch e).  The
// epoch counters live in device memory and are advanced by the kernel itself, so a captured CUDA graph replays.
// Each (model, layer, direction) owns a slot; every rank issues the same sequence of collectives in stream order and,
// in this mode, the teacher forward runs on the student's stream (two streams that spin on peers could map to one
// hardware queue in a different order on different ranks).
namespace {
constexpr int kSyncSlots = 16;
constexpr int kSyncSlotBytes = 8192;          // 2048 floats: the block-0 backward accumulator is the largest message

__device__ __forceinline__ double ld_cv(const double* p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_cv(const float* p) {
    float v;
    asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
}  // namespace

struct dcase_syncbn {
    int world, rank;
    unsigned char* mail;            // local: [slot][parity][rank][kSyncSlotBytes] | flags u32 [slot][world] | epochs u32 [slot]
    size_t flags_off, epochs_off, bytes;
    unsigned char** d_peer_mail;    // device array [world]
    void* opened[kMaxWorld];
    int n_opened;
};

namespace {

template <typename T>
__global__ void __launch_bounds__(256)
syncbn_allreduce_kernel(T* __restrict__ vals, int n, int slot, unsigned char* mail, unsigned char* const* __restrict__ peer_mail,
                        size_t flags_off, size_t epochs_off, int world, int rank) {
    __shared__ uint32_t epoch_s;
    uint32_t* epochs = reinterpret_cast<uint32_t*>(mail + epochs_off);
    if (threadIdx.x == 0) {
        epoch_s = ld_volatile_u32(epochs + slot) + 1u;
        st_volatile_u32(epochs + slot, epoch_s);
    }
    __syncthreads();
    const uint32_t epoch = epoch_s;
    const size_t box = ((size_t)(slot * 2 + (int)(epoch & 1u)) * world) * kSyncSlotBytes;
    for (int r = 0; r < world; ++r) {                       // push: coalesced stores into every rank's mailbox
        T* dst = reinterpret_cast<T*>(peer_mail[r] + box + (size_t)rank * kSyncSlotBytes);
        for (int i = threadIdx.x; i < n; i += 256) dst[i] = vals[i];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world) {
        st_volatile_u32(reinterpret_cast<uint32_t*>(peer_mail[threadIdx.x] + flags_off) + slot * world + rank, epoch);
        const uint32_t* mine = reinterpret_cast<const uint32_t*>(mail + flags_off) + slot * world + threadIdx.x;
        while (!reached(ld_volatile_u32(mine), epoch)) __nanosleep(32);
    }
    __syncthreads();
    __threadfence_system();
    for (int i = threadIdx.x; i < n; i += 256) {
        T s = 0;
        for (int r = 0; r < world; ++r)                     // fixed order: bit-identical on every rank
            s += ld_cv(reinterpret_cast<const T*>(mail + box + (size_t)r * kSyncSlotBytes) + i);
        vals[i] = s;
    }
}

template <typename T>
int syncbn_allreduce(const dcase_syncbn* h, T* vals, int n, int slot, cudaStream_t s) {
    DCASE_REQUIRE(h && h->d_peer_mail, "SyncBN group not connected");
    DCASE_REQUIRE(slot >= 0 && slot < kSyncSlots && n >= 1 && (size_t)n * sizeof(T) <= (size_t)kSyncSlotBytes, "bad slot / size");
    DCASE_PROF("syncbn_allreduce", s);
    syncbn_allreduce_kernel<T><<<1, 256, 0, s>>>(vals, n, slot, h->mail, h->d_peer_mail, h->flags_off, h->epochs_off,
                                                 h->world, h->rank);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

}  // namespace

int syncbn_world(const dcase_syncbn* h) { return h ? h->world : 1; }
int syncbn_allreduce_f64(const dcase_syncbn* h, double* vals, int n, int slot, cudaStream_t s) {
    return syncbn_allreduce<double>(h, vals, n, slot, s);
}
int syncbn_allreduce_f32(const dcase_syncbn* h, float* vals, int n, int slot, cudaStream_t s) {
    return syncbn_allreduce<float>(h, vals, n, slot, s);
}

extern "C" {

int dcase_syncbn_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int dcase_syncbn_create(dcase_ctx* ctx, int world, int rank, dcase_syncbn** out, void* handle_out) {
    DCASE_REQUIRE(ctx && out && handle_out, "null argument");
    DCASE_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "bad world / rank");
    dcase_syncbn* h = new dcase_syncbn();
    h->world = world; h->rank = rank; h->n_opened = 0; h->mail = nullptr; h->d_peer_mail = nullptr;
    h->flags_off = (size_t)kSyncSlots * 2 * world * kSyncSlotBytes;
    h->epochs_off = h->flags_off + (size_t)kSyncSlots * world * sizeof(uint32_t);
    h->bytes = h->epochs_off + kSyncSlots * sizeof(uint32_t);
    DCASE_CUDA_CHECK(cudaMalloc(&h->mail, h->bytes));
    DCASE_CUDA_CHECK(cudaMemset(h->mail, 0, h->bytes));
    DCASE_CUDA_CHECK(cudaMalloc(&h->d_peer_mail, world * sizeof(unsigned char*)));
    DCASE_CUDA_CHECK(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle_out, h->mail));
    DCASE_CUDA_CHECK(cudaDeviceSynchronize());
    *out = h;
    return DCASE_OK;
}

// all_handles: world x dcase_syncbn_handle_bytes() bytes, rank-major
int dcase_syncbn_connect(dcase_syncbn* h, const void* all_handles) {
    DCASE_REQUIRE(h && all_handles, "null argument");
    unsigned char* pm[kMaxWorld];
    const cudaIpcMemHandle_t* hs = (const cudaIpcMemHandle_t*)all_handles;
    for (int r = 0; r < h->world; ++r) {
        if (r == h->rank) { pm[r] = h->mail; continue; }
        void* a = nullptr;
        DCASE_CUDA_CHECK(cudaIpcOpenMemHandle(&a, hs[r], cudaIpcMemLazyEnablePeerAccess));
        h->opened[h->n_opened++] = a;
        pm[r] = (unsigned char*)a;
    }
    DCASE_CUDA_CHECK(cudaMemcpy(h->d_peer_mail, pm, h->world * sizeof(unsigned char*), cudaMemcpyHostToDevice));
    return DCASE_OK;
}

// Attach (or with NULL detach) the group: from then on every train-mode forward / backward through this context
// normalises over the global batch.
int dcase_ctx_set_syncbn(dcase_ctx* ctx, dcase_syncbn* h) {
    DCASE_REQUIRE(ctx, "null context");
    DCASE_REQUIRE(!h || h->d_peer_mail, "SyncBN group not connected");
    ctx->syncbn = h;
    return DCASE_OK;
}

int dcase_syncbn_allreduce(dcase_syncbn* h, void* vals, int n, int is_double, int slot, void* stream) {
    DCASE_REQUIRE(h && vals, "null argument");
    return is_double ? syncbn_allreduce<double>(h, (double*)vals, n, slot, (cudaStream_t)stream)
                     : syncbn_allreduce<float>(h, (float*)vals, n, slot, (cudaStream_t)stream);
}

int dcase_syncbn_destroy(dcase_syncbn* h) {
    if (!h) return DCASE_OK;
    cudaDeviceSynchronize();
    for (int i = 0; i < h->n_opened; ++i) cudaIpcCloseMemHandle(h->opened[i]);
    cudaFree(h->d_peer_mail); cudaFree(h->mail);
    delete h;
    return DCASE_OK;
}

}  // extern "C"
