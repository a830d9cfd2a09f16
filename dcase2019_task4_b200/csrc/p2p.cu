// Data-parallel gradient exchange FUSED with the optimizer over NVLink peer memory -- EXPERIMENTAL.
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
// STATUS (round 1): compiles for sm_100a; written after the round's GPU budget was spent -- not yet run on hardware.
// Enabled only by DCASE_DP_P2P=1 (dcase2019_task4_b200/dp.py); the default data-parallel step uses NCCL.
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
