// conv3x3 (64 -> 64 channels, stride 1, pad 1; CNN.py:46-47) on the 5th-gen tensor cores: forward, data
// gradient and weight gradient as implicit GEMMs (tcgen05.mma kind::tf32, fp32 accumulators in TMEM).
//
// Tile = 16 frames x 8 mel slots = 128 output pixels of one clip.  The 18 x PITCH input halo (zero filled
// outside the clip) is staged ONCE in shared memory in the tensor core's swizzled operand layout; the 9 taps are
// the SAME buffer at 9 start addresses (the swizzle is a function of the absolute shared-memory address, so row
// shifts need no re-layout; pinned by tests/test_gpu_tcgen05.py).  Layer 1 (F = 16) covers the mel axis with two
// 8-wide halves (PITCH = 10); layer 2 (F = 4) uses PITCH = 8 with 4 valid slots.
//
//   forward / dgrad : D[pixel][n] = sum_tap sum_c halo[pixel + tap][c] * W[tap][n][c]     M=128, N=64, K=9*64
//                     the 144 KB weight image arrives by ONE bulk-async copy group (cp.async.bulk, mbarrier tx)
//   wgrad           : D[n][c]    += sum_pixel d_pre[pixel][n] * halo[pixel + tap][c]      M=64, N=32, K=pixels
//                     (MN-major operands, SWIZZLE_128B_BASE32B), all 9 taps of one input-channel half per CTA,
//                     accumulated in TMEM over all tiles of the persistent CTA
// Both are warp-specialised (TMA producer / MMA issuer warps / epilogue warps) around mbarrier rings.
#include <cuda_fp16.h>

#include "cnn.cuh"
#include "tc.cuh"
#include "tma.cuh"            // CUtensorMap helpers (the encoder is fetched through cudaGetDriverEntryPoint)

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;

int encode_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
               CUtensorMapSwizzle swz, CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT32) {
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = g_encode_tiled(map, dtype, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box,
                                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { dcase_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return DCASE_ERR_CUDA; }
    return DCASE_OK;
}

}  // namespace

int dcase_tma_init() {
    if (!g_encode_tiled) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        DCASE_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) { dcase_set_error("cuTensorMapEncodeTiled is not available"); return DCASE_ERR_STATE; }
        g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
    }
    return DCASE_OK;
}

int make_act_map(CUtensorMap* map, const float* base, int B, int T_l, int F, int box_rows, int pitch, CUtensorMapSwizzle swz) {
    const cuuint64_t dims[4] = {64, (cuuint64_t)F, (cuuint64_t)T_l, (cuuint64_t)B};
    const cuuint64_t strides[3] = {64 * sizeof(float), (cuuint64_t)F * 64 * sizeof(float), (cuuint64_t)T_l * F * 64 * sizeof(float)};
    const cuuint32_t box[4] = {32, (cuuint32_t)pitch, (cuuint32_t)box_rows, 1};
    return encode_map(map, base, 4, dims, strides, box, swz);
}

// the same view of an fp16 copy of the activation: one 128-byte row holds all 64 channels of a pixel
int make_act_map_h(CUtensorMap* map, const void* base, int B, int T_l, int F, int box_rows, int pitch) {
    const cuuint64_t dims[4] = {64, (cuuint64_t)F, (cuuint64_t)T_l, (cuuint64_t)B};
    const cuuint64_t strides[3] = {64 * 2, (cuuint64_t)F * 64 * 2, (cuuint64_t)T_l * F * 64 * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)pitch, (cuuint32_t)box_rows, 1};
    return encode_map(map, base, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
}

int make_rows_map(CUtensorMap* map, const float* base, long long n_rows, int box_rows, CUtensorMapSwizzle swz) {
    const cuuint64_t dims[2] = {64, (cuuint64_t)n_rows};
    const cuuint64_t strides[1] = {64 * sizeof(float)};
    const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    return encode_map(map, base, 2, dims, strides, box, swz);
}

int make_matrix_map(CUtensorMap* map, const float* base, long long cols, long long rows, long long ld, int box_cols,
                    int box_rows, CUtensorMapSwizzle swz) {
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    return encode_map(map, base, 2, dims, strides, box, swz);
}

namespace {

constexpr int kTile = 128;
constexpr int kHaloRows = 192;                    // >= 18 * 10 + 2, multiple of 8
constexpr int kHaloBlk = kHaloRows * 128;         // bytes per 32-channel block
constexpr int kWImgBytes = 9 * 16384;             // [9 taps][2 k-blocks][64 rows][128 B]

// W [n][c][tap] -> shared-memory images (K-major SW128 B operands, rows = output channel of the pass)
//   forward: image[tap ][c / 32][row n][c % 32] = W[n][c][tap]
//   dgrad  : image[8-tap][n / 32][row c][n % 32] = W[n][c][tap]      (mirrored taps, transposed channels)
//   forward, fp16 (the forward pass multiplies fp16 copies: same 10 explicit mantissa bits as tf32, half the operand bytes):
//            image[tap][row n][c] = W[n][c][tap], rows of 64 fp16 = 128 B, 16-byte chunks swizzled by row
struct ConvWPrepArgs {     // blockIdx.y = layer (both 64 -> 64 conv layers of the CNN in one launch)
    const float* w[2];
    float* img_fwd[2];     // nullable
    float* img_dgrad[2];
    __half* img_fwd_h[2];  // nullable
};
__global__ void conv_w_image_kernel(ConvWPrepArgs a) {
    const float* __restrict__ w = a.w[blockIdx.y];
    float* __restrict__ img_fwd = a.img_fwd[blockIdx.y];
    float* __restrict__ img_dgrad = a.img_dgrad[blockIdx.y];
    __half* __restrict__ img_h = a.img_fwd_h[blockIdx.y];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 64 * 64 * 9; i += gridDim.x * blockDim.x) {
        const int tap = i % 9, c = (i / 9) & 63, n = i / 576;
        const float wv = __ldg(w + i);
        const float v = tc::tf32_rn(wv);
        if (img_fwd) img_fwd[(tap * 16384 + (c >> 5) * 8192 + tc::sw128_off(n, c & 31)) >> 2] = v;
        img_dgrad[((8 - tap) * 16384 + (n >> 5) * 8192 + tc::sw128_off(c, n & 31)) >> 2] = v;
        if (img_h) img_h[(tap * 8192 + (n >> 3) * 1024 + (n & 7) * 128 + ((((c >> 3) ^ n) & 7) << 4) + (c & 7) * 2) >> 1] = __float2half_rn(wv);
    }
}

struct TileGeom {
    int b, t0, f0;
};
__device__ __forceinline__ TileGeom decode_tile(int tile, int halves, int tblocks) {
    TileGeom g;
    const int h = tile % halves;
    const int r = tile / halves;
    g.f0 = h * 8;
    g.t0 = (r % tblocks) * 16;
    g.b = r / tblocks;
    return g;
}

// ---------------------------------------------------------------------------------------------
// forward / dgrad, warp-specialised:  warp 0 = TMA producer, warps 1 and 6 = MMA issuers (even / odd tiles: one warp
// sustains ~7 uniform-datapath instructions = ~105 cycles per MMA, two keep the tensor core at its 48-cycle
// shared-memory operand floor for M = 128, N = 64 tf32), warps 2..5 = epilogue.
// The 144 KB weight image (all 64 output channels: N = 64 halves the shared-memory operand traffic per MAC
// compared with two N = 32 passes) leaves room for THREE 24 KB staging units, each one 32-channel block of the
// 18 x PITCH input halo of a tile (one TMA box).  The MMA loop runs k-block outer, so a unit is released after 36
// MMAs and refilled with the next tile's block while the other block is being multiplied: loads, MMAs and
// epilogues of consecutive tiles overlap (full / empty mbarrier ring, two TMEM accumulators).
// ---------------------------------------------------------------------------------------------
// HALF = true (the forward pass): input and weights are fp16 copies, kind::f16 MMAs with K = 16.  A pixel's 64 channels
// are ONE 128-byte row, so a tile's whole halo is one staging unit (one TMA box set) and a tile takes 36 MMAs instead of
// 72; the weight image shrinks to 72 KB.
constexpr int kUnits = 3;
constexpr int kConv2SmemBytes = kWImgBytes + kUnits * kHaloBlk + 256 + 18 * 8 + 4 * 2048;
constexpr int kConv2Threads = 352;          // warps 0 TMA, 1 and 6 MMA, 2..5 epilogue; HALF: 7..10 a second epilogue quartet
constexpr int kHalfImgBytes = 9 * 8192;

template <int PITCH, bool HALF>
__global__ void __launch_bounds__(kConv2Threads, 1)
conv3x3_tma_kernel(const __grid_constant__ CUtensorMap in_map4, const __grid_constant__ CUtensorMap in_map2, int B, int T_l, int F,
                   const float* __restrict__ w_img,
                   const float* __restrict__ bias, float* __restrict__ out, double* __restrict__ stats) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* Wi = smem;                               // 144 KB
    unsigned char* units = smem + kWImgBytes;               // kUnits x 24 KB (1024-B aligned)
    float* bias_s = reinterpret_cast<float*>(units + kUnits * kHaloBlk);      // [64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 64);
    uint64_t* full = bars;                                  // [kUnits]   TMA -> MMA
    uint64_t* empty = bars + kUnits;                        // [kUnits]   MMA done -> TMA
    uint64_t* acc_full = bars + 2 * kUnits;                 // [2]        MMA done -> epilogue
    uint64_t* acc_empty = bars + 2 * kUnits + 2;            // [2]        epilogue drained -> MMA
    uint64_t* w_bar = bars + 2 * kUnits + 4;
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 16);
    // epilogue staging, 2 KB per epilogue warp (128-B aligned): behind the barriers, or -- HALF: eight warps -- in the half
    // of the weight region the fp16 image leaves free
    unsigned char* stage_base = HALF ? Wi + kHalfImgBytes : reinterpret_cast<unsigned char*>(bars + 18);
    constexpr int kEpiWarps = HALF ? 8 : 4;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform for the compiler: role branches stay uniform
    constexpr int halves = PITCH == 10 ? 2 : 1;
    const int tblocks = (T_l + 15) / 16;
    const int n_tiles = B * tblocks * halves;
    if ((tc::smem_u32(smem) & 1023u) != 0) __trap();

    if (tid == 0) {
        for (int i = 0; i < kUnits; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], 32 * kEpiWarps); }
        tc::mbar_init(w_bar, 1);
        tc::fence_mbar_init();
    }
    if (tid < 64) bias_s[tid] = bias ? __ldg(bias + tid) : 0.f;
    if (warp == 1) tc::tmem_alloc(tmem_base_s, 128);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            constexpr int kTapBytes = HALF ? 8192 : 16384;
            mbar_expect_tx(w_bar, 9 * kTapBytes);
            for (int i = 0; i < 9; ++i)
                bulk_g2s(Wi + i * kTapBytes, reinterpret_cast<const unsigned char*>(w_img) + i * kTapBytes, kTapBytes, w_bar);
            int u = 0, n = 0;                       // staging unit of the next box and how often it has been used
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const TileGeom g = decode_tile(tile, halves, tblocks);
#pragma unroll 1
                for (int kb = 0; kb < (HALF ? 1 : 2); ++kb) {
                    if (n >= 1) tc::mbar_wait(&empty[u], (n - 1) & 1);
                    // the 18 halo rows arrive as five boxes (4 + 4 + 4 + 4 + 2 frames, 1024-byte aligned pieces): the TMA
                    // unit works on boxes concurrently but walks the 128-byte rows of one box slowly (~30 ns per row)
                    mbar_expect_tx(&full[u], 18 * PITCH * 128);
                    unsigned char* dst = units + u * kHaloBlk;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        tma_load_4d(dst + q * 4 * PITCH * 128, &in_map4, 32 * kb, g.f0 - 1, g.t0 - 1 + 4 * q, g.b, &full[u]);
                    tma_load_4d(dst + 16 * PITCH * 128, &in_map2, 32 * kb, g.f0 - 1, g.t0 + 15, g.b, &full[u]);
                    if (++u == kUnits) { u = 0; ++n; }
                }
            }
        }
    } else if (warp == 1 || warp == 6) {
        // all 32 lanes run the loop with warp-uniform values; one lane is elected per MMA / commit (tc.cuh).
        // Issuer w takes the tiles it = w, w + 2, ... (accumulator w); unit j = 2 it + kb of the producer's sequence.
        const int w = warp == 1 ? 0 : 1;
        tc::mbar_wait(w_bar, 0);
        constexpr uint32_t idesc = HALF ? tc::idesc_f16(128, 64, 0, 0) : tc::idesc_tf32(128, 64, 0, 0);
        constexpr int kKb = HALF ? 1 : 2;              // staging units (32-channel blocks) per tile
        constexpr int kTapBytes = HALF ? 8192 : 16384;
        const uint32_t a_hi = tc::desc_hi(PITCH * 128, 2), b_hi = tc::desc_hi(1024, 2);
        const uint32_t b_lo0 = tc::desc_lo(tc::smem_u32(Wi), 16);
        const uint32_t u_lo0 = tc::desc_lo(tc::smem_u32(units), 16);
        const uint32_t d = tmem + w * 64;
        int use = 0;                                   // how often this issuer's accumulator has been used
#ifdef DCASE_CONV_TIMING
        long long t_acc = 0, t_full = 0, t_issue = 0, tt;
        const long long t_begin = clock64();
#define TICK() tt = clock64()
#define TOCK(acc) acc += clock64() - tt
#else
#define TICK()
#define TOCK(acc)
#endif
        for (int it = w; blockIdx.x + (long long)it * gridDim.x < n_tiles; it += 2, ++use) {
            TICK();
            if (use >= 1) tc::mbar_wait(&acc_empty[w], (use - 1) & 1);
            TOCK(t_acc);
#pragma unroll 1
            for (int kb = 0; kb < kKb; ++kb) {
                const int j = kKb * it + kb, u = j % kUnits, n = j / kUnits;
                TICK();
                tc::mbar_wait(&full[u], n & 1);
                TOCK(t_full);
                TICK();
                tc::fence_after_sync();
                const uint32_t a_lo0 = u_lo0 + u * (kHaloBlk >> 4);
                const uint32_t b_lo1 = b_lo0 + kb * (8192 >> 4);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const uint32_t al = a_lo0 + ((((1 + dy) * PITCH + 1 + dx) * 128 + k4 * 32) >> 4);
                        const uint32_t bl = b_lo1 + ((tap * kTapBytes + k4 * 32) >> 4);
                        const uint32_t acc = (kb > 0 || tap > 0 || k4 > 0) ? 1u : 0u;
                        if (HALF) tc::umma_f16_elect(d, al, a_hi, bl, b_hi, idesc, acc);
                        else tc::umma_tf32_elect(d, al, a_hi, bl, b_hi, idesc, acc);
                    }
                }
                tc::umma_commit_elect(&empty[u]);        // unit free once these MMAs have read it
                TOCK(t_issue);
            }
            tc::umma_commit_elect(&acc_full[w]);
        }
#ifdef DCASE_CONV_TIMING
        if (blockIdx.x == 3 && lane == 0)
            printf("mma warp %d: total %lld  wait acc_empty %lld  wait full %lld  issue %lld  tiles %d\n", w, clock64() - t_begin, t_acc,
                   t_full, t_issue, use);
#endif
    } else if (warp <= 5 || HALF) {
        // HALF: the MMAs of a tile take half as long, so the epilogue (accumulator -> staging -> 128-byte lines + statistics,
        // ~3300 cycles per tile and quartet) would bound the kernel: two quartets split the 64 output channels
        const int wq = warp & 3;                   // TMEM lane quadrant of this warp
        const int eh = warp >= 7 ? 1 : 0;          // second quartet: channels 32..63
        const int slot = wq + 4 * eh;
        const int ch_lo = HALF ? eh : 0, ch_hi = HALF ? eh + 1 : 2;
        const uint32_t stage_a = tc::smem_u32(stage_base) + (uint32_t)slot * 2048u;
        int it = 0;
        // BatchNorm batch statistics of the output (CNN.py:49) ride in the epilogue: in the read-back loop below a lane
        // always sees the same 16-byte channel chunk (lane & 7), so it keeps that chunk's sum / sum of squares for both
        // channel halves in 16 registers over its whole tile stream (fp32 over ~300 values, then fp64 atomics)
        float4 bs[2], bq[2];
        bs[0] = bs[1] = bq[0] = bq[1] = make_float4(0.f, 0.f, 0.f, 0.f);
#ifdef DCASE_CONV_TIMING
        long long t_wait = 0, t_ld = 0, t_st = 0, tt;
        const long long t_begin = clock64();
#endif
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int ac = it & 1;
            const TileGeom g = decode_tile(tile, halves, tblocks);
            TICK();
            tc::mbar_wait(&acc_full[ac], (it >> 1) & 1);
            TOCK(t_wait);
            TICK();
            tc::fence_after_sync();
            float acc[64];
            if (HALF) {
                const uint32_t ta = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(ac * 64 + 32 * eh);
                tc::tmem_ld16(ta, acc);              // this quartet's 32 columns sit in acc[0..31]
                tc::tmem_ld16(ta + 16, acc + 16);
                tc::tmem_ld_wait();
            } else {
                tc::tmem_ld_row64(tmem, wq, ac * 64, acc);
            }
            tc::fence_before_sync();
            mbar_arrive(&acc_empty[ac]);
            TOCK(t_ld);
            TICK();
            // A thread owns one pixel row (256 B), so direct stores would touch 32 different lines with 16 B each per
            // instruction (measured: 4450 cycles per tile, the kernel's bottleneck).  Instead 16 rows x 128 B at a time
            // go through a 2 KB per-warp staging tile (XOR-swizzled chunks) and leave as full 128-byte lines.
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                if (ch < ch_lo || ch >= ch_hi) continue;
#pragma unroll
                for (int rh = 0; rh < 2; ++rh) {
                    if ((lane >> 4) == rh) {
                        const uint32_t rb = stage_a + (uint32_t)(lane & 15) * 128u;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + 32 * ch + 4 * q);
                            const int e = (HALF ? 0 : 32 * ch) + 4 * q;
                            st_shared_v4(rb + (uint32_t)((q ^ (lane & 7)) << 4), acc[e] + b4.x, acc[e + 1] + b4.y, acc[e + 2] + b4.z,
                                         acc[e + 3] + b4.w);
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int idx = lane + 32 * q, r = idx >> 3, c = idx & 7;
                        const float4 v = ld_shared_v4(stage_a + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4));
                        const int prow = 32 * wq + 16 * rh + r;                // pixel row of the tile
                        const int t = g.t0 + (prow >> 3), f = g.f0 + (prow & 7);
                        if (t < T_l && f < F) {
                            *reinterpret_cast<float4*>(out + (((long long)g.b * T_l + t) * F + f) * 64 + 32 * ch + 4 * c) = v;
                            bs[ch].x += v.x; bs[ch].y += v.y; bs[ch].z += v.z; bs[ch].w += v.w;
                            bq[ch].x = fmaf(v.x, v.x, bq[ch].x); bq[ch].y = fmaf(v.y, v.y, bq[ch].y);
                            bq[ch].z = fmaf(v.z, v.z, bq[ch].z); bq[ch].w = fmaf(v.w, v.w, bq[ch].w);
                        }
                    }
                    __syncwarp();
                }
            }
            TOCK(t_st);
        }
#ifdef DCASE_CONV_TIMING
        if (blockIdx.x == 3 && tid == 64)
            printf("epilogue: total %lld  wait acc_full %lld  ldtm %lld  stores %lld  tiles %d\n", clock64() - t_begin, t_wait, t_ld, t_st, it);
#endif
        if (stats) {
            // lanes with equal (lane & 7) hold the same channels: fold over lane bits 3, 4; lanes 0..7 then park the warp's
            // 128 partials in its staging tile, the four epilogue warps meet, and thread e adds entry e of all four
            float part[16] = {bs[0].x, bs[0].y, bs[0].z, bs[0].w, bs[1].x, bs[1].y, bs[1].z, bs[1].w,
                              bq[0].x, bq[0].y, bq[0].z, bq[0].w, bq[1].x, bq[1].y, bq[1].z, bq[1].w};
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                part[i] += __shfl_xor_sync(0xffffffffu, part[i], 8);
                part[i] += __shfl_xor_sync(0xffffffffu, part[i], 16);
            }
            float* mine = reinterpret_cast<float*>(stage_base + slot * 2048);
            if (lane < 8) {
#pragma unroll
                for (int i = 0; i < 16; ++i)      // entry = (sum | sumsq) * 64 + channel, channel = 32 ch + 4 (lane & 7) + e
                    mine[(i >> 3) * 64 + ((i >> 2) & 1) * 32 + 4 * lane + (i & 3)] = part[i];
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
            if (eh == 0) {                             // the first quartet's 128 threads own one entry each
                const int e = 32 * wq + lane;
                const float* all = reinterpret_cast<const float*>(stage_base) + (HALF ? ((e >> 5) & 1) * 2048 : 0);   // the quartet that owns
                const double tot = (double)all[e] + (double)all[512 + e] + (double)all[1024 + e] + (double)all[1536 + e];   // the channel's half
                atomicAdd(stats + e, tot);             // [0,64): sum, [64,128): sum of squares
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 128);
}

// ---------------------------------------------------------------------------------------------
// weight gradient, warp-specialised:  warp 0 = TMA producer, warps 1..3 = MMA issuers (one tap row dy each),
// warps 4..7 = final read-out.  grid = (CTAs, 2): blockIdx.y selects 32 of the 64 INPUT channels c, so the nine
// [64 n][32 c] accumulators (288 TMEM columns) of all taps stay resident over the CTA's whole tile stream and every
// tile is fetched once per half: d_pre tile (32 KB, two TMA boxes) + one 32-channel halo block (one box), both
// landing directly in the MN-major (SWIZZLE_128B_ATOM_32B) operand layout; 4-stage full / empty ring.
//   D_tap[n][c] += sum over the 8 mel slots of frame row r:  d_pre[(r, j)][n] * in[(r + dy, j + dx)][c]
// ---------------------------------------------------------------------------------------------
constexpr int kWgStages = 4;
constexpr int kWgStageBytes = 32768 + kHaloBlk;          // d_pre 2 x 16 KB | halo block 24 KB (18 * PITCH rows written)
constexpr int kWgSmemBytes = kWgStages * kWgStageBytes + 16 * 8 + 16;
constexpr int kWgThreads = 256;

template <int PITCH>
__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_tma_kernel(const __grid_constant__ CUtensorMap dpre_map, const __grid_constant__ CUtensorMap in_map, int B, int T_l,
                      int F, float* __restrict__ g_w) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgStages * kWgStageBytes);
    uint64_t* full = bars;                     // [kWgStages]  TMA -> MMA
    uint64_t* empty = bars + kWgStages;        // [kWgStages]  3 issuers done -> TMA
    uint64_t* done = bars + 2 * kWgStages;     // all MMAs of the CTA complete -> read-out
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 16);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int ch = blockIdx.y;
    constexpr int halves = PITCH == 10 ? 2 : 1;
    const int tblocks = (T_l + 15) / 16;
    const int n_tiles = B * tblocks * halves;
    if ((tc::smem_u32(smem) & 1023u) != 0) __trap();

    // the 8 halo rows behind the TMA box (read by the dx = +1 tap of the last frame row when PITCH = 8) must be finite:
    // they only ever meet zero d_pre rows
    for (int i = tid; i < kWgStages * 64; i += kWgThreads)
        reinterpret_cast<float4*>(smem + (i >> 6) * kWgStageBytes + 32768 + 18 * PITCH * 128)[i & 63] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
        for (int i = 0; i < kWgStages; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 3); }
        tc::mbar_init(done, 3);
        tc::fence_mbar_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_base_s, 512);
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_base_s;
    const bool any = (int)blockIdx.x < n_tiles;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const int st = it % kWgStages, n = it / kWgStages;
                if (n >= 1) tc::mbar_wait(&empty[st], (n - 1) & 1);
                const TileGeom g = decode_tile(tile, halves, tblocks);
                unsigned char* dst = smem + st * kWgStageBytes;
                mbar_expect_tx(&full[st], 32768 + 18 * PITCH * 128);
                tma_load_4d(dst, &dpre_map, 0, g.f0, g.t0, g.b, &full[st]);
                tma_load_4d(dst + 16384, &dpre_map, 32, g.f0, g.t0, g.b, &full[st]);
                tma_load_4d(dst + 32768, &in_map, 32 * ch, g.f0 - 1, g.t0 - 1, g.b, &full[st]);
            }
        }
    } else if (warp <= 3) {
        // issuer of tap row dy = warp - 2: 3 taps x 16 K steps per tile; all lanes run the loop, one is elected per MMA
        const int dyi = warp - 1;                              // 0, 1, 2
        constexpr uint32_t idesc = tc::idesc_tf32(64, 32, 1, 1);
        const uint32_t hi = tc::desc_hi(512, 1);
        const uint32_t s_lo0 = tc::desc_lo(tc::smem_u32(smem), 16384);
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int st = it % kWgStages;
            tc::mbar_wait(&full[st], (it / kWgStages) & 1);
            tc::fence_after_sync();
            const uint32_t a_lo = s_lo0 + st * (kWgStageBytes >> 4);
            const uint32_t b_lo = a_lo + ((32768 + dyi * PITCH * 128) >> 4);
            const uint32_t acc1 = it > 0 ? 1u : 0u;
#pragma unroll
            for (int dxi = 0; dxi < 3; ++dxi) {
                const uint32_t d = tmem + (dyi * 3 + dxi) * 32;
#pragma unroll
                for (int r = 0; r < 16; ++r)               // K step = the 8 mel slots of frame row r of the tile
                    tc::umma_tf32_elect(d, a_lo + r * (1024 >> 4), hi, b_lo + ((r * PITCH + dxi) * 128 >> 4), hi, idesc,
                                        r > 0 ? 1u : acc1);
            }
            tc::umma_commit_elect(&empty[st]);
        }
        tc::umma_commit_elect(done);
    } else if (any) {
        const int wq = warp & 3;
        tc::mbar_wait(done, 0);
        tc::fence_after_sync();
        const int n = 16 * wq + lane;                       // M = 64 accumulators: row m in TMEM lane 32*(m/16) + m%16
        const bool own = lane < 16;
        const uint32_t tbase = tmem + ((uint32_t)(wq * 32) << 16);
        // g_w[n][c][tap]: the 9 taps of 4 consecutive channels are 36 contiguous, 16-byte aligned floats -> 9 vector REDs
#pragma unroll 1
        for (int c0 = 0; c0 < 32; c0 += 4) {
            float v[9][4];
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) tc::tmem_ld4(tbase + tap * 32 + c0, v[tap]);
            tc::tmem_ld_wait();
            if (own) {
                float* dst = g_w + n * 576 + (32 * ch + c0) * 9;
#pragma unroll
                for (int q = 0; q < 9; ++q) {
                    float e[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) { const int idx = 4 * q + k; e[k] = v[idx % 9][idx / 9]; }
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * q), "f"(e[0]), "f"(e[1]), "f"(e[2]), "f"(e[3])
                                 : "memory");
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

#define DCASE_TRY_RC(expr) do { int rc__ = (expr); if (rc__ != DCASE_OK) return rc__; } while (0)

int grid_for(int n_tiles, int num_sms, int per_sm) {
    const int g = num_sms * per_sm;
    return n_tiles < g ? n_tiles : g;
}

}  // namespace

int conv_tc_kernels_init() {
    { const int rc = dcase_tma_init(); if (rc != DCASE_OK) return rc; }
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(conv_wgrad_tma_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes));
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(conv_wgrad_tma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes));
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_tma_kernel<10, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConv2SmemBytes));
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_tma_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConv2SmemBytes));
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_tma_kernel<10, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConv2SmemBytes));
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_tma_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConv2SmemBytes));
    return DCASE_OK;
}

int launch_conv_w_prep(const float* w1, float* w_fwd1, float* w_dgrad1, void* w_fwd1_h, const float* w2, float* w_fwd2,
                       float* w_dgrad2, void* w_fwd2_h, cudaStream_t s) {
    DCASE_PROF("conv_w_prep", s);
    ConvWPrepArgs a{{w1, w2}, {w_fwd1, w_fwd2}, {w_dgrad1, w_dgrad2}, {(__half*)w_fwd1_h, (__half*)w_fwd2_h}};
    conv_w_image_kernel<<<dim3(72, 2), 256, 0, s>>>(a);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int launch_conv3x3(const float* in, int B, int T_l, int F, const float* w_img, const float* bias, float* out,
                   double* stats, int num_sms, cudaStream_t s) {
    DCASE_PROF(bias ? (F == 16 ? "conv3x3_fwd_l1" : "conv3x3_fwd_l2") : (F == 16 ? "conv3x3_dgrad_l1" : "conv3x3_dgrad_l2"), s);
    DCASE_REQUIRE(F == 16 || F == 4, "conv3x3 is built for the 16- and 4-bin layers of cfg.crnn_kwargs");
    const int n_tiles = B * ((T_l + 15) / 16) * (F == 16 ? 2 : 1);
    const int gx = n_tiles < num_sms ? n_tiles : num_sms;
    CUtensorMap in_map4, in_map2;
    const int pitch = F == 16 ? 10 : 8;
    DCASE_TRY_RC(make_act_map(&in_map4, in, B, T_l, F, 4, pitch, CU_TENSOR_MAP_SWIZZLE_128B));
    DCASE_TRY_RC(make_act_map(&in_map2, in, B, T_l, F, 2, pitch, CU_TENSOR_MAP_SWIZZLE_128B));
    // BatchNorm batch statistics ([64] sums | [64] sums of squares, fp64) are accumulated by the conv epilogue itself
    if (stats) DCASE_CUDA_CHECK(cudaMemsetAsync(stats, 0, 128 * sizeof(double), s));
    if (F == 16) conv3x3_tma_kernel<10, false><<<gx, kConv2Threads, kConv2SmemBytes, s>>>(in_map4, in_map2, B, T_l, F, w_img, bias, out, stats);
    else conv3x3_tma_kernel<8, false><<<gx, kConv2Threads, kConv2SmemBytes, s>>>(in_map4, in_map2, B, T_l, F, w_img, bias, out, stats);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

// forward pass on fp16 copies of the input activation and of the weights (image from launch_conv_w_prep); fp32 output
int launch_conv3x3_h(const void* in_h, int B, int T_l, int F, const void* w_img_h, const float* bias, float* out,
                     double* stats, int num_sms, cudaStream_t s) {
    DCASE_PROF(F == 16 ? "conv3x3_fwd_l1" : "conv3x3_fwd_l2", s);
    DCASE_REQUIRE(F == 16 || F == 4, "conv3x3 is built for the 16- and 4-bin layers of cfg.crnn_kwargs");
    const int n_tiles = B * ((T_l + 15) / 16) * (F == 16 ? 2 : 1);
    const int gx = n_tiles < num_sms ? n_tiles : num_sms;
    CUtensorMap in_map4, in_map2;
    const int pitch = F == 16 ? 10 : 8;
    DCASE_TRY_RC(make_act_map_h(&in_map4, in_h, B, T_l, F, 4, pitch));
    DCASE_TRY_RC(make_act_map_h(&in_map2, in_h, B, T_l, F, 2, pitch));
    if (stats) DCASE_CUDA_CHECK(cudaMemsetAsync(stats, 0, 128 * sizeof(double), s));
    const float* w = reinterpret_cast<const float*>(w_img_h);
    if (F == 16) conv3x3_tma_kernel<10, true><<<gx, kConv2Threads, kConv2SmemBytes, s>>>(in_map4, in_map2, B, T_l, F, w, bias, out, stats);
    else conv3x3_tma_kernel<8, true><<<gx, kConv2Threads, kConv2SmemBytes, s>>>(in_map4, in_map2, B, T_l, F, w, bias, out, stats);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int launch_conv_wgrad(const float* d_pre, const float* in, int B, int T_l, int F, float* g_w, int num_sms,
                      cudaStream_t s) {
    DCASE_PROF(F == 16 ? "conv3x3_wgrad_l1" : "conv3x3_wgrad_l2", s);
    DCASE_REQUIRE(F == 16 || F == 4, "conv wgrad is built for the 16- and 4-bin layers of cfg.crnn_kwargs");
    const int n_tiles = B * ((T_l + 15) / 16) * (F == 16 ? 2 : 1);
    int gx = num_sms / 2;
    if (gx > n_tiles) gx = n_tiles;
    if (gx < 1) gx = 1;
    const int pitch = F == 16 ? 10 : 8;
    CUtensorMap dpre_map, in_map;
    DCASE_TRY_RC(make_act_map(&dpre_map, d_pre, B, T_l, F, 16, 8, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    DCASE_TRY_RC(make_act_map(&in_map, in, B, T_l, F, 18, pitch, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    if (F == 16) conv_wgrad_tma_kernel<10><<<dim3(gx, 2), kWgThreads, kWgSmemBytes, s>>>(dpre_map, in_map, B, T_l, F, g_w);
    else conv_wgrad_tma_kernel<8><<<dim3(gx, 2), kWgThreads, kWgSmemBytes, s>>>(dpre_map, in_map, B, T_l, F, g_w);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}
