// CNN block 0, fused:  conv3x3(1->64) + bias -> BatchNorm (folded) -> GLU -> Dropout(0.5) -> AvgPool(2,4)
// forward and backward, never materialising the [B,64,T,64] activation (340 MB at B = 24).
//
// Replaces (reference file:line): baseline/models/CNN.py:42-67 (block 0), :5-16 (GLU) and their autograd backward.
//
// Tile = 128 pixels = 2 frames x 64 mel bins of one clip = 16 pool windows.  Per tile, on the tensor core
// (tcgen05.mma kind::tf32, fp32 accumulators in TMEM):
//   MMA0  y   [p][c] = sum_j T[p][j] W0[c][j]         T = [9 taps | 1], W0 = [BN-folded conv weights | bias], K = 16
//   MMA1  lin [p][n] = sum_k y[p][k] Wg[n][k]         the GLU linear (K = 64)
//   MMA2  out [n][w] = sum_p z[p][n] P[w][p]          forward: the (2,4) average pool as a 0/1 matrix (K = 128)
//   MMA3  [U|S][m][j] += sum_p [DL|D2][p][m] T[p][j]  backward: every parameter gradient of the block (K = 128)
// CUDA cores only do the gate: g = sigmoid(y), dropout select, z = (lin + b) g, and in the backward
//   DL = dz g (grad wrt lin),  D2 = dz (lin + b) g (1 - g) (grad wrt y through the gate).
// Backward algebra: y is LINEAR in the 10 columns of T, so with U = DL^T T and S2 = D2^T T
//   dWg[n][k] = sum_p DL[p][n] y[p][k]  = sum_j U[n][j] W0[k][j]
//   S[c][j]   = sum_p dY[p][c] T[p][j]  = sum_n Wg[n][c] U[n][j] + S2[c][j]      (dY = DL Wg + D2)
// and S gives dbeta, dgamma and the conv0 weight gradient (cnn0_bwd_finalize).  No [pixels x 64] gradient tile is
// ever multiplied by a [pixels x 64] activation tile: one M=128, N=16 accumulator per tile stream.
//
// tf32 operands: kind::tf32 truncates fp32 operands to 10 mantissa bits, a mean relative shrink of
// 0.7213 * 2^-11 (log-uniform mantissa).  Tiles written per element (y, z, DL, D2) are stored unrounded and the
// shrink is undone in the constant operand (Wg) or in the epilogue scale; small per-pixel operands (taps) and
// weights are rounded to nearest.
#include "cnn.cuh"
#include "tc.cuh"

namespace {

constexpr int kTile = 128;
constexpr float kTruncComp = 1.f + 3.5221e-4f;
constexpr float kLog2e = 1.4426950408889634f;
// The gate sigmoid(y) is the block's MUFU load (84.9 M activations per pass) and the per-element CUDA-core work around it
// bounds both kernels, so everything that can ride on the tensor core does:
//   * MMA0 produces a pre-scaled ys = kGateScale * y (W0 carries the factor), MMA1's constant operand carries the inverse;
//   * the GLU bias rides in MMA1 as one more K = 8 slice: the tap rows hold two constant-1 columns and rows 64..127 of the
//     T0 block (free: W0 only has 64 rows) hold [0.5 b_hi | 0.5 b_lo] against them, so MMA1 delivers u = (lin + b) / 2;
//   * with a = tanh(y / 2) (one MUFU): sigmoid(y) = (1 + a) / 2, hence
//       forward   z  = (lin + b) g           = fma(u, a, u)                 dropout: a := -1 gives exactly 0
//       backward  DL = dz g                  = 1/2 * fma(dz, a, dz)
//                 D2 = dz (lin + b) g (1-g)  = -1/2 * (dz u) (a a - 1)
//     all as packed f32x2 instructions; the +-1/2 and the pool / dropout scale of dz are applied once to the [128][16]
//     accumulator at the end of the backward (DL and D2 only feed that linear reduction).
//   DCASE_GATE_TANH (default): ys = y / 2, a = tanh.approx(ys);  otherwise ys = -log2(e) y, a = 2 rcp(1 + ex2(ys)) - 1.
// tanh.approx.f32 has a relative error of up to 2^-11 (|dg| <= 2.4e-4); measured effect on the frame posteriors:
// tests/test_gpu_fullsize.py prints it, tests/scripts/precision_modes.py models it.
#ifndef DCASE_GATE_TANH
#define DCASE_GATE_TANH 1
#endif
#if DCASE_GATE_TANH
constexpr float kGateScale = 0.5f;
constexpr float kGateUnscale = 2.0f;
#else
constexpr float kGateScale = -kLog2e;
constexpr float kGateUnscale = -0.6931471805599453f;
#endif
constexpr float kLinScale = 0.5f * kGateUnscale;      // MMA1 delivers u = (lin + b) / 2 from the pre-scaled ys

// Parameter gradients of the block from the {U | S2} accumulator (cnn0_bwd_finalize): a kernel of its own, or the last CTA
// of cnn0_bwd_kernel (enabled = 1)
struct Cnn0FinArgs {
    int enabled;
    const double* mom;
    long long n_pix;
    const float *w, *b, *glu_w_raw;
    float pgs;
    float *g_w, *g_b, *g_gamma, *g_beta, *g_glu_w, *g_glu_b;
};

struct Cnn0Args {
    const float* x;        // [B][T][64] z-scored log-mel
    int B, T;
    const float* fold0;    // bn0_finalize output (cnn.cuh)
    const float* glu_w;    // [64][64]
    const float* glu_b;    // [64]
    DropoutCfg drop;
    float* out;            // fwd: [B][T/2][16][64], nullable
    void* out_h;           // fwd: the same as fp16 (operand of conv1's forward), nullable
    unsigned int* tile_ctr; // fwd: tile counter of the dynamic schedule (zeroed by the launcher)
    const float* d_out;    // bwd: grad of out
    float* us;             // bwd: [128][16] accumulators {U[64][16], S2[64][16]} (zeroed by the caller; us[10] = CTA ticket)
    Cnn0FinArgs fin;       // bwd: the finalize pass, run by the last CTA to finish
};

__device__ __forceinline__ float ex2_ftz(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_ftz(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void bar_sync_named(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// ---- shared-memory operand tiles, addressed with 32-bit shared-space addresses ----------------------------
// Both tcgen05 layouts used here are XOR swizzles of the 16-byte chunk index inside a 128-byte row, so the address
// of chunk c4 (channels 4 c4 .. 4 c4 + 3) of a thread's row is  (row_base ^ ((c4 & 7) << 4)) + (c4 >> 3) * 16 KB
// with the row's own swizzle bits folded into row_base (region bases are 1024-byte aligned).
__device__ __forceinline__ uint32_t krow_base(uint32_t region, int r) {    // K-major SWIZZLE_128B (tc::sw128_chunk)
    return region + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((r & 7) << 4));
}
__device__ __forceinline__ uint32_t mnrow_base(uint32_t region, int r) {   // MN-major SWIZZLE_128B_BASE32B
    return region + (uint32_t)(r * 128 + ((r & 3) << 5));
}
__device__ __forceinline__ uint32_t chunk_addr(uint32_t row_base, int c4) {
    return (row_base ^ (uint32_t)((c4 & 7) << 4)) + (uint32_t)(c4 >> 3) * 16384u;
}
__device__ __forceinline__ void sts128(uint32_t addr, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

// packed fp32 pairs (FFMA2 / FMUL2 on sm_100): half the issue slots of the per-element gate arithmetic
__device__ __forceinline__ uint64_t pk(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ void sts128_2(uint32_t addr, uint64_t lo, uint64_t hi) {
    asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(addr), "l"(lo), "l"(hi) : "memory");
}

// x rows t0-1 .. t0+2 (zero padded) of a tile = xs[4][66]; thread t of 256 owns elements t and (t < 8) t + 256.
// The tile's (clip, first frame) pair is carried incrementally (32-bit): a 64-bit division per tile and thread was
// 30 % of the kernel's instructions.
// rounds to the nearest tf32 (ties away), finite inputs only: 2 instructions instead of cvt.rna's 3
__device__ __forceinline__ float tf32_round_fast(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

struct XsRegs { float v0, v1; };
struct TilePos {          // tile -> clip b, first frame t0
    int b, t0;
    __device__ __forceinline__ void init(long long tile, int T) {
        const unsigned r0 = (unsigned)(2 * tile);
        b = (int)(r0 / (unsigned)T);
        t0 = (int)(r0 - (unsigned)b * (unsigned)T);
    }
    __device__ __forceinline__ void advance(int rows, int T) {      // rows >= 0
        t0 += rows;
        while (t0 >= T) { t0 -= T; ++b; }
    }
};
__device__ __forceinline__ float xs_value(const float* __restrict__ x, const TilePos& p, int T, int i) {
    const int hr = i / 66, hc = i - hr * 66;
    const int tt = p.t0 - 1 + hr, ff = hc - 1;
    const bool ok = tt >= 0 && tt < T && ff >= 0 && ff < 64;
    return ok ? __ldg(x + ((long long)p.b * T + tt) * 64 + ff) : 0.f;
}
__device__ __forceinline__ XsRegs xs_prefetch(const float* __restrict__ x, const TilePos& p, int T, int t) {
    XsRegs r;
    r.v0 = xs_value(x, p, T, t);
    r.v1 = t + 256 < 4 * 66 ? xs_value(x, p, T, t + 256) : 0.f;
    return r;
}
__device__ __forceinline__ void xs_commit(const XsRegs& r, float* xs, int t) {     // MMA0 operand: rounded to tf32 once, here
    xs[t] = tf32_round_fast(r.v0);
    if (t + 256 < 4 * 66) xs[t + 256] = tf32_round_fast(r.v1);
}

// operand rows of MMA0 for pixel `row` of a tile: [tap0..8, 1, 1, 0] -> chunks 0..2 of its row (column 9 meets the folded
// conv bias in MMA0 and 0.5 b_hi in MMA1's bias slice, column 10 meets 0.5 b_lo; W0 is zero there)
__device__ __forceinline__ void write_taps(const float* xs, int row, uint32_t t0_rowbase) {
    const int tr = row >> 6, f = row & 63;
    float tap[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) tap[k] = xs[(tr + k / 3) * 66 + f + (k % 3)];
    sts128(chunk_addr(t0_rowbase, 0), tap[0], tap[1], tap[2], tap[3]);
    sts128(chunk_addr(t0_rowbase, 1), tap[4], tap[5], tap[6], tap[7]);
    sts128(chunk_addr(t0_rowbase, 2), tap[8], 1.f, 1.f, 0.f);
}

// Wg[n][k] -> K-major SW128 B operand (two blocks of 64 rows), rounded to tf32 after the truncation compensation
__device__ __forceinline__ void stage_wg(const float* __restrict__ glu_w, unsigned char* Wb, int t, int nt) {
    // eight loads in flight per thread: as a rolled loop each of the 8-16 iterations waited out an L2 round trip, 5-10 us of
    // prologue in every one of the grid's CTAs (%globaltimer, round 2)
    for (int i0 = 0; i0 < 4096; i0 += 8 * nt) {
        float w8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int i = i0 + t + j * nt; w8[j] = i < 4096 ? __ldg(glu_w + i) : 0.f; }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = i0 + t + j * nt;
            if (i < 4096) {
                const int n = i >> 6, k = i & 63;
                *reinterpret_cast<float*>(Wb + (k >> 5) * 8192 + tc::sw128_off(n, k & 31)) = tc::tf32_rn(kLinScale * kTruncComp * w8[j]);
            }
        }
    }
}

// W0 (logical columns 16..31 of rows 0..63 of the T0 block): [wf[0..8][n], bf[n], 0...]; chunk 3 of the tap rows = 0
__device__ __forceinline__ void stage_w0(const float* __restrict__ fold0, unsigned char* T0, int t, int nt) {
    for (int i = t; i < 256; i += nt) {
        const int n = i >> 2, c = i & 3;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = 4 * c + e;
            v[e] = kGateScale * (k < 9 ? __ldg(fold0 + kFold0Wf + k * 64 + n) : (k == 9 ? __ldg(fold0 + kFold0Bf + n) : 0.f));
        }
        *reinterpret_cast<float4*>(T0 + tc::sw128_chunk(n, 4 + c)) = tc::tf32_rn4(make_float4(v[0], v[1], v[2], v[3]));
    }
    for (int r = t; r < 128; r += nt) *reinterpret_cast<float4*>(T0 + tc::sw128_chunk(r, 3)) = make_float4(0.f, 0.f, 0.f, 0.f);
}
// GLU bias as the B operand of MMA1's extra K = 8 slice: rows 64 + n of the T0 block, logical columns 24..31 (they face
// the tap rows' columns 8..15 = [tap8, 1, 1, 0 | 0 0 0 0]): [0, hi, lo, 0 | 0 0 0 0] with hi + lo = b[n] / 2 to 2^-22
__device__ __forceinline__ void stage_bias(const float* __restrict__ glu_b, unsigned char* T0, int t, int nt) {
    for (int n = t; n < 64; n += nt) {
        const float b = 0.5f * __ldg(glu_b + n);
        const float hi = tc::tf32_rn(b), lo = tc::tf32_rn(b - hi);
        *reinterpret_cast<float4*>(T0 + tc::sw128_chunk(64 + n, 4)) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(T0 + tc::sw128_chunk(64 + n, 5)) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(T0 + tc::sw128_chunk(64 + n, 6)) = make_float4(0.f, hi, lo, 0.f);
        *reinterpret_cast<float4*>(T0 + tc::sw128_chunk(64 + n, 7)) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// The MMA issue helpers are called by ALL lanes of one warp with warp-uniform arguments (tc.cuh: one lane is elected
// per instruction; issuing under `if (tid == 0)` costs ~16 instructions per MMA on the critical warp).
__device__ __forceinline__ void issue_mma0(uint32_t d_tmem, uint32_t t0_addr) {   // y = T W0^T, K = 16
    constexpr uint32_t idesc = tc::idesc_tf32(128, 64, 0, 0);
    const uint32_t lo = tc::desc_lo(t0_addr, 16), hi = tc::desc_hi(1024, 2);
#pragma unroll
    for (int k = 0; k < 2; ++k) tc::umma_tf32_elect(d_tmem, lo + 2 * k, hi, lo + 4 + 2 * k, hi, idesc, k);
}
__device__ __forceinline__ void issue_mma1(uint32_t d_tmem, uint32_t a_addr, uint32_t wb_addr) {   // lin = y Wg^T, K = 64
    constexpr uint32_t idesc = tc::idesc_tf32(128, 64, 0, 0);
    const uint32_t a_lo = tc::desc_lo(a_addr, 16), b_lo = tc::desc_lo(wb_addr, 16), hi = tc::desc_hi(1024, 2);
#pragma unroll
    for (int j = 0; j < 8; ++j)
        tc::umma_tf32_elect(d_tmem, a_lo + (((j >> 2) * 16384 + (j & 3) * 32) >> 4), hi, b_lo + (((j >> 2) * 8192 + (j & 3) * 32) >> 4), hi,
                            idesc, j > 0 ? 1u : 0u);
}

// lin = y Wg^T with y read IN PLACE from the MMA0 accumulator in tensor memory (an M = 128 accumulator has exactly the
// [lane = row][column = k] layout of a TMEM A operand): no shared-memory round trip for y, and the GEMM can be issued the
// moment MMA0 has completed, before any thread has touched y
// The first instruction is the bias slice (stage_bias): A = columns 8..15 of the tap rows, still in shared memory.  The
// next tile's taps may be written under it: the constant columns are rewritten with the same values and tap8 meets a
// zero row of B.
__device__ __forceinline__ void issue_mma1_tmem(uint32_t d_tmem, uint32_t y_tmem, uint32_t wb_addr, uint32_t t0_addr) {
    constexpr uint32_t idesc = tc::idesc_tf32(128, 64, 0, 0);
    const uint32_t b_lo = tc::desc_lo(wb_addr, 16), hi = tc::desc_hi(1024, 2);
    tc::umma_tf32_elect(d_tmem, tc::desc_lo(t0_addr, 16) + 2, hi, tc::desc_lo(t0_addr + 8192, 16) + 4 + 2, hi, idesc, 0u);
#pragma unroll
    for (int j = 0; j < 8; ++j)
        tc::umma_tf32_tmem_a_elect(d_tmem, y_tmem + 8 * j, b_lo + (((j >> 2) * 8192 + (j & 3) * 32) >> 4), hi, idesc, 1u);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    tc::tmem_ld16(taddr, v);
    tc::tmem_ld16(taddr + 16, v + 16);
    tc::tmem_ld_wait();
}

// a[i] = tanh(y[i] / 2) = 2 sigmoid(y[i]) - 1 for 32 values of the pre-scaled ys (see kGateScale)
__device__ __forceinline__ void tanh32(const float (&y)[32], float (&a)[32]) {
#if DCASE_GATE_TANH
#pragma unroll
    for (int i = 0; i < 32; ++i) asm("tanh.approx.f32 %0, %1;" : "=f"(a[i]) : "f"(y[i]));
#else
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = ex2_ftz(y[i]);
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = fmaf(2.f, rcp_ftz(1.f + a[i]), -1.f);
#endif
}

__device__ __forceinline__ void require_aligned_smem(const void* p) {
    if (tc::smem_u32(p) & 1023u) __trap();      // the operand swizzles assume 1024-byte aligned regions
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
// Software pipeline per CTA (4 CTAs / SM), two block barriers per tile:
//   top      MMA0 of this tile has completed (it was issued during the previous tile): MMA1 is issued at once
//   phase P  the PREVIOUS tile's z is average-pooled from shared memory on the CUDA cores (thread = one window x four
//            channels: 8 conflict-free 16-byte loads, one coalesced 16-byte store) -- under MMA1.  Round 2 replaced the
//            pooling MMA (16 x M64 N16 K8 + commit, ~900 tensor-pipe cycles per tile and a third round trip, whose result
//            also blocked the lin columns of tensor memory) by these ~40 instructions per thread
//   phase D  operand rows of the NEXT tile (taps / dropout bits), x rows of the tile after it prefetched into registers,
//            a = tanh(y / 2) from tensor memory
//   barrier A, then MMA0 of the next tile;  phase F  z = u (1 + a) -> shared memory;  barrier C
// z is held as fp16 (round to nearest: 11 significant bits, one more than the tf32 operand the pooling MMA used to read)
// and summed in fp32: 16 KB instead of 32 KB, which together with 63 registers per thread lets FOUR CTAs share an SM.
// smem (1024-B aligned, all dynamic): Wb 16 KB | T0 16 KB | Z 16 KB (128 rows x 64 fp16, 16-byte chunks swizzled by row) |
//                                     xs[4][66] | keep_lo[128] | 2 mbarriers | tmem base
// TMEM (128 columns, 4 CTAs = all 512): [0,64) y (the next tile's y from barrier A on), [64,128) u = (lin + b) / 2
constexpr int kFwdWb = 0, kFwdT0 = 16384, kFwdA = 32768, kFwdMisc = 49152;
constexpr int kFwdSmemBytes = kFwdMisc + 4 * 66 * 4 + 128 * 4 + 2 * 8 + 8 + 16;
constexpr int kFwdThreads = 256;
constexpr int kFwdCtasPerSm = 4;

__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float2 unpack_half2(uint32_t v) {
    float2 r;
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(r.x), "=f"(r.y) : "r"(v));
    return r;
}
// z row of pixel p: 128 bytes, 16-byte chunk c (8 channels) stored at chunk c ^ (p & 7)
__device__ __forceinline__ uint32_t zrow_base(uint32_t region, int p) { return region + (uint32_t)(p * 128); }

// average pool of one tile from its z rows in shared memory: thread = (window w, channels 4 c4 .. 4 c4 + 3)
__device__ __forceinline__ void pool_tile(uint32_t z_region, int tid, float scale, float* __restrict__ out_tile,
                                          uint2* __restrict__ out_tile_h) {
    const int w = tid >> 4, c4 = tid & 15;
    uint64_t acc01 = pk(0.f, 0.f), acc23 = acc01;     // packed fp32 sums (FADD2: half the adds)
    const uint64_t one2 = pk(1.f, 1.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int p = 4 * w + j;                      // frame 0; frame 1 is 64 rows = 8192 bytes further (same p & 7)
        const uint32_t addr = z_region + (uint32_t)(p * 128 + ((((c4 >> 1) ^ p) & 7) << 4) + (c4 & 1) * 8);
#pragma unroll
        for (int tr = 0; tr < 2; ++tr) {
            uint32_t v0, v1;
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v0), "=r"(v1) : "r"(addr + tr * 8192) : "memory");
            const float2 a = unpack_half2(v0), b = unpack_half2(v1);
            acc01 = fma2(pk(a.x, a.y), one2, acc01);
            acc23 = fma2(pk(b.x, b.y), one2, acc23);
        }
    }
    float4 acc;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(acc01));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.z), "=f"(acc.w) : "l"(acc23));
    const float4 o = make_float4(scale * acc.x, scale * acc.y, scale * acc.z, scale * acc.w);
    if (out_tile)                                          // operand of the weight-gradient MMAs: rounded to tf32
        *reinterpret_cast<float4*>(out_tile + w * 64 + 4 * c4) =
            make_float4(tf32_round_fast(o.x), tf32_round_fast(o.y), tf32_round_fast(o.z), tf32_round_fast(o.w));
    if (out_tile_h)                                        // operand of conv1's forward MMAs: fp16
        out_tile_h[w * 16 + c4] = make_uint2(pack_half2(o.x, o.y), pack_half2(o.z, o.w));
}

__global__ void __launch_bounds__(kFwdThreads, 4)
cnn0_fwd_kernel(Cnn0Args a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    require_aligned_smem(smem);
    unsigned char* Wb = smem + kFwdWb;
    unsigned char* T0 = smem + kFwdT0;
    float* xs = reinterpret_cast<float*>(smem + kFwdMisc);
    uint32_t* keep_lo = reinterpret_cast<uint32_t*>(xs + 4 * 66);
    uint64_t* bars = reinterpret_cast<uint64_t*>(keep_lo + 128);     // [0] MMA0, [1] MMA1
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 2);
    int* tq = reinterpret_cast<int*>(tmem_base_s + 1);               // [3] tile indices claimed by thread 0 (dynamic schedule)
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);       // warp-uniform for the compiler
    const int row = tid & 127, half = tid >> 7;

    stage_wg(a.glu_w, Wb, tid, kFwdThreads);
    stage_w0(a.fold0, T0, tid, kFwdThreads);
    stage_bias(a.glu_b, T0, tid, kFwdThreads);
    if (tid == 0) { tc::mbar_init(&bars[0], 1); tc::mbar_init(&bars[1], 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(tmem_base_s, 128);
    uint64_t seed = a.drop.seed; uint32_t step = a.drop.step;
    if (a.drop.sc) { seed = a.drop.sc->seed; step = a.drop.sc->step; }
    // Tiles are handed out DYNAMICALLY: a CTA's first tile is its block index, every further one comes from a global
    // counter.  The block scheduler needs ~16 us to start the grid's 592 CTAs (globaltimer: the fourth CTA of an SM starts
    // 11-16 us after the first) and the other model's kernels free SMs at their own pace, so with a static split the CTAs
    // that started last set the kernel's duration while the early ones had finished 30 us before.  The pipeline looks two
    // tiles ahead (operand rows of `nxt`, prefetched x rows of `nxt2`), so thread 0 claims the tile after those at the top of
    // every iteration and publishes it through shared memory before barrier A.
    const int n_tiles = (int)((long long)a.B * a.T / 2);
    int cur = blockIdx.x, nxt, nxt2;
    if (tid == 0) {
        tq[0] = (int)gridDim.x + (int)atomicAdd(a.tile_ctr, 1u);
        tq[1] = (int)gridDim.x + (int)atomicAdd(a.tile_ctr, 1u);
    }
    const uint32_t wb_a = tc::smem_u32(Wb), t0_a = tc::smem_u32(T0), a_a = tc::smem_u32(smem + kFwdA);
    const uint32_t t0_rb = krow_base(t0_a, row), z_rb = zrow_base(a_a, row);
    const bool drop = a.drop.enabled != 0;
    float* const out_f = a.out;
    uint2* const out_h = reinterpret_cast<uint2*>(a.out_h);

    // prologue: operand rows of the first tile and its MMA0; x rows of the second tile in shared memory, of the third in
    // registers
    uint32_t keep_next = 0xffffffffu;       // keep bits (this thread's 32 channels) of the tile whose MMA0 is in flight
    TilePos pos;                            // position of the tile whose x rows are fetched
    pos.init(cur, a.T);
    {
        XsRegs x0 = xs_prefetch(a.x, pos, a.T, tid);
        xs_commit(x0, xs, tid);
    }
    __syncthreads();
    const uint32_t tmem = *tmem_base_s;
    nxt = tq[0];
    nxt2 = tq[1];
    XsRegs xr{0.f, 0.f};
    if (nxt < n_tiles) { pos.init(nxt, a.T); xr = xs_prefetch(a.x, pos, a.T, tid); }
    if (half == 0) write_taps(xs, row, t0_rb);
    else if (drop) {
        const uint4 r = philox4x32_10((uint64_t)((long long)cur * kTile + row), a.drop.stream, step, seed);
        keep_lo[row] = r.x;
        keep_next = r.y;
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    if (warp == 0) { issue_mma0(tmem, t0_a); tc::umma_commit_elect(&bars[0]); }
    if (drop && half == 0) keep_next = keep_lo[row];
    xs_commit(xr, xs, tid);                 // the second tile's rows (the first tile's have been consumed)
    __syncthreads();

    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t ph0 = 0, ph1 = 0;
    const float pool_scale = drop ? 0.25f : 0.125f;      // 1/8 window, x2 inverted dropout
    int prev = -1;

#ifdef DCASE_CNN0_TIMING
    long long tm[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tt;
    int n_done = 0;
    const long long t_begin = clock64();
#define TICK() tt = clock64()
#define TOCK(i) tm[i] += clock64() - tt
#else
#define TICK()
#define TOCK(i)
#endif
    while (cur < n_tiles) {
        const bool has_next = nxt < n_tiles;
        const uint32_t keep = keep_next;
        TICK();
        tc::mbar_wait(&bars[0], ph0);            // y of `cur` is in TMEM; T0 may be rewritten
        ph0 ^= 1;
        tc::fence_after_sync();
        if (warp == 0) {                          // MMA1 at once: y is read from tensor memory; the u columns were drained by
            issue_mma1_tmem(tmem + 64, tmem, wb_a, t0_a);      // every thread before barrier C of the previous tile
            tc::umma_commit_elect(&bars[1]);
        }
        if (tid == 0) tq[2] = nxt2 < n_tiles ? (int)gridDim.x + (int)atomicAdd(a.tile_ctr, 1u) : n_tiles;   // the tile after `nxt2`
        TOCK(0);
        TICK();
        if (prev >= 0) pool_tile(a_a, tid, pool_scale, out_f ? out_f + (long long)prev * 16 * 64 : nullptr, out_h ? out_h + (long long)prev * 16 * 16 : nullptr);     // z of `prev`: complete since barrier C
        TOCK(1);
        TICK();
        uint32_t keep_hi_next = 0xffffffffu;
        if (has_next) {
            if (half == 0) write_taps(xs, row, t0_rb);
            else if (drop) {
                const uint4 r = philox4x32_10((uint64_t)((long long)nxt * kTile + row), a.drop.stream, step, seed);
                keep_lo[row] = r.x;
                keep_hi_next = r.y;
            }
            if (nxt2 < n_tiles) { pos.init(nxt2, a.T); xr = xs_prefetch(a.x, pos, a.T, tid); }
        }
        TOCK(2);
        TICK();
        float g[32];                              // a = tanh(y / 2); z = u (1 + a)
        {
            float y[32];
            tmem_ld32(tmem + lane_base + 32 * half, y);
            tc::fence_before_sync();
            tanh32(y, g);
        }
        tc::fence_proxy_async();                  // the taps of `nxt`
        TOCK(3);
        TICK();
        __syncthreads();                          // A: everybody holds its y, has pooled `prev` and written the rows of `nxt`
        const int nxt3 = tq[2];
        TOCK(4);
        TICK();
        if (warp == 0 && has_next) {
            tc::mbar_wait(&bars[1], ph1);         // MMA1 has read y (long done: it ran under phases P and D); an MMA that
            tc::fence_after_sync();               // overwrites a TMEM A operand must not be queued behind its reader
            issue_mma0(tmem, t0_a);
            tc::umma_commit_elect(&bars[0]);
        }
        if (drop) {
#pragma unroll
            for (int i = 0; i < 32; ++i) g[i] = (keep & (1u << i)) ? g[i] : -1.f;     // u * (-1) + u = 0 exactly
            keep_next = half ? keep_hi_next : keep_lo[row];
        }
        tc::mbar_wait(&bars[1], ph1);
        TOCK(5);
        TICK();
        ph1 ^= 1;
        tc::fence_after_sync();
        {
            float u[32];
            tmem_ld32(tmem + 64 + lane_base + 32 * half, u);
            tc::fence_before_sync();
#pragma unroll
            for (int q = 0; q < 4; ++q) {         // 8 channels = one 16-byte chunk of fp16
                uint32_t h[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int i = 8 * q + 2 * e;
                    const uint64_t uu = pk(u[i], u[i + 1]);
                    float z0, z1;
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(z0), "=f"(z1) : "l"(fma2(uu, pk(g[i], g[i + 1]), uu)));
                    h[e] = pack_half2(z0, z1);
                }
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(z_rb + (uint32_t)((((4 * half + q) ^ row) & 7) << 4)),
                             "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
            }
        }
        if (has_next && nxt2 < n_tiles) xs_commit(xr, xs, tid);      // rows of `nxt2` (those of `nxt` were consumed before A)
        TOCK(6);
        TICK();
        __syncthreads();                          // C: z of `cur` is complete, its u columns are drained
        TOCK(7);
#ifdef DCASE_CNN0_TIMING
        ++n_done;
#endif
        prev = cur;
        cur = nxt; nxt = nxt2; nxt2 = nxt3;
    }
#ifdef DCASE_CNN0_TIMING
    if (blockIdx.x == 7 && (tid == 0 || tid == 160))
        printf("cnn0_fwd tid %d: total %lld tiles %d | wait y + mma1 %lld | pool %lld | taps %lld | phase d %lld | sync A %lld | mma0+mask+wait %lld | phase f %lld | sync C %lld\n",
               tid, clock64() - t_begin, n_done, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], tm[6], tm[7]);
#endif
    if (prev >= 0) pool_tile(a_a, tid, pool_scale, out_f ? out_f + (long long)prev * 16 * 64 : nullptr, out_h ? out_h + (long long)prev * 16 * 16 : nullptr);
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

// ---------------------------------------------------------------------------------------------
// backward (recomputes the forward per tile)
// ---------------------------------------------------------------------------------------------
// One CTA per SM, 512 threads = two independent groups of 256 (named barriers), each streaming its own tiles with
// the same software pipeline as the forward (MMA0 of the next tile rides with MMA1 of the current one), so one
// group's CUDA-core phases overlap the other's tensor-core round trips.
// smem (1024-B aligned, all dynamic): Wb 16 KB | per group: T0 16 KB | E 16 KB (MN-major [taps | 1]) | DL 32 KB
//                        (first y, K-major) | D2 32 KB (DL and D2 contiguous: one M = 128 MN-major A operand) | misc
// TMEM (512 columns): group g at 256 g: [0,64) y, [64,128) lin, [128,144) accumulator {U | S2}[128][16]
constexpr int kBwdThreads = 512;
constexpr int kBwdGroupBytes = 16384 + 16384 + 32768 + 32768;
constexpr int kBwdMiscOff = 16384 + 2 * kBwdGroupBytes;
constexpr int kBwdMiscGroupFloats = 4 * 66 + 128;          // xs | keep_lo
constexpr int kBwdSmemBytes = kBwdMiscOff + (64 + 2 * kBwdMiscGroupFloats) * 4 + 6 * 8 + 8;

__device__ __forceinline__ void cnn0_bwd_finalize_body(const Cnn0FinArgs& f, const float* __restrict__ fold0, const float* us, float* sm,
                                                       int tid, int nt);

__global__ void __launch_bounds__(kBwdThreads, 1)
cnn0_bwd_kernel(Cnn0Args a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    require_aligned_smem(smem);
    const int tid = threadIdx.x, grp = tid >> 8, gt = tid & 255;
    const int lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);       // warp-uniform for the compiler
    const bool issuer = (warp & 7) == 0;                           // warp 0 of each group issues its MMAs
    const int row = gt & 127, half = gt >> 7;
    unsigned char* Wb = smem;
    unsigned char* gbase = smem + 16384 + grp * kBwdGroupBytes;
    unsigned char* T0 = gbase;
    unsigned char* E = gbase + 16384;
    float* bg = reinterpret_cast<float*>(smem + kBwdMiscOff);
    float* xs = bg + 64 + grp * kBwdMiscGroupFloats;
    uint32_t* keep_lo = reinterpret_cast<uint32_t*>(xs + 4 * 66);
    uint64_t* bars_all = reinterpret_cast<uint64_t*>(bg + 64 + 2 * kBwdMiscGroupFloats);
    uint64_t* bars = bars_all + 3 * grp;                 // [0] MMA0, [1] MMA1, [2] MMA3 (accumulate)
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars_all + 6);

    stage_wg(a.glu_w, Wb, tid, kBwdThreads);
    stage_w0(a.fold0, T0, gt, 256);
    stage_bias(a.glu_b, T0, gt, 256);
    for (int r = gt; r < 128; r += 256)                    // E chunk 3 (columns 12..15) stays zero
        *reinterpret_cast<float4*>(E + tc::sw128b32_chunk(r, 3)) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gt == 0) { tc::mbar_init(&bars[0], 1); tc::mbar_init(&bars[1], 1); tc::mbar_init(&bars[2], 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(tmem_base_s, 512);
    uint64_t seed = a.drop.seed; uint32_t step = a.drop.step;
    if (a.drop.sc) { seed = a.drop.sc->seed; step = a.drop.sc->step; }
    const long long n_tiles = (long long)a.B * a.T / 2;
    const long long stride = (long long)gridDim.x * 2;
    long long cur = (long long)blockIdx.x * 2 + grp;
    const uint32_t wb_a = tc::smem_u32(Wb), t0_a = tc::smem_u32(T0), e_a = tc::smem_u32(E), dl_a = tc::smem_u32(gbase + 32768);
    const uint32_t t0_rb = krow_base(t0_a, row), e_rb = mnrow_base(e_a, row), y_rb = krow_base(dl_a, row),
                   dl_rb = mnrow_base(dl_a, row);       // D2 rows: dl_rb + 32768 (chunk_addr's block stride continues)
    const bool drop = a.drop.enabled != 0;
    const bool active = cur < n_tiles;
    const int bar_id = 1 + grp;

    uint32_t keep_next = 0xffffffffu;
    TilePos pos;
    pos.init(active ? cur : 0, a.T);
    const int pos_step = (int)(2 * stride);
    if (active) {
        XsRegs xr0 = xs_prefetch(a.x, pos, a.T, gt);
        xs_commit(xr0, xs, gt);
    }
    __syncthreads();
    const uint32_t tmem = *tmem_base_s + 256u * grp;
    if (active) {
        if (half == 0) write_taps(xs, row, t0_rb);
        else if (drop) {
            const uint4 r = philox4x32_10((uint64_t)(cur * kTile + row), a.drop.stream, step, seed);
            keep_lo[row] = r.x;
            keep_next = r.y;
        }
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    if (active && issuer) { issue_mma0(tmem, t0_a); tc::umma_commit_elect(&bars[0]); }
    if (active && drop && half == 0) keep_next = keep_lo[row];
    pos.advance(pos_step, a.T);
    XsRegs xr = cur + stride < n_tiles ? xs_prefetch(a.x, pos, a.T, gt) : XsRegs{0.f, 0.f};

    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t ph0 = 0, ph1 = 0, ph2 = 0;
    bool pending = false, first = true;
    const float dz_scale = drop ? 0.25f : 0.125f;

    for (; cur < n_tiles; cur += stride) {
        const long long nxt = cur + stride;
        const bool has_next = nxt < n_tiles;
        const uint32_t keep = keep_next;
        tc::mbar_wait(&bars[0], ph0);            // y of `cur` is in TMEM; T0 may be read / rewritten
        ph0 ^= 1;
        tc::fence_after_sync();
        if (issuer) {                             // MMA1 straight away: y is read from tensor memory, and the lin columns
            issue_mma1_tmem(tmem + 64, tmem, wb_a, t0_a);   // were drained by every thread before the previous tile's last barrier
            tc::umma_commit_elect(&bars[1]);
        }
        // gate phase first: it only needs y, and the accumulating MMA of the previous tile finishes under it
        float g[32];
        {
            float y[32];
            tmem_ld32(tmem + lane_base + 32 * half, y);
            tc::fence_before_sync();
            tanh32(y, g);                         // a = tanh(y / 2)
        }
        if (pending) { tc::mbar_wait(&bars[2], ph2); ph2 ^= 1; pending = false; }   // E / DL / D2 free again
        tc::fence_after_sync();
        if (half == 0) {                          // E (MN-major copy of this tile's operand rows) for MMA3
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4 v = lds128(chunk_addr(t0_rb, c));
                sts128(chunk_addr(e_rb, c), v.x, v.y, v.z, v.w);
            }
        }
        uint32_t keep_hi_next = 0xffffffffu;
        if (has_next) {
            xs_commit(xr, xs, gt);
            bar_sync_named(bar_id, 256);
            if (half == 0) write_taps(xs, row, t0_rb);
            else if (drop) {
                const uint4 r = philox4x32_10((uint64_t)(nxt * kTile + row), a.drop.stream, step, seed);
                keep_lo[row] = r.x;
                keep_hi_next = r.y;
            }
            pos.advance(pos_step, a.T);
            if (nxt + stride < n_tiles) xr = xs_prefetch(a.x, pos, a.T, gt);
        }
        tc::fence_proxy_async();
        bar_sync_named(bar_id, 256);              // everybody holds its y: MMA0 of `nxt` may overwrite the columns
        if (issuer && has_next) {
            tc::mbar_wait(&bars[1], ph1);         // MMA1 has read y: see the forward kernel
            tc::fence_after_sync();
            issue_mma0(tmem, t0_a);
            tc::umma_commit_elect(&bars[0]);
        }
        // gradient of the pooled output for this pixel's window (raw: the pool / dropout scale is applied to the accumulator
        // at the end), dropout mask applied; overlaps MMA1
        float dz[32];
        {
            const int f = row & 63;
            const float4* dsrc = reinterpret_cast<const float4*>(a.d_out + (cur * 16 + (f >> 2)) * 64) + 8 * half;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 d = __ldg(dsrc + q);
                dz[4 * q] = d.x; dz[4 * q + 1] = d.y; dz[4 * q + 2] = d.z; dz[4 * q + 3] = d.w;
            }
            if (drop) {
#pragma unroll
                for (int i = 0; i < 32; ++i) dz[i] = (keep & (1u << i)) ? dz[i] : 0.f;
                keep_next = half ? keep_hi_next : keep_lo[row];
            }
        }
        tc::mbar_wait(&bars[1], ph1);
        ph1 ^= 1;
        tc::fence_after_sync();
        {
            float u[32];                                  // (lin + b) / 2
            tmem_ld32(tmem + 64 + lane_base + 32 * half, u);
            tc::fence_before_sync();
            const uint64_t m1 = pk(-1.f, -1.f);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                uint64_t dl[2], d2[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int i = 4 * q + 2 * e;
                    const uint64_t A = pk(g[i], g[i + 1]), D = pk(dz[i], dz[i + 1]), U = pk(u[i], u[i + 1]);
                    dl[e] = fma2(D, A, D);                        // 2 x grad wrt lin:  dz (1 + a)
                    d2[e] = mul2(mul2(D, U), fma2(A, A, m1));     // -2 x grad wrt y through the gate:  dz u (a a - 1)
                }
                const uint32_t addr = chunk_addr(dl_rb, 8 * half + q);
                sts128_2(addr, dl[0], dl[1]);                     // overwrites y (MMA1 has completed)
                sts128_2(addr + 32768u, d2[0], d2[1]);
            }
        }
        tc::fence_proxy_async();
        bar_sync_named(bar_id, 256);
        if (issuer) {                      // MMA3: {U | S2}[m][j] += sum_p [DL | D2][p][m] E[p][j]
            tc::fence_after_sync();
            constexpr uint32_t idesc = tc::idesc_tf32(128, 16, 1, 1);
            const uint32_t d_lo = tc::desc_lo(dl_a, 16384), e_lo = tc::desc_lo(e_a, 16384), mn_hi = tc::desc_hi(512, 1);
            const uint32_t acc1 = first ? 0u : 1u;
#pragma unroll
            for (int j = 0; j < 16; ++j)
                tc::umma_tf32_elect(tmem + 128, d_lo + (j * 1024 >> 4), mn_hi, e_lo + (j * 1024 >> 4), mn_hi, idesc, j > 0 ? 1u : acc1);
            tc::umma_commit_elect(&bars[2]);
        }
        pending = true;
        first = false;
    }
    if (pending) { tc::mbar_wait(&bars[2], ph2); ph2 ^= 1; }
    tc::fence_after_sync();
    if (!first && (warp & 7) < 4) {        // warps 0..3 of each group: accumulator rows 32 q .. 32 q + 31
        float v[16];
        tc::tmem_ld16(tmem + 128 + lane_base, v);
        tc::tmem_ld_wait();
        const int m = 32 * (warp & 3) + lane;
        const float sc = ((warp & 2) ? -0.5f : 0.5f) * dz_scale * kTruncComp;     // rows < 64: U (from 2 DL), rows >= 64: S2 (from -2 D2)
#pragma unroll
        for (int j = 0; j < 10; ++j) atomicAdd(a.us + m * 16 + j, sc * v[j]);
    }
    tc::fence_before_sync();
    __threadfence();                       // this CTA's accumulator atomics are visible before its ticket
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(*tmem_base_s, 512);
    if (a.fin.enabled) {                   // the last CTA to finish turns the accumulator into the block's parameter gradients
        __shared__ int is_last;
        if (tid == 0) is_last = atomicAdd(reinterpret_cast<unsigned int*>(a.us + 10), 1u) == gridDim.x - 1;
        __syncthreads();
        if (is_last) {
            __threadfence();
            cnn0_bwd_finalize_body(a.fin, a.fold0, a.us, reinterpret_cast<float*>(smem + 16384), tid, kBwdThreads);   // the operand tiles are dead
        }
    }
}

// {U, S2} -> GLU parameter gradients and S = Wg^T U + S2; then the BatchNorm / conv0 parameter gradients from S and the
// tap moments, task (c, k) owning tap k of channel c (the fp64 terms are spread over nine tasks per channel: B200's fp64
// rate is low and this is the last link of the backward chain before Adam).  `sm`: 6,080 floats of shared memory.
__device__ __forceinline__ void cnn0_bwd_finalize_body(const Cnn0FinArgs& f, const float* __restrict__ fold0,
                                                       const float* us, float* sm, int tid, int nt) {
    float (*U)[10] = reinterpret_cast<float (*)[10]>(sm);
    float (*S)[10] = reinterpret_cast<float (*)[10]>(sm + 640);
    float (*W0e)[10] = reinterpret_cast<float (*)[10]>(sm + 1280);
    float* Wg = sm + 1920;                               // [n][c], pitch 65: the S pass walks a column
    for (int i = tid; i < 640; i += nt) {
        const int r = i / 10, j = i - r * 10;
        U[r][j] = __ldcg(us + r * 16 + j);               // from L2: other CTAs' atomics
        W0e[r][j] = j < 9 ? fold0[kFold0Wf + j * 64 + r] : fold0[kFold0Bf + r];
    }
    for (int i0 = 0; i0 < 4096; i0 += 8 * nt) {         // eight loads in flight per thread (this runs in ONE CTA, on the chain)
        float w8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { const int i = i0 + tid + k * nt; w8[k] = i < 4096 ? __ldg(f.glu_w_raw + i) : 0.f; }
#pragma unroll
        for (int k = 0; k < 8; ++k) { const int i = i0 + tid + k * nt; if (i < 4096) Wg[(i >> 6) * 65 + (i & 63)] = w8[k]; }
    }
    __syncthreads();
    for (int i = tid; i < 4096; i += nt) {              // dWg[n][k] = sum_j U[n][j] W0e[k][j]
        const int n = i >> 6, k = i & 63;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 10; ++j) acc = fmaf(U[n][j], W0e[k][j], acc);
        f.g_glu_w[i] = f.pgs * acc;
    }
    if (tid < 64) f.g_glu_b[tid] = f.pgs * U[tid][9];
    for (int task = tid; task < 640; task += nt) {      // S[c][j] = sum_n Wg[n][c] U[n][j] + S2[c][j]
        const int c = task / 10, j = task - c * 10;
        float acc = __ldcg(us + (64 + c) * 16 + j);
#pragma unroll 8
        for (int n = 0; n < 64; ++n) acc = fmaf(Wg[n * 65 + c], U[n][j], acc);
        S[c][j] = acc;
    }
    __syncthreads();
    for (int task = tid; task < 640; task += nt) {
        const int c = task / 10, k = task - c * 10;     // k = 9: the channel's gamma / beta / bias task
        const double n = (double)f.n_pix;
        const double mean = fold0[kFold0Mean + c], invstd = fold0[kFold0Invstd + c], av = fold0[kFold0A + c];
        const double S1 = S[c][9];                      // sum dY
        double wG = 0.0;
        for (int l = 0; l < 9; ++l) wG += (double)f.w[c * 9 + l] * (double)S[c][l];   // sum dY * (conv output - bias)
        const double bm = (double)f.b[c] - mean;
        const double S2 = invstd * (wG + bm * S1);      // sum dY * xhat
        if (k == 9) {
            f.g_gamma[c] = f.pgs * (float)S2;
            f.g_beta[c] = f.pgs * (float)S1;
            f.g_b[c] = 0.f;                             // BN cancels the conv bias
        } else {
            double sxx = 0.0;                           // sum_p xhat_c * x_k
            for (int l = 0; l < 9; ++l) {
                const int lo = l <= k ? l : k, hi = l <= k ? k : l;
                sxx += (double)f.w[c * 9 + l] * f.mom[9 + lo * 9 - (lo * (lo - 1)) / 2 + (hi - lo)];
            }
            sxx = invstd * (sxx + bm * f.mom[k]);
            f.g_w[c * 9 + k] = f.pgs * (float)(av * ((double)S[c][k] - (S1 / n) * f.mom[k] - (S2 / n) * sxx));
        }
    }
}

// the same as a kernel of its own (SyncBN: the accumulator crosses the ranks between cnn0_bwd_kernel and this)
constexpr int kBwdFinThreads = 640;
__global__ void __launch_bounds__(kBwdFinThreads)
cnn0_bwd_finalize_kernel(Cnn0FinArgs f, const float* __restrict__ fold0, const float* __restrict__ us) {
    __shared__ float sm[6080];
    cnn0_bwd_finalize_body(f, fold0, us, sm, threadIdx.x, kBwdFinThreads);
}

}  // namespace

int cnn0_kernels_init() {
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(cnn0_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
    DCASE_CUDA_CHECK(cudaFuncSetAttribute(cnn0_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmemBytes));
    return DCASE_OK;
}

int launch_cnn0_fwd(const float* x, int B, int T, const float* fold0, const float* glu_w, const float* glu_b,
                    DropoutCfg drop, float* out, void* out_h, unsigned int* tile_ctr, int num_sms, cudaStream_t s) {
    DCASE_PROF("cnn0_fused_fwd", s);
    DCASE_REQUIRE((long long)B * T / 2 < (1ll << 30), "batch too large");
    DCASE_CUDA_CHECK(cudaMemsetAsync(tile_ctr, 0, sizeof(unsigned int), s));
    Cnn0Args a{};
    a.x = x; a.B = B; a.T = T; a.fold0 = fold0; a.glu_w = glu_w; a.glu_b = glu_b; a.drop = drop; a.out = out; a.out_h = out_h;
    a.tile_ctr = tile_ctr;
    const long long n_tiles = (long long)B * T / 2;
    const long long grid = n_tiles < (long long)kFwdCtasPerSm * num_sms ? n_tiles : (long long)kFwdCtasPerSm * num_sms;
    cnn0_fwd_kernel<<<(int)grid, kFwdThreads, kFwdSmemBytes, s>>>(a);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int launch_cnn0_bwd(const float* x, int B, int T, const float* fold0, const float* glu_w, const float* glu_b,
                    DropoutCfg drop, const float* d_out, float* us, const Cnn0BwdFinalize* fin, int num_sms, cudaStream_t s) {
    DCASE_PROF("cnn0_fused_bwd", s);
    Cnn0Args a{};
    a.x = x; a.B = B; a.T = T; a.fold0 = fold0; a.glu_w = glu_w; a.glu_b = glu_b; a.drop = drop; a.d_out = d_out; a.us = us;
    if (fin)
        a.fin = Cnn0FinArgs{1, fin->mom, fin->n_pix, fin->conv_w, fin->conv_b, glu_w, fin->param_grad_scale, fin->g_conv_w,
                            fin->g_conv_b, fin->g_gamma, fin->g_beta, fin->g_glu_w, fin->g_glu_b};
    const long long n_tiles = (long long)B * T / 2;
    const long long grid = (n_tiles + 1) / 2 < num_sms ? (n_tiles + 1) / 2 : num_sms;
    cnn0_bwd_kernel<<<(int)grid, kBwdThreads, kBwdSmemBytes, s>>>(a);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}

int launch_cnn0_bwd_finalize(const Cnn0BwdFinalize& fin, const float* fold0, const float* glu_w, const float* us, cudaStream_t s) {
    DCASE_PROF("cnn0_bwd_finalize", s);
    const Cnn0FinArgs f{1, fin.mom, fin.n_pix, fin.conv_w, fin.conv_b, glu_w, fin.param_grad_scale, fin.g_conv_w, fin.g_conv_b,
                        fin.g_gamma, fin.g_beta, fin.g_glu_w, fin.g_glu_b};
    cnn0_bwd_finalize_kernel<<<1, kBwdFinThreads, 0, s>>>(f, fold0, us);
    DCASE_LAUNCH_CHECK();
    return DCASE_OK;
}
