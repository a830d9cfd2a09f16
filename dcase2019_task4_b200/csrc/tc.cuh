// tcgen05 / TMEM / mbarrier primitives for sm_100a (inline PTX), plus the shared-memory operand layout used
// by every tensor-core kernel here.
//
// Operand layout ("SW128 block"): [rows][32 fp32] = rows x 128 B, 8-row swizzle atoms of 1024 B, the 16-byte
// chunk c of row r stored at chunk (c ^ (r & 7)).  One block read as
//   * K-major operand  : rows = M or N index, the 32 floats = 32 consecutive K   (SBO = 1024 B between 8-row groups)
//   * MN-major operand : rows = K index,      the 32 floats = 32 consecutive M/N (SBO = 1024 B between 8-K groups,
//                        LBO = byte distance to the block holding the next 32 M/N values)
// kind::tf32 consumes K = 8 per instruction: K-major advances the start address by 32 B inside the 128-B row,
// MN-major by one 8-row atom (1024 B).  fp32 values are stored unrounded (the tensor core reads the top 19 bits),
// so epilogues can re-read exact values from the same tile.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row r, k in [0,32)) inside an SW128 block
__device__ __forceinline__ uint32_t sw128_off(int r, int k) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 2) ^ r) & 7) << 4) + (k & 3) * 4);
}
// byte offset of 16-byte chunk c (0..7) of row r
__device__ __forceinline__ uint32_t sw128_chunk(int r, int c) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + (((c ^ r) & 7) << 4));
}

// kind::tf32 TRUNCATES fp32 operands to 19 bits; truncation shrinks every product by ~2^-11 and the bias compounds
// layer after layer (measured: -0.4 % on the layer-1 batch variance).  Every value that feeds an MMA is therefore
// rounded to nearest tf32 when it is produced (unbiased), by the thread that writes the operand tile / tensor.
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ float4 tf32_rn4(float4 v) { return make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w)); }

// ---- mbarrier ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM -------------------------------------------------------------------------------------------
// warp-collective; ncols power of two >= 32; writes the TMEM base address to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread = lane = accumulator row)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {   // 32 lanes x 4 consecutive fp32 columns
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// accumulator row `lane 32*warp + l`, 64 columns starting at `col`, into v[64]
__device__ __forceinline__ void tmem_ld_row64(uint32_t tmem_base, int warp, int col, float (&v)[64]) {
    const uint32_t a = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)col;
    tmem_ld16(a, v);
    tmem_ld16(a + 16, v + 16);
    tmem_ld16(a + 32, v + 32);
    tmem_ld16(a + 48, v + 48);
    tmem_ld_wait();
}

// ---- descriptors ------------------------------------------------------------------------------------
// shared-memory matrix descriptor, version 1 (sm_100); layout_type: 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    d |= (uint64_t)layout_type << 61;
    return d;
}
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return smem_desc(smem_addr, lbo_bytes, sbo_bytes, 2);
}
// MN-major fp32 operands must use SWIZZLE_128B_BASE32B: [rows = K][32 fp32], 4-row atoms of 512 B, the 32-byte
// chunk c of row r stored at chunk (c ^ (r & 3)).  Byte offset of 16-byte chunk c16 (0..7) of row r:
__device__ __forceinline__ uint32_t sw128b32_chunk(int r, int c16) {
    return (uint32_t)(r * 128 + ((((c16 >> 1) ^ r) & 3) << 5) + ((c16 & 1) << 4));
}
// instruction descriptor: kind::tf32, fp32 accumulate, dense; majors: 0 = K-major, 1 = MN-major
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4)                      // c_format = F32
           | (2u << 7) | (2u << 10)       // a_format = b_format = TF32
           | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16)
           | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::f16 with fp16 operands (a_format = b_format = 0), fp32 accumulate: K = 16 per instruction, i.e. twice the K of
// kind::tf32 for the same 32 operand bytes per row -- the shared-memory operand fetch that bounds N = 64 MMAs halves
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

// the same with bf16 operands (a_format = b_format = 1)
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return idesc_f16(M, N, a_mn_major, b_mn_major) | (1u << 7) | (1u << 10);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// ---- warp-uniform issue path ---------------------------------------------------------------------------
// The single-thread form above, called under `if (tid == 0)`, makes the compiler wrap every MMA in a convergence
// loop and rebuild both 64-bit descriptors (~16 instructions per MMA on the issuing thread).  The functions below
// are called by ALL lanes of the issuing warp with warp-uniform arguments: one lane is elected inside the asm, the
// descriptors are (lo, hi) pairs whose low word (start address >> 4 | LBO << 16) advances by a constant.
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
    return ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout_type) {
    return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout_type << 29);
}
// one MMA (K = 8): D[tmem] (+)= A * B, issued by one elected lane
__device__ __forceinline__ void umma_tf32_elect(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_f16_elect(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same with the A operand read from TENSOR MEMORY: [a_tmem] = 128 lanes (rows m) x K fp32 columns, i.e. exactly the
// layout an M = 128 accumulator has -- a GEMM can consume the previous GEMM's result without a shared-memory round
// trip (values are truncated to tf32 as they are read).  K = 8 per instruction: advance a_tmem by 8 columns.
__device__ __forceinline__ void umma_tf32_tmem_a_elect(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi,
                                                       uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        ".reg .b64 db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(smem_u32(bar))
        : "memory");
}

// arrive on the mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// D[128 x 64] = A[128 x 64] * B[64 x 64]^T  (both K-major, two 32-wide K blocks each)
//   a_base: two blocks of 128 rows (16 KB each, contiguous); b_base: two blocks of 64 rows (8 KB each)
__device__ __forceinline__ void umma_128x64x64_kmajor(uint32_t d_tmem, uint32_t a_base, uint32_t b_base, bool accumulate) {
    constexpr uint32_t idesc = idesc_tf32(128, 64, 0, 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t a = a_base + (j >> 2) * (128 * 128) + (j & 3) * 32;
        const uint32_t b = b_base + (j >> 2) * (64 * 128) + (j & 3) * 32;
        umma_tf32(d_tmem, smem_desc_sw128(a, 16, 1024), smem_desc_sw128(b, 16, 1024), idesc, (accumulate || j > 0) ? 1u : 0u);
    }
}

// D[64 x 64] (+)= sum over `rows` pixels of A[p][m] * B[p][n]  (both MN-major: rows = K = pixels)
//   a_base / b_base: two blocks (channels 0..31 | 32..63) of `rows` rows each; blk_bytes = rows * 128
__device__ __forceinline__ void umma_64x64_mnmajor(uint32_t d_tmem, uint32_t a_base, uint32_t b_base, int rows, bool accumulate) {
    constexpr uint32_t idesc = idesc_tf32(64, 64, 1, 1);
    const uint32_t blk = (uint32_t)rows * 128;
    for (int j = 0; j < rows / 8; ++j) {
        umma_tf32(d_tmem, smem_desc(a_base + j * 1024, blk, 512, 1), smem_desc(b_base + j * 1024, blk, 512, 1), idesc,
                  (accumulate || j > 0) ? 1u : 0u);
    }
}

}  // namespace tc
