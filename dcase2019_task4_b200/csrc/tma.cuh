// TMA (cp.async.bulk.tensor) helpers shared by the tensor-core kernels: host-side tensor-map construction through
// the driver entry point (no libcuda link dependency) and the device-side issue / mbarrier wrappers.
#pragma once
#include <cuda.h>            // CUtensorMap (types only)
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "tc.cuh"

int dcase_tma_init();        // conv_tc.cu: resolves cuTensorMapEncodeTiled once

// 4-D view [B][T_l][F][64] of a channels-last activation; box = [1][rows][pitch][32 channels] lands in shared memory as
// rows * pitch consecutive 128-byte rows in the tensor core's swizzled operand layout; out-of-range frames / mel bins
// (the zero padding of the convolution, tile tails) are zero-filled by the TMA unit.
int make_act_map(CUtensorMap* map, const float* base, int B, int T_l, int F, int box_rows, int pitch, CUtensorMapSwizzle swz);
// 2-D view [n_rows][64] of the same memory; box = [box_rows][32 channels]
int make_rows_map(CUtensorMap* map, const float* base, long long n_rows, int box_rows, CUtensorMapSwizzle swz);

// 2-D view of a row-major fp32 matrix [rows][cols] with row stride `ld` floats; box = [box_rows][box_cols]
int make_matrix_map(CUtensorMap* map, const float* base, long long cols, long long rows, long long ld, int box_cols,
                    int box_rows, CUtensorMapSwizzle swz);

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     tc::smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(tc::smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst_smem, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            tc::smem_u32(dst_smem)),
        "l"(map), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     tc::smem_u32(dst_smem)),
                 "l"(map), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
// shared -> global tile store (bulk async group of the issuing thread)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src_smem, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(tc::smem_u32(src_smem)),
                 "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ float4 ld_shared_v4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
