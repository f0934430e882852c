// TMA (cp.async.bulk.tensor) helpers for the row-gather kernels: tensor maps over bf16
// [rows][64] shadow activations (box = one 128-byte row, SWIZZLE_128B -- gathered rows land
// in shared memory exactly in the UMMA K-major SW128 operand layout; layout facts pinned on
// a B200 by tools/probe_gather4.cu) and the tile::gather4 / 2-D tile loads.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_common.cuh"

namespace nlam {

// bf16 [rows][64] row-major, one-row box (for tile::gather4) -- host side, no GPU work
int make_row_map_bf16(CUtensorMap* map, const void* base, long long rows, int box_rows);

namespace tc {

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// four rows r0..r3 (64 bf16 columns from column c0) -> 4 consecutive 128-byte rows at dst
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int r0, int r1, int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}
// box of `box_rows` consecutive rows starting at row r0 -> dst (box_rows * 128 bytes)
__device__ __forceinline__ void tma_load_rows(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                              int r0) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(r0)
      : "memory");
}

}  // namespace tc
}  // namespace nlam
