// Shared pieces of the bf16 tensor-core backward kernels (dgrad, wgrad, multi-context dgrad).
#pragma once
#include "rowmlp_tc.cuh"

namespace nlam {
namespace tc {

struct BGeo {
  int n1, n2, nmax;
  int k1, k2, ko;        // K of GEMM1 (pad 16), GEMM2 (= pad16(dh)), GEMM3 (= pad16(dout))
  int kb1, kb2, kbo;     // 64-wide blocks of z, of the hidden tile, of the dY tile
  int rb;                // z blocks gathered per round
  int tmem_cols, cY, cZ;
  uint32_t off_t, off_w1, off_w2, off_par, off_lnx, off_bar, smem_bytes;
  int total_tiles, tiles_per_batch;
  int need_dz;           // any source gradient requested
  int parts;             // 1 = bf16 operands, 2 = fp32 operands split into bf16 hi + lo tiles
  // scratch
  uint8_t* a_img;
  uint8_t* dy_img;
  uint8_t* dh_img;
  float* partial;        // wgrad: [w_slots][n_chunks][p_total] (matrix entries)
  float* vec_partial;    // dgrad: [d_slots][n_chunks][vec_len] = [db1 | db2 | dLNg | dLNb]
  int p_total, vec_len;
  // wgrad
  int w_tmem_cols, w_mchunks;
  uint32_t w_off_a, w_off_dy, w_off_dh, w_off_bar, w_smem_bytes;
};

// swizzled fp32 staging tile [128][n] (n multiple of 16): 16-byte chunk c4 of row r
__device__ __forceinline__ int stg_idx(int r, int c4, int n) {
  const int nc = n >> 2;
  const int m = (nc < 8 ? nc : 8) - 1;
  return r * n + ((c4 ^ (r & m)) << 2);
}

// Column sums over the 32 lanes of a warp of a 16-column chunk held one row per
// lane; lane l ends up with the total of column (l & 15).  16 shuffles.
__device__ __forceinline__ float warp_colsum16(const float (&v)[16], int lane) {
  float w8[8], w4[4], w2[2];
  const bool b3 = lane & 8, b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = b3 ? v[i] : v[i + 8];
    const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
    w8[i] = (b3 ? v[i + 8] : v[i]) + recv;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b2 ? w8[i] : w8[i + 4];
    const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
    w4[i] = (b2 ? w8[i + 4] : w8[i]) + recv;
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b1 ? w4[i] : w4[i + 2];
    const float recv = __shfl_xor_sync(0xffffffffu, send, 2);
    w2[i] = (b1 ? w4[i + 2] : w4[i]) + recv;
  }
  const float send = b0 ? w2[0] : w2[1];
  const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
  float w1 = (b0 ? w2[1] : w2[0]) + recv;
  w1 += __shfl_xor_sync(0xffffffffu, w1, 16);
  return w1;
}

// MN-major SW128 descriptor: tile stored [K rows][64 MN elements] (128-byte
// rows, same physical layout as the K-major tiles), 8-row K groups 1024 B apart
// (SBO), 64-element MN blocks `lbo_bytes` apart (LBO).
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         (64ull << 32) | (1ull << 46) | (2ull << 61);
}

template <int TNT = NT>
__device__ __forceinline__ void copy_tile_out(const uint8_t* s, uint8_t* g, int bytes) {
  const uint4* src = reinterpret_cast<const uint4*>(s);
  uint4* dst = reinterpret_cast<uint4*>(g);
  for (int i = threadIdx.x; i < bytes / 16; i += 2 * TNT) {  // 16 KB blocks (1024 chunks)
    uint4 v[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) v[j] = src[i + j * TNT];
#pragma unroll
    for (int j = 0; j < 2; ++j) dst[i + j * TNT] = v[j];
  }
}
template <int TNT = NT, int U = 2>
__device__ __forceinline__ void copy_tile_in(const uint8_t* g, uint8_t* s, int bytes) {
  // blocks of U * TNT chunks (16 KB; 32 KB for the split-mode tiles of a 512-thread CTA):
  // U independent 128-bit loads per thread in flight
  const uint4* src = reinterpret_cast<const uint4*>(g);
  uint4* dst = reinterpret_cast<uint4*>(s);
  for (int i = threadIdx.x; i < bytes / 16; i += U * TNT) {
    uint4 v[U];
#pragma unroll
    for (int j = 0; j < U; ++j) v[j] = __ldg(src + i + j * TNT);
#pragma unroll
    for (int j = 0; j < U; ++j) dst[i + j * TNT] = v[j];
  }
}


}  // namespace tc

// multi-context dgrad (rowmlp_tc_bwd_mc.cu)
bool tc_dgrad_mc_supported(const KParams& p, const tc::BGeo& g);
int tc_dgrad_mc_grid(const tc::BGeo& g);
int tc_rowmlp_dgrad_mc(const KParams& p, const tc::BGeo& g, cudaStream_t st);
// fused input + weight gradient kernel (rowmlp_tc_bwd_fused.cu)
int tc_bwd_fused_kind(const KParams& p);  // 0 no, 1 narrow inputs (no source gradients), 2 yes
int tc_bwd_fused_grid(const tc::BGeo& g);
int tc_bwd_fused_grid_bsum(const tc::BGeo& g);  // src0_batch_sum: row tiles are the dealt unit
int tc_rowmlp_bwd_fused(const KParams& p, const tc::BGeo& g, cudaStream_t st);

}  // namespace nlam
