// Kernel parameter block and tile bookkeeping shared by the fp32 (FFMA) and
// bf16 (tcgen05) row-MLP kernels.
#pragma once
#include "common.cuh"

namespace nlam {

constexpr float LN_EPS = 1e-5f;  // nn.LayerNorm default (utils.py:212)

struct KParams {
  nlam_rowmlp d;
  int k_total;  // sum of source widths
  int koff[NLAM_MAX_SRC + 1];
  int vec_ok[NLAM_MAX_SRC];  // 128-bit gather allowed
  int out_vec_ok;
  ParamLayout lay;
  // backward only
  const float* g0;
  const float* g1;
  const int32_t* g1_idx;
  const float* g1_scale;
  long long g1_batch_stride;
  float* d_src[NLAM_MAX_SRC];
  void* d_src_bf16[NLAM_MAX_SRC];  // optional bf16 shadows of the per-row source gradients
  const int32_t* g0_idx;
  const int32_t* d_src_idx[NLAM_MAX_SRC];
  int reduce_src, reduce_accumulate;
  int g0_nsum;              // > 1: dOut = sum of g0_nsum slices of g0, g0_sum_stride floats apart
  long long g0_sum_stride;
  int inputs_stable;  // fwd.src rows were not written by the kernel just before this one
  int sp_src, n_sp;         // sender pre-reduction (fused backward): source index or -1
  const int32_t* sp_tile_ptr;
  const int32_t* sp_row_ptr;
  const int32_t* sp_rows;
  int src0_batch_sum;       // d_src[0] = [1, rows, w] summed over the batch
  int split;                // tensor-core kernels: fp32 operands as bf16 hi + lo tiles
  float* a_save;
  float* dy_save;
  float* dh_save;
  float* ln_partial;  // [batch][n_tiles][2][d_out]
};


template <int TILE>
__device__ __forceinline__ void tile_range(const nlam_rowmlp& d, int tile, int& row0, int& cnt,
                                           int& chunk) {
  if (d.tile_ptr) {
    row0 = d.tile_ptr[tile];
    cnt = d.tile_ptr[tile + 1] - row0;
    chunk = d.tile_chunk ? d.tile_chunk[tile] : 0;
  } else {
    row0 = tile * TILE;
    cnt = min(TILE, d.rows - row0);
    chunk = 0;
  }
}


inline int n_tiles_of(const nlam_rowmlp& d, int tile_rows = NLAM_TILE_ROWS) {
  return d.tile_ptr ? d.n_tiles : (d.rows + tile_rows - 1) / tile_rows;
}

inline int fill_params(const nlam_rowmlp& d, KParams& p) {
  p.d = d;
  NLAM_CHECK(d.n_src >= 1 && d.n_src <= NLAM_MAX_SRC, "rowmlp: n_src=%d out of range", d.n_src);
  NLAM_CHECK(d.batch >= 1 && d.rows >= 0, "rowmlp: bad batch/rows");
  NLAM_CHECK(d.n_chunks >= 1, "rowmlp: n_chunks must be >= 1");
  NLAM_CHECK(d.n_chunks == 1 || (d.tile_ptr && d.tile_chunk && d.chunk_ptr),
             "rowmlp: chunked weights need tile_ptr/tile_chunk/chunk_ptr");
  int k = 0;
  for (int s = 0; s < d.n_src; ++s) {
    p.koff[s] = k;
    const nlam_src& src = d.src[s];
    NLAM_CHECK(src.ptr && src.width > 0 && src.ld >= src.width, "rowmlp: bad source %d", s);
    p.vec_ok[s] = (src.width % 4 == 0) && (k % 4 == 0) && (src.ld % 4 == 0) &&
                  (src.batch_stride % 4 == 0) && (((uintptr_t)src.ptr) % 16 == 0);
    k += src.width;
  }
  for (int s = d.n_src; s <= NLAM_MAX_SRC; ++s) p.koff[s] = k;
  p.k_total = k;
  NLAM_CHECK(d.residual_src == -1 || d.residual_src == 0, "rowmlp: residual_src must be -1 or 0");
  NLAM_CHECK(d.residual_src < 0 || d.src[0].width == d.d_out,
             "rowmlp: residual source width %d != d_out %d", d.src[0].width, d.d_out);
  NLAM_CHECK(!d.out_res || (d.residual_src < 0 && d.src[0].width == d.d_out),
             "rowmlp: out_res needs residual_src == -1 and src[0].width == d_out");
  p.out_vec_ok = (d.d_out % 4 == 0) && (((uintptr_t)d.out) % 16 == 0) &&
                 (((uintptr_t)d.out_res) % 16 == 0);
  p.lay = ParamLayout{k, d.d_hidden, d.d_out, d.w.ln_g != nullptr};
  p.reduce_src = -1;
  p.sp_src = -1;
  NLAM_CHECK(!(d.agg.out || d.agg.out_bf16) || (d.agg.seg_ptr && d.agg.tile_seg && d.tile_ptr && d.n_chunks == 1),
             "rowmlp: agg needs seg_ptr, tile_seg and a (receiver-aligned) tile table");
  NLAM_CHECK(d.out || d.out_res || d.agg.out || true, "unreachable");
  return 0;
}


}  // namespace nlam
