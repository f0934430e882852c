// bf16 tensor-core (tcgen05) row-MLP BACKWARD for sm_100a: two kernels.
//
// dgrad (per 128-row tile, recompute-based -- only layer inputs were saved):
//   gather z -> GEMM1 (H) -> a = SiLU(H+b1)            [a tile -> HBM image]
//   GEMM2 (Y) -> LayerNorm stats, dOut rows (fp32, staged coalesced) ->
//   dY = LN'(dOut)                                       [dY tile -> HBM image]
//   GEMM3: dA = dY . W2   (W2 tile re-used as MN-major B operand)
//   dH = dA * SiLU'(H+b1)                                [dH tile -> HBM image]
//   GEMM4: dZ = dH . W1   (W1 blocks re-used as MN-major B), 64 input columns
//   at a time, double-buffered in TMEM; rows stored coalesced per source.
//   Column sums for db1, db2, dLN gamma/beta: 16-shuffle warp transposes of the
//   fp32 epilogue registers, accumulated per CTA across tiles (no atomics).
// wgrad (per tile): dW1^T += z^T . dH, dW2^T += a^T . dY as M=128 UMMAs whose
//   A/B operands are the SAME bf16 tiles viewed MN-major; fp32 accumulators stay
//   in TMEM across all tiles of a persistent CTA, then go to a per-CTA partial
//   that a fixed-order reduction sums (deterministic).
//
// Reference: autograd of utils.make_mlp / InteractionNet.message / aggr_mlp
// (utils.py:191-214, interaction_net.py:106,117-121).
#include <stdlib.h>

#include "rowmlp_tc_bwd.cuh"

namespace nlam {

// defined in rowmlp_simt.cu
int launch_reduce_params(const float* partial, int splits, int n_chunks, int p_total,
                         float* out, int accumulate, const float* vec_partial, int vec_slots,
                         int vec_len, ParamLayout lay, cudaStream_t st, bool defer);

namespace tc {

constexpr int MAXCH = 4;  // 16-column chunks per thread (n <= 128, two column halves)

// TNT = threads per CTA (256, or 512 at d = 128 where one CTA fits per SM; see rowmlp_tc.cu)
// SP: fp32 operands as split bf16 tiles (rowmlp_tc.cuh, put8): every tile below has a lo twin
// right behind it, every product is three UMMAs, the HBM tile images carry both parts.
template <int FN, bool FG, int TNT, bool SP = false>
__global__ void __launch_bounds__(TNT, TNT == 256 ? 2 : 1)
rowmlp_tc_dgrad_kernel(const __grid_constant__ KParams p, const __grid_constant__ BGeo g) {
  constexpr bool F = FN > 0;  // square fast path: sizes are compile-time constants
  constexpr int NG = TNT / 128;   // column groups (threads) per tile row
  constexpr int ZC = 64 / NG;     // columns of a 64-wide dZ block per thread
  constexpr int ORP = TNT / 16, ONP = TM / ORP;  // output phases: rows per pass, passes
  const int n1 = F ? FN : g.n1, n2 = F ? FN : g.n2;
  const int k2 = F ? FN : g.k2, ko = F ? FN : g.ko;
  const int kb2 = F ? FN / 64 : g.kb2, kbo = F ? FN / 64 : g.kbo;
  extern __shared__ __align__(1024) uint8_t sm[];
  if (smem_u32(sm) & 1023u) __trap();
  uint8_t* sA = sm;                     // z blocks of one gather round | fp32 staging
  float* stg = reinterpret_cast<float*>(sm);
  uint8_t* sT = sm + g.off_t;           // a -> dY -> dH bf16 tile
  uint8_t* sW1 = sm + g.off_w1;
  uint8_t* sW2 = sm + g.off_w2;
  float* sPar = reinterpret_cast<float*>(sm + g.off_par);
  float* sLnx = reinterpret_cast<float*>(sm + g.off_lnx);  // [TM][2]
  float* sRed = reinterpret_cast<float*>(sm + g.off_t);    // [warps][4][64], aliases sT
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + g.off_bar);  // [0] main, [1],[2] dZ
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dh = F ? FN : p.d.d_hidden, dout = F ? FN : p.d.d_out;
  const bool has_ln = p.d.w.ln_g != nullptr;

  if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)g.tmem_cols);
  if (tid == 32) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tH = tmem_base, tY = tmem_base + (uint32_t)g.cY, tZ = tmem_base + (uint32_t)g.cZ;
  uint32_t ph_main = 0, ph_z = 0;
  int loaded_chunk = -1;

  const uint32_t a_blk = TM * 128u;
  constexpr uint32_t P = SP ? 2u : 1u;
  // split mode: byte offsets of the lo tiles (z round region, a / dH tile, dY tile, W1, W2)
  const uint32_t z_lo = (uint32_t)min(g.kb1, g.rb) * a_blk;
  const uint32_t t2_lo = (uint32_t)kb2 * a_blk, to_lo = (uint32_t)kbo * a_blk;
  const uint32_t w1_lo = (uint32_t)g.kb1 * (uint32_t)n1 * 128u, w2_lo = (uint32_t)kb2 * (uint32_t)n2 * 128u;
  const uint32_t idesc1 = make_idesc_bf16(TM, n1);
  const uint32_t idesc2 = make_idesc_bf16(TM, n2);
  const uint32_t idesc3 = make_idesc_bf16(TM, n1, 0, 1);  // B = W2 viewed MN-major
  const uint32_t idesc4 = make_idesc_bf16(TM, 64, 0, 1);    // B = W1 block viewed MN-major

  const int q = warp & 3, hf = warp >> 2, r = q * 32 + lane;
  const int cp1 = n1 >= 16 * NG ? n1 / NG : n1, cp2 = n2 >= 16 * NG ? n2 / NG : n2;
  const bool act1 = n1 >= 16 * NG || hf == 0, act2 = n2 >= 16 * NG || hf == 0;
  const bool split2 = n2 >= 16 * NG;
  const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
  const float* sB2 = sPar + n1;
  const float* sG = sB2 + n2;
  // Sum over the column groups of row r of one per-thread partial value, returned to every
  // thread of the row; fixed order (deterministic).  The exchange array is [TM][2] floats
  // (shared memory is full at d = 128): four groups go through it in two steps, pair sums
  // (0+1), (2+3) first.  Ends with a barrier: the array may be reused right away.
  auto row_allsum = [&](float v) {
    float tot;
    if (NG == 4 && split2) {
      if (hf & 1) sLnx[r * 2 + (hf >> 1)] = v;
      __syncthreads();
      float pair = v;
      if (!(hf & 1)) pair = v + sLnx[r * 2 + (hf >> 1)];
      __syncthreads();
      if (!(hf & 1)) sLnx[r * 2 + (hf >> 1)] = pair;
      __syncthreads();
      tot = sLnx[r * 2] + sLnx[r * 2 + 1];
    } else {
      if (hf < 2) sLnx[r * 2 + hf] = v;  // (four groups, narrow output: group 0 holds the row)
      __syncthreads();
      tot = sLnx[r * 2] + (split2 ? sLnx[r * 2 + 1] : 0.f);
    }
    __syncthreads();
    return tot;
  };

  // per-CTA column-sum accumulators (lane l owns column c0 + (l & 15) of each chunk)
  float acc_db1[MAXCH], acc_db2[MAXCH], acc_dg[MAXCH], acc_dbt[MAXCH];
#pragma unroll
  for (int i = 0; i < MAXCH; ++i) acc_db1[i] = acc_db2[i] = acc_dg[i] = acc_dbt[i] = 0.f;

  auto flush_colsums = [&](int chunk) {
    // sRed[warp][which][col in half]; quarters q = 0..3 are summed in fixed order
    __syncthreads();
    if (lane < 16) {
#pragma unroll
      for (int i = 0; i < MAXCH; ++i) {
        sRed[(warp * 4 + 0) * 64 + i * 16 + lane] = acc_db1[i];
        sRed[(warp * 4 + 1) * 64 + i * 16 + lane] = acc_db2[i];
        sRed[(warp * 4 + 2) * 64 + i * 16 + lane] = acc_dg[i];
        sRed[(warp * 4 + 3) * 64 + i * 16 + lane] = acc_dbt[i];
      }
    }
    __syncthreads();
    float* dst = g.vec_partial + ((size_t)blockIdx.x * p.d.n_chunks + chunk) * g.vec_len;
    // which: 0 db1 (n1 cols), 1 db2, 2 dgamma, 3 dbeta (n2 cols)
    for (int e = tid; e < 4 * 128; e += TNT) {
      const int which = e >> 7, col = e & 127;
      const int n = which == 0 ? n1 : n2, cp = which == 0 ? cp1 : cp2;
      const int real = which == 0 ? dh : dout;
      if (col >= n || col >= real) continue;
      if (which >= 2 && !has_ln) continue;
      const int h = col / cp, cc = col % cp;  // column half and offset inside it
      float s = 0.f;
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) s += sRed[((h * 4 + qq) * 4 + which) * 64 + cc];
      const int off = which == 0 ? 0 : which == 1 ? dh : which == 2 ? dh + dout : dh + 2 * dout;
      // several weight sets: a CTA may come back to a chunk (zero-initialised slots)
      dst[off + col] = p.d.n_chunks > 1 ? dst[off + col] + s : s;
    }
#pragma unroll
    for (int i = 0; i < MAXCH; ++i) acc_db1[i] = acc_db2[i] = acc_dg[i] = acc_dbt[i] = 0.f;
  };

  // pipelined gather (64-wide sources): row indices are fetched one tile ahead
  // (off in this kernel: the extra live registers spill and cost more than the gather gains)
  constexpr bool PIPE = false;
  int nidx[NLAM_MAX_SRC] = {-1, -1, -1};
  if (PIPE && (int)blockIdx.x < g.total_tiles) {
    int r0, c0, ch0;
    tile_range<TM>(p.d, blockIdx.x / p.d.batch, r0, c0, ch0);
    load_row_idx<TNT>(p, r0, c0, tid, nidx);
  }

  for (int t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
    const int b = t % p.d.batch, tile = t / p.d.batch;  // batch innermost: shared rows hit L2
    int row0, cnt, chunk;
    tile_range<TM>(p.d, tile, row0, cnt, chunk);
    const size_t grow0 = (size_t)b * p.d.rows + row0;

    if (chunk != loaded_chunk) {
      if (loaded_chunk >= 0) flush_colsums(loaded_chunk);
      stage_weight<TNT, SP>(p.d.w.w1 + (size_t)chunk * dh * p.k_total, dh, p.k_total, n1, g.k1, sW1);
      stage_weight<TNT, SP>(p.d.w.w2 + (size_t)chunk * dout * dh, dout, dh, n2, k2, sW2);
      stage_params<TNT>(p.d, chunk, n1, n2, sPar, 2);  // beta is not needed backward
      loaded_chunk = chunk;
    }

    // dOut = g0 rows (+ scale * gathered g1 rows): batches of 4 units per thread,
    // row indices first, then all data loads; the combine + store happens later
    auto dm_load = [&](int base, float4 (&va)[4], float4 (&vb)[4], float (&gs)[4],
                       int (&rowv)[4], int (&c4v)[4]) {
      const int w4 = n2 >> 2;
      const bool vec = (dout & 3) == 0;
      const float* g0p[4];
      const float* g1p[4];
      int colv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int u = base + j * TNT;
        g0p[j] = g1p[j] = nullptr;
        gs[j] = 1.f;
        rowv[j] = -1, c4v[j] = 0, colv[j] = 0;
        if (u < TM * w4) {
          const int row = u / w4, c4 = u % w4, col = c4 * 4;
          rowv[j] = row, c4v[j] = c4, colv[j] = col;
          if (row < cnt && col < dout) {
            if (p.g0) {
              const size_t gr = p.g0_idx ? (size_t)b * p.d.rows + __ldg(p.g0_idx + row0 + row)
                                         : grow0 + row;
              g0p[j] = p.g0 + gr * dout + col;
            }
            if (p.g1) {
              const int gi = __ldg(p.g1_idx + row0 + row);
              if (p.g1_scale) gs[j] = __ldg(p.g1_scale + gi);
              g1p[j] = p.g1 + (size_t)b * p.g1_batch_stride + (size_t)gi * dout + col;
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        va[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        vb[j] = va[j];
        if (vec) {
          if (g0p[j]) va[j] = __ldg(reinterpret_cast<const float4*>(g0p[j]));
          if (g1p[j]) vb[j] = __ldg(reinterpret_cast<const float4*>(g1p[j]));
        } else {
          float t0[4] = {0.f, 0.f, 0.f, 0.f}, t1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (rowv[j] >= 0 && colv[j] + e < dout) {
              if (g0p[j]) t0[e] = __ldg(g0p[j] + e);
              if (g1p[j]) t1[e] = __ldg(g1p[j] + e);
            }
          va[j] = make_float4(t0[0], t0[1], t0[2], t0[3]);
          vb[j] = make_float4(t1[0], t1[1], t1[2], t1[3]);
        }
      }
    };
    auto dm_store = [&](const float4 (&va)[4], const float4 (&vb)[4], const float (&gs)[4],
                        const int (&rowv)[4], const int (&c4v)[4]) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (rowv[j] < 0) continue;
        const float4 v = make_float4(va[j].x + gs[j] * vb[j].x, va[j].y + gs[j] * vb[j].y,
                                     va[j].z + gs[j] * vb[j].z, va[j].w + gs[j] * vb[j].w);
        *reinterpret_cast<float4*>(stg + stg_idx(rowv[j], c4v[j], n2)) = v;
      }
    };
    float4 dm0_a[4], dm0_b[4];
    float dm0_s[4];
    int dm0_row[4], dm0_c4[4];

    // ---------------- gather rounds + GEMM 1 (H = z . W1^T)
    for (int kb0 = 0; kb0 < g.kb1; kb0 += g.rb) {
      const int kbe = min(g.kb1, kb0 + g.rb);
      const int k_begin = kb0 * 64, k_end = min(g.k1, kbe * 64);
      if (PIPE) {  // one round covers all (<= 3) sources
        const int cidx[NLAM_MAX_SRC] = {nidx[0], nidx[1], nidx[2]};
        gather_rows_pipe<64, TNT>(p, b, cidx, sA, tid);
      } else if (F && FG) {
        gather_rows_fast<(F ? FN : 64), TNT, SP>(p, b, row0, cnt, k_begin / (F ? FN : 64),
                                                 k_end / (F ? FN : 64), sA, z_lo);
      } else {
        gather_rows<TNT, SP>(p, b, row0, cnt, k_begin, k_end, sA, z_lo);
      }
      fence_async_smem();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint32_t a0 = smem_u32(sA), w0 = smem_u32(sW1);
        const uint32_t w_blk = (uint32_t)n1 * 128u;
        for (int ks = k_begin / 16; ks < k_end / 16; ++ks) {
          const uint32_t kb = ks >> 2, kin = (ks & 3) * 32;
          umma_bf16(tH, make_desc_k_sw128(a0 + (kb - kb0) * a_blk + kin),
                    make_desc_k_sw128(w0 + kb * w_blk + kin), idesc1, ks > 0);
          if (SP) {
            umma_bf16(tH, make_desc_k_sw128(a0 + (kb - kb0) * a_blk + kin),
                      make_desc_k_sw128(w0 + w1_lo + kb * w_blk + kin), idesc1, 1u);
            umma_bf16(tH, make_desc_k_sw128(a0 + z_lo + (kb - kb0) * a_blk + kin),
                      make_desc_k_sw128(w0 + kb * w_blk + kin), idesc1, 1u);
          }
        }
        umma_commit(&bars[0]);
      }
      // first batch of dOut loads flies while the last GEMM-1 round executes
      if (kbe == g.kb1) dm_load(tid, dm0_a, dm0_b, dm0_s, dm0_row, dm0_c4);
      mbar_wait(&bars[0], ph_main);
      ph_main ^= 1;
      tc_fence_after();
    }

    // ---------------- dOut rows -> fp32 staging (coalesced), rows >= cnt are zero.
    // The first batch of loads was issued before waiting for GEMM 1 (dm0_*).
    dm_store(dm0_a, dm0_b, dm0_s, dm0_row, dm0_c4);
    for (int base = tid + TNT * 4; base < TM * (n2 >> 2); base += TNT * 4) {
      float4 va[4], vb[4];
      float gs[4];
      int rowv[4], c4v[4];
      dm_load(base, va, vb, gs, rowv, c4v);
      dm_store(va, vb, gs, rowv, c4v);
    }

    {  // L2 prefetch of the next tile's input and dOut rows
      const int tn = t + gridDim.x;
      if (tn < g.total_tiles) {
        int r0n, cn, chn;
        const int bn = tn % p.d.batch;
        tile_range<TM>(p.d, tn / p.d.batch, r0n, cn, chn);
        if (PIPE) {
          load_row_idx<TNT>(p, r0n, cn, tid, nidx);
          prefetch_rows_of(p, bn, nidx, (tid & 31) < 16);
        } else {
          prefetch_sources(p, bn, r0n, cn);
        }
        if (p.g0)
          prefetch_tile_rows(p.g0 + (size_t)bn * p.d.rows * dout, p.g0_idx, dout, dout, r0n, cn);
        if (p.g1)
          prefetch_tile_rows(p.g1 + (size_t)bn * p.g1_batch_stride, p.g1_idx, dout, dout, r0n, cn);
      }
    }

    // ---------------- epilogue 1: a = SiLU(H + b1) -> bf16 tile
    if (act1) {
      for (int cc = 0; cc < cp1; cc += 16) {
        const int c0 = hf * cp1 + cc;
        float v[16];
        tmem_ld16(tH + lane_addr + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = silu_sel<SP>(v[j] + sPar[c0 + j]);
        put16<SP>(sT, t2_lo, r, c0, a_blk, v);
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();

    // ---------------- GEMM 2: Y = a . W2^T ; meanwhile the a tile goes to HBM
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a0 = smem_u32(sT), w0 = smem_u32(sW2);
      const uint32_t w_blk = (uint32_t)n2 * 128u;
      for (int ks = 0; ks < k2 / 16; ++ks) {
        const uint32_t kb = ks >> 2, kin = (ks & 3) * 32;
        umma_bf16(tY, make_desc_k_sw128(a0 + kb * a_blk + kin),
                  make_desc_k_sw128(w0 + kb * w_blk + kin), idesc2, ks > 0);
        if (SP) {
          umma_bf16(tY, make_desc_k_sw128(a0 + kb * a_blk + kin),
                    make_desc_k_sw128(w0 + w2_lo + kb * w_blk + kin), idesc2, 1u);
          umma_bf16(tY, make_desc_k_sw128(a0 + t2_lo + kb * a_blk + kin),
                    make_desc_k_sw128(w0 + kb * w_blk + kin), idesc2, 1u);
        }
      }
      umma_commit(&bars[0]);
    }
    copy_tile_out<TNT>(sT, g.a_img + (size_t)t * P * kb2 * a_blk, P * kb2 * a_blk);
    mbar_wait(&bars[0], ph_main);
    ph_main ^= 1;
    tc_fence_after();

    // ---------------- epilogue 2: LayerNorm backward -> dY (bf16 tile) + column sums
    float mean = 0.f, rstd = 1.f, m1 = 0.f, m2 = 0.f;
    if (has_ln) {
      float s = 0.f;
      if (act2)
        for (int cc = 0; cc < cp2; cc += 16) {
          const int c0 = hf * cp2 + cc;
          float v[16];
          tmem_ld16(tY + lane_addr + (uint32_t)c0, v);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (F || c0 + j < dout) s += v[j] + sB2[c0 + j];
        }
      mean = row_allsum(s) / (float)dout;
      float qq = 0.f;
      if (act2)
        for (int cc = 0; cc < cp2; cc += 16) {
          const int c0 = hf * cp2 + cc;
          float v[16];
          tmem_ld16(tY + lane_addr + (uint32_t)c0, v);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (F || c0 + j < dout) {
              const float dl = v[j] + sB2[c0 + j] - mean;
              qq += dl * dl;
            }
        }
      rstd = rsqrtf(row_allsum(qq) / (float)dout + LN_EPS);
      // row means of dyhat and dyhat*yhat; column sums for dgamma / dbeta
      float s1 = 0.f, s2 = 0.f;
      if (act2) {
#pragma unroll
        for (int ci = 0; ci < MAXCH; ++ci) {
          const int cc = ci * 16;
          if (cc < cp2) {
            const int c0 = hf * cp2 + cc;
            float v[16], dmv[16], pv[16];
            tmem_ld16(tY + lane_addr + (uint32_t)c0, v);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 d4 =
                  *reinterpret_cast<const float4*>(stg + stg_idx(r, (c0 >> 2) + j4, n2));
              dmv[j4 * 4] = d4.x, dmv[j4 * 4 + 1] = d4.y, dmv[j4 * 4 + 2] = d4.z,
                       dmv[j4 * 4 + 3] = d4.w;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float yh = (v[j] + sB2[c0 + j] - mean) * rstd;
              const float dyh = dmv[j] * sG[c0 + j];
              pv[j] = (F || c0 + j < dout) ? dmv[j] * yh : 0.f;
              if (F || c0 + j < dout) {
                s1 += dyh;
                s2 += dyh * yh;
              }
            }
            acc_dg[ci] += warp_colsum16(pv, lane);
            acc_dbt[ci] += warp_colsum16(dmv, lane);
          }
        }
      }
      m1 = row_allsum(s1) / (float)dout;
      m2 = row_allsum(s2) / (float)dout;
    }
    __syncthreads();  // every thread is done copying the a tile out of sT
    if (act2) {
#pragma unroll
      for (int ci = 0; ci < MAXCH; ++ci) {
        const int cc = ci * 16;
        if (cc < cp2) {
          const int c0 = hf * cp2 + cc;
          float v[16], dmv[16];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 d4 =
                *reinterpret_cast<const float4*>(stg + stg_idx(r, (c0 >> 2) + j4, n2));
            dmv[j4 * 4] = d4.x, dmv[j4 * 4 + 1] = d4.y, dmv[j4 * 4 + 2] = d4.z,
                     dmv[j4 * 4 + 3] = d4.w;
          }
          if (has_ln) {
            tmem_ld16(tY + lane_addr + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float yh = (v[j] + sB2[c0 + j] - mean) * rstd;
              const float dyh = dmv[j] * sG[c0 + j];
              v[j] = ((F || c0 + j < dout) && r < cnt) ? rstd * (dyh - m1 - yh * m2) : 0.f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = (F || c0 + j < dout) ? dmv[j] : 0.f;
          }
          acc_db2[ci] += warp_colsum16(v, lane);
          put16<SP>(sT, to_lo, r, c0, a_blk, v);
        }
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();

    // ---------------- GEMM 3: dA = dY . W2 (into Y's columns) ; dY tile -> HBM
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a0 = smem_u32(sT), w0 = smem_u32(sW2);
      const uint32_t lbo = (uint32_t)n2 * 128u;  // 64-wide blocks of the hidden dim
      for (int ks = 0; ks < ko / 16; ++ks) {
        const uint32_t kb = ks >> 2, kin = (ks & 3) * 32;
        umma_bf16(tY, make_desc_k_sw128(a0 + kb * a_blk + kin),
                  make_desc_mn_sw128(w0 + (uint32_t)ks * 2048u, lbo), idesc3, ks > 0);
        if (SP) {
          umma_bf16(tY, make_desc_k_sw128(a0 + kb * a_blk + kin),
                    make_desc_mn_sw128(w0 + w2_lo + (uint32_t)ks * 2048u, lbo), idesc3, 1u);
          umma_bf16(tY, make_desc_k_sw128(a0 + to_lo + kb * a_blk + kin),
                    make_desc_mn_sw128(w0 + (uint32_t)ks * 2048u, lbo), idesc3, 1u);
        }
      }
      umma_commit(&bars[0]);
    }
    copy_tile_out<TNT>(sT, g.dy_img + (size_t)t * P * kbo * a_blk, P * kbo * a_blk);
    mbar_wait(&bars[0], ph_main);
    ph_main ^= 1;
    tc_fence_after();
    __syncthreads();  // dY tile fully copied before dH overwrites it

    // ---------------- epilogue 3: dH = dA * SiLU'(H + b1) -> bf16 tile
    if (act1) {
#pragma unroll
      for (int ci = 0; ci < MAXCH; ++ci) {
        const int cc = ci * 16;
        if (cc < cp1) {
          const int c0 = hf * cp1 + cc;
          float v[16], h[16];
          tmem_ld16(tY + lane_addr + (uint32_t)c0, v);
          tmem_ld16(tH + lane_addr + (uint32_t)c0, h);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            v[j] = (F || c0 + j < dh) ? v[j] * silu_grad_sel<SP>(h[j] + sPar[c0 + j]) : 0.f;
          acc_db1[ci] += warp_colsum16(v, lane);
          put16<SP>(sT, t2_lo, r, c0, a_blk, v);
        }
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    copy_tile_out<TNT>(sT, g.dh_img + (size_t)t * P * kb2 * a_blk, P * kb2 * a_blk);

    // ---------------- GEMM 4 + epilogue 4: dZ = dH . W1, 64 input columns at a time
    if (g.need_dz) {
      auto issue_dz = [&](int kb) {
        tc_fence_after();
        const uint32_t a0 = smem_u32(sT);
        const uint32_t w0 = smem_u32(sW1) + (uint32_t)kb * (uint32_t)n1 * 128u;
        for (int ks = 0; ks < k2 / 16; ++ks) {
          const uint32_t kbb = ks >> 2, kin = (ks & 3) * 32;
          umma_bf16(tZ + (uint32_t)(kb & 1) * 64u, make_desc_k_sw128(a0 + kbb * a_blk + kin),
                    make_desc_mn_sw128(w0 + (uint32_t)ks * 2048u, 0), idesc4, ks > 0);
          if (SP) {
            umma_bf16(tZ + (uint32_t)(kb & 1) * 64u, make_desc_k_sw128(a0 + kbb * a_blk + kin),
                      make_desc_mn_sw128(w0 + w1_lo + (uint32_t)ks * 2048u, 0), idesc4, 1u);
            umma_bf16(tZ + (uint32_t)(kb & 1) * 64u,
                      make_desc_k_sw128(a0 + t2_lo + kbb * a_blk + kin),
                      make_desc_mn_sw128(w0 + (uint32_t)ks * 2048u, 0), idesc4, 1u);
          }
        }
        umma_commit(&bars[1 + (kb & 1)]);
      };
      if (tid == 0) issue_dz(0);
      for (int kb = 0; kb < g.kb1; ++kb) {
        if (tid == 0 && kb + 1 < g.kb1) issue_dz(kb + 1);
        // fast path: block kb = 64 columns [col0, col0+64) of source fs
        constexpr int FNN = F ? FN : 64;
        const int fs = (kb * 64) / FNN, col0 = (kb * 64) % FNN;
        float* fdst = (F && FG) ? p.d_src[fs] : nullptr;
        const bool fres = F && FG && fdst && (fs == p.d.residual_src) && p.g0;
        const int32_t* didx = fdst ? p.d_src_idx[fs] : nullptr;
        int orow_i[ONP];
        if (didx && fs != p.reduce_src) {  // scatter targets, fetched while the MMA runs
#pragma unroll
          for (int i = 0; i < ONP; ++i) {
            const int row = (tid >> 4) + ORP * i;
            orow_i[i] = row < cnt ? __ldg(didx + row0 + row) : 0;
          }
        }
        float4 e[ONP];
        if (fres) {  // residual rows requested early: they arrive while the MMA runs
#pragma unroll
          for (int i = 0; i < ONP; ++i) {
            const int row = (tid >> 4) + ORP * i;
            e[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < cnt) {
              const size_t gr = p.g0_idx ? (size_t)b * p.d.rows + __ldg(p.g0_idx + row0 + row)
                                         : grow0 + row;
              e[i] = __ldg(reinterpret_cast<const float4*>(p.g0 + gr * FNN + col0) + (tid & 15));
            }
          }
        }
        mbar_wait(&bars[1 + (kb & 1)], (ph_z >> (kb & 1)) & 1u);
        ph_z ^= 1u << (kb & 1);
        tc_fence_after();
        // TMEM -> swizzled fp32 staging [128][64]
        for (int cc = 0; cc < ZC; cc += 16) {
          const int c0 = hf * ZC + cc;
          float v[16];
          tmem_ld16(tZ + (uint32_t)(kb & 1) * 64u + lane_addr + (uint32_t)c0, v);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4)
            *reinterpret_cast<float4*>(stg + stg_idx(r, (c0 >> 2) + j4, 64)) =
                make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
        }
        tc_fence_before();
        __syncthreads();
        if (F && FG) {
          if (fdst && fs == p.reduce_src) {
            // receiver-aligned tile: sum the gradient rows of each segment (fixed order)
            const int seg_lo = __ldg(p.d.agg.tile_seg + tile);
            const int seg_hi = __ldg(p.d.agg.tile_seg + tile + 1);
            float* ro = fdst + (size_t)b * p.d.agg.n_seg * FNN + col0 + (tid & 15) * 4;
            for (int seg = seg_lo + (tid >> 4); seg < seg_hi; seg += ORP) {
              const int r0 = __ldg(p.d.agg.seg_ptr + seg) - row0;
              const int r1 = __ldg(p.d.agg.seg_ptr + seg + 1) - row0;
              float4* o4 = reinterpret_cast<float4*>(ro + (size_t)seg * FNN);
              float4 old = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.reduce_accumulate) old = *o4;  // in flight during the shared-memory sum
              float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
              for (int rr = r0; rr < r1; ++rr) {
                const float4 v = *reinterpret_cast<const float4*>(stg + stg_idx(rr, tid & 15, 64));
                acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
              }
              acc.x += old.x, acc.y += old.y, acc.z += old.z, acc.w += old.w;
              *o4 = acc;
            }
          } else if (fdst) {
            float* o = fdst + col0 + (tid & 15) * 4;
#pragma unroll
            for (int i = 0; i < ONP; ++i) {
              const int row = (tid >> 4) + ORP * i;
              if (row < cnt) {
                float4 v = *reinterpret_cast<const float4*>(stg + stg_idx(row, tid & 15, 64));
                if (fres) v.x += e[i].x, v.y += e[i].y, v.z += e[i].z, v.w += e[i].w;
                const size_t orow = didx ? (size_t)b * p.d.rows + orow_i[i] : grow0 + row;
                *reinterpret_cast<float4*>(o + orow * FNN) = v;
              }
            }
          }
        } else {
          // generic: coalesced per-source row stores
          for (int u = tid; u < cnt * 16; u += TNT) {
            const int row = u >> 4, c4 = u & 15, kg = kb * 64 + c4 * 4;
            if (kg >= p.k_total) continue;
            int s = 0;
            while (s + 1 < p.d.n_src && kg >= p.koff[s + 1]) ++s;
            float* dst = p.d_src[s];
            const int w = p.d.src[s].width, col = kg - p.koff[s];
            const float4 v = *reinterpret_cast<const float4*>(stg + stg_idx(row, c4, 64));
            float tmp[4] = {v.x, v.y, v.z, v.w};
            const bool res = (s == p.d.residual_src) && p.g0;
            if (dst && (w & 3) == 0) {
              float* o = dst + (grow0 + row) * w + col;
              if (res) {
                const float4 ee =
                    __ldg(reinterpret_cast<const float4*>(p.g0 + (grow0 + row) * dout + col));
                tmp[0] += ee.x, tmp[1] += ee.y, tmp[2] += ee.z, tmp[3] += ee.w;
              }
              *reinterpret_cast<float4*>(o) = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
            } else {
              // element by element: a 4-column group may straddle sources (odd widths)
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int k = kg + j;
                if (k >= p.k_total) continue;
                int s2 = s;
                while (s2 + 1 < p.d.n_src && k >= p.koff[s2 + 1]) ++s2;
                float* d2 = p.d_src[s2];
                if (!d2) continue;
                const int c2 = k - p.koff[s2];
                const bool res2 = (s2 == p.d.residual_src) && p.g0;
                d2[(grow0 + row) * p.d.src[s2].width + c2] =
                    tmp[j] + (res2 ? __ldg(p.g0 + (grow0 + row) * dout + c2) : 0.f);
              }
            }
          }
        }
        __syncthreads();
      }
    }
    tc_fence_before();
    __syncthreads();  // sT / staging free for the next tile
  }
  if (loaded_chunk >= 0) flush_colsums(loaded_chunk);

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
}

// -------------------------------------------------------------------- wgrad
template <int FN, bool FG, int TNT, bool SP = false>
__global__ void __launch_bounds__(TNT, TNT == 256 ? 2 : 1)
rowmlp_tc_wgrad_kernel(const __grid_constant__ KParams p, const __grid_constant__ BGeo g) {
  constexpr bool F = FN > 0;
  const int n1 = F ? FN : g.n1, n2 = F ? FN : g.n2;
  const int kb2 = F ? FN / 64 : g.kb2, kbo = F ? FN / 64 : g.kbo;
  extern __shared__ __align__(1024) uint8_t sm[];
  if (smem_u32(sm) & 1023u) __trap();
  uint8_t* sZ = sm;
  uint8_t* sAi = sm + g.w_off_a;
  uint8_t* sDY = sm + g.w_off_dy;
  uint8_t* sDH = sm + g.w_off_dh;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + g.w_off_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dh = F ? FN : p.d.d_hidden, dout = F ? FN : p.d.d_out;
  if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)g.w_tmem_cols);
  if (tid == 32) {
    mbar_init(&bars[0], 1);
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tW2 = tmem_base + (uint32_t)(g.w_mchunks * n1);
  uint32_t ph = 0;
  const uint32_t a_blk = TM * 128u;
  constexpr uint32_t P = SP ? 2u : 1u;
  const uint32_t z_lo = (uint32_t)g.kb1 * a_blk, t2_lo = (uint32_t)kb2 * a_blk, to_lo = (uint32_t)kbo * a_blk;
  const uint32_t idesc_w1 = make_idesc_bf16(TM, n1, 1, 1);
  const uint32_t idesc_w2 = make_idesc_bf16(TM, n2, 1, 1);
  const int q = warp & 3, hf = warp >> 2, r = q * 32 + lane;
  const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
  int cur_chunk = -1;
  bool first = true;

  auto flush = [&](int chunk) {
    // accumulators -> per-CTA partial in final [n][k] orientation (added to the
    // zero-initialised slot when there are several weight sets: a CTA may revisit one)
    const bool multi = p.d.n_chunks > 1;
    tc_fence_after();
    const ParamLayout lay = p.lay;
    float* dst = g.partial + ((size_t)blockIdx.x * p.d.n_chunks + chunk) * g.p_total;
    constexpr int NG = TNT / 128;
    const int cpa = n1 >= 16 * NG ? n1 / NG : n1;
    if (n1 >= 16 * NG || hf == 0) {
      for (int mc = 0; mc < g.w_mchunks; ++mc) {
        const int kg = mc * 128 + r;  // input column
        for (int cc = 0; cc < cpa; cc += 16) {
          const int c0 = hf * cpa + cc;
          float v[16];
          tmem_ld16(tmem_base + (uint32_t)(mc * n1) + lane_addr + (uint32_t)c0, v);
          if (kg < p.k_total) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (F || c0 + j < dh) {
                float* o = dst + lay.off_w1() + (size_t)(c0 + j) * p.k_total + kg;
                *o = multi ? *o + v[j] : v[j];
              }
          }
        }
      }
    }
    const int cpb = n2 >= 16 * NG ? n2 / NG : n2;
    if (n2 >= 16 * NG || hf == 0) {
      for (int cc = 0; cc < cpb; cc += 16) {
        const int c0 = hf * cpb + cc;
        float v[16];
        tmem_ld16(tW2 + lane_addr + (uint32_t)c0, v);
        if (r < dh) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (F || c0 + j < dout) {
              float* o = dst + lay.off_w2() + (size_t)(c0 + j) * dh + r;
              *o = multi ? *o + v[j] : v[j];
            }
        }
      }
    }
    tc_fence_before();
    __syncthreads();
  };

  constexpr bool PIPE = F && FG && FN == 64 && TNT == 256 && !SP;
  int nidx[NLAM_MAX_SRC] = {-1, -1, -1};
  if (PIPE && (int)blockIdx.x < g.total_tiles) {
    int r0, c0, ch0;
    tile_range<TM>(p.d, blockIdx.x / p.d.batch, r0, c0, ch0);
    load_row_idx<TNT>(p, r0, c0, tid, nidx);
  }

  for (int t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
    const int b = t % p.d.batch, tile = t / p.d.batch;  // batch innermost: shared rows hit L2
    int row0, cnt, chunk;
    tile_range<TM>(p.d, tile, row0, cnt, chunk);
    if (chunk != cur_chunk) {
      if (cur_chunk >= 0) flush(cur_chunk);
      cur_chunk = chunk;
      first = true;
    }
    if (PIPE) {
      const int cidx[NLAM_MAX_SRC] = {nidx[0], nidx[1], nidx[2]};
      gather_rows_pipe<64, TNT>(p, b, cidx, sZ, tid);
    } else if (F && FG) {
      gather_rows_fast<(F ? FN : 64), TNT, SP>(p, b, row0, cnt, 0, p.d.n_src, sZ, z_lo);
    } else {
      gather_rows<TNT, SP>(p, b, row0, cnt, 0, g.k1, sZ, z_lo);
    }
    constexpr int CU = (SP && TNT == 512) ? 4 : 2;  // split tiles are multiples of 32 KB
    copy_tile_in<TNT, CU>(g.a_img + (size_t)t * P * kb2 * a_blk, sAi, P * kb2 * a_blk);
    copy_tile_in<TNT, CU>(g.dy_img + (size_t)t * P * kbo * a_blk, sDY, P * kbo * a_blk);
    copy_tile_in<TNT, CU>(g.dh_img + (size_t)t * P * kb2 * a_blk, sDH, P * kb2 * a_blk);
    {  // L2 prefetch of the next tile's rows and bf16 tile images
      const int tn = t + gridDim.x;
      if (tn < g.total_tiles) {
        int r0n, cn, chn;
        tile_range<TM>(p.d, tn / p.d.batch, r0n, cn, chn);
        if (PIPE) {
          load_row_idx<TNT>(p, r0n, cn, tid, nidx);
          prefetch_rows_of(p, tn % p.d.batch, nidx, (tid & 31) < 16);
        } else {
          prefetch_sources(p, tn % p.d.batch, r0n, cn);
        }
        const int li = (int)(P * kb2 * a_blk) >> 7, lo = (int)(P * kbo * a_blk) >> 7;
        for (int u = tid; u < li; u += TNT) {
          prefetch_l2(g.a_img + (size_t)tn * P * kb2 * a_blk + (size_t)u * 128);
          prefetch_l2(g.dh_img + (size_t)tn * P * kb2 * a_blk + (size_t)u * 128);
        }
        for (int u = tid; u < lo; u += TNT)
          prefetch_l2(g.dy_img + (size_t)tn * P * kbo * a_blk + (size_t)u * 128);
      }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t z0 = smem_u32(sZ), a0 = smem_u32(sAi), y0 = smem_u32(sDY), h0 = smem_u32(sDH);
      for (int mc = 0; mc < g.w_mchunks; ++mc)
        for (int ks = 0; ks < TM / 16; ++ks) {
          const uint32_t za = z0 + (uint32_t)mc * 2u * a_blk + (uint32_t)ks * 2048u;
          const uint32_t hb = h0 + (uint32_t)ks * 2048u;
          umma_bf16(tmem_base + (uint32_t)(mc * n1), make_desc_mn_sw128(za, a_blk),
                    make_desc_mn_sw128(hb, a_blk), idesc_w1, (!first || ks > 0) ? 1u : 0u);
          if (SP) {
            umma_bf16(tmem_base + (uint32_t)(mc * n1), make_desc_mn_sw128(za, a_blk),
                      make_desc_mn_sw128(hb + t2_lo, a_blk), idesc_w1, 1u);
            umma_bf16(tmem_base + (uint32_t)(mc * n1), make_desc_mn_sw128(za + z_lo, a_blk),
                      make_desc_mn_sw128(hb, a_blk), idesc_w1, 1u);
          }
        }
      for (int ks = 0; ks < TM / 16; ++ks) {
        const uint32_t aa = a0 + (uint32_t)ks * 2048u, yb = y0 + (uint32_t)ks * 2048u;
        umma_bf16(tW2, make_desc_mn_sw128(aa, a_blk), make_desc_mn_sw128(yb, a_blk), idesc_w2,
                  (!first || ks > 0) ? 1u : 0u);
        if (SP) {
          umma_bf16(tW2, make_desc_mn_sw128(aa, a_blk), make_desc_mn_sw128(yb + to_lo, a_blk),
                    idesc_w2, 1u);
          umma_bf16(tW2, make_desc_mn_sw128(aa + t2_lo, a_blk), make_desc_mn_sw128(yb, a_blk),
                    idesc_w2, 1u);
        }
      }
      umma_commit(&bars[0]);
    }
    first = false;
    mbar_wait(&bars[0], ph);
    ph ^= 1;
    tc_fence_after();
    __syncthreads();
  }
  if (cur_chunk >= 0) flush(cur_chunk);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)g.w_tmem_cols);
}

static int pow2_cols(int c) {
  int v = 32;
  while (v < c) v *= 2;
  return v;
}

static int make_bgeo(const KParams& p, BGeo& g) {
  const nlam_rowmlp& d = p.d;
  const uint32_t P = p.split ? 2u : 1u;  // operand tiles: hi (+ lo), see rowmlp_tc.cuh put8
  g.parts = (int)P;
  g.n1 = pad_n(d.d_hidden), g.n2 = pad_n(d.d_out);
  g.nmax = g.n1 > g.n2 ? g.n1 : g.n2;
  g.k1 = (p.k_total + 15) / 16 * 16;
  g.k2 = (d.d_hidden + 15) / 16 * 16;
  g.ko = (d.d_out + 15) / 16 * 16;
  g.kb1 = (g.k1 + 63) / 64, g.kb2 = (g.n1 + 63) / 64, g.kbo = (g.n2 + 63) / 64;
  // z blocks gathered per GEMM-1 round (the fast gather works on whole sources: 2 blocks each
  // at d = 128, where the 64 KB staging region takes two sources per round)
  g.rb = (fast_n(p) == 128 && fast_gather(p)) ? (option_rb128() == 2 ? 2 : 4) : 3;
  if (p.split) g.rb = 2;  // hi + lo of two blocks = the 64 KB the fp32 staging tile may need
  g.cY = g.n1, g.cZ = g.n1 + g.nmax;
  g.tmem_cols = pow2_cols(g.cZ + 128);
  const uint32_t blk = TM * 128u;
  const int rb = g.kb1 < g.rb ? g.kb1 : g.rb;
  uint32_t r0 = P * (uint32_t)rb * blk;
  const uint32_t stg_bytes = (uint32_t)TM * (g.n2 > 64 ? g.n2 : 64) * 4u;
  if (stg_bytes > r0) r0 = stg_bytes;
  auto al = [](uint32_t x) { return (x + 1023u) & ~1023u; };
  uint32_t o = al(r0);
  const int kbt = g.kb2 > g.kbo ? g.kb2 : g.kbo;
  g.off_t = o, o += P * (uint32_t)kbt * blk;
  g.off_w1 = o, o += al(P * (uint32_t)g.kb1 * g.n1 * 128u);
  g.off_w2 = o, o += al(P * (uint32_t)g.kb2 * g.n2 * 128u);
  g.off_par = o, o += (uint32_t)(g.n1 + 2 * g.n2) * 4u;
  g.off_lnx = o, o += (uint32_t)TM * 2u * 4u;  // [TM][2] (see row_allsum)
  g.off_bar = o, o += 64;
  g.smem_bytes = o;
  g.tiles_per_batch = n_tiles_of(d, TM);
  g.total_tiles = g.tiles_per_batch * d.batch;
  g.p_total = p.lay.total();
  g.vec_len = d.d_hidden + d.d_out * (p.lay.has_ln ? 3 : 1);
  // wgrad
  g.w_mchunks = (g.kb1 + 1) / 2;
  g.w_tmem_cols = pow2_cols(g.w_mchunks * g.n1 + g.n2);
  o = P * (uint32_t)g.kb1 * blk;
  g.w_off_a = o, o += P * (uint32_t)g.kb2 * blk;
  g.w_off_dy = o, o += P * (uint32_t)g.kbo * blk;
  g.w_off_dh = o, o += P * (uint32_t)g.kb2 * blk;
  o += blk;  // MN-major M=128 views may run one (ignored) block past a tile
  g.w_off_bar = o, o += 64;
  g.w_smem_bytes = o;
  NLAM_CHECK(g.smem_bytes <= 232448 && g.w_smem_bytes <= 232448 && g.w_tmem_cols <= 512 &&
                 g.tmem_cols <= 512,
             "rowmlp_bwd(bf16): %u / %u bytes of shared memory, %d / %d TMEM columns", g.smem_bytes,
             g.w_smem_bytes, g.tmem_cols, g.w_tmem_cols);
  return 0;
}

int make_bgeo_probe(const KParams& p) {
  BGeo g{};
  return make_bgeo(p, g);
}

static int grid_for(uint32_t smem, int tmem_cols, int total_tiles) {
  int per_sm = smem <= 113 * 1024 ? 2 : 1;
  if (tmem_cols * per_sm > 512) per_sm = 1;
  int grid = 148 * per_sm;
  return grid < total_tiles ? grid : total_tiles;
}

// wgrad CTAs keep their accumulators in TMEM across tiles and flush a full
// parameter block at the end: few, long-lived CTAs keep the partial traffic small
static int wgrad_grid(const BGeo& g) {
  int grid = grid_for(g.w_smem_bytes, g.w_tmem_cols, g.total_tiles);
  const int want = (g.total_tiles + 3) / 4;
  if (grid > want) grid = want;
  return grid < 1 ? 1 : grid;
}

// three tiles in flight per SM with shared weights (rowmlp_tc_bwd_mc.cu), when there
// is enough work to fill the machine that way
static bool use_dgrad_mc(const KParams& p, const BGeo& g) {
  // default off: in the full train step the two-CTA kernel is ~1 % faster (B200, MEPS)
  const int env = option_dgrad_mc();
  if (env == 0 || p.split) return false;
  if (!tc_dgrad_mc_supported(p, g)) return false;
  // measured: +4 % on the 3-source edge MLPs, -6 % on the 2-source node MLP
  return env > 0 || (g.total_tiles > 296 && p.d.n_src == 3);
}

// one kernel for input and weight gradients (rowmlp_tc_bwd_fused.cu): square 64-wide MLPs
// `need_dz`: -1 unknown (workspace sizing), else whether a source gradient is requested
static bool use_bwd_fused(const KParams& p, int need_dz) {
  if (option_bwd_fused() == 0 || p.split) return false;
  const int kind = tc_bwd_fused_kind(p);
  return kind == 2 || (kind == 1 && need_dz == 0);
}

struct TcBwdWs {
  size_t a_img, dy_img, dh_img, partial, vec_partial, total;  // float offsets
  int d_slots, w_slots;
};
static TcBwdWs tc_bwd_ws(const KParams& p, const BGeo& g, bool fused) {
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  TcBwdWs w;
  const size_t blk_f = TM * 128 / 4;  // floats per 16 KB block
  size_t o = 0;
  // fused: no bf16 tile images, one partial slot per CTA
  w.a_img = o, o += fused ? 0 : al((size_t)g.total_tiles * g.parts * g.kb2 * blk_f);
  w.dy_img = o, o += fused ? 0 : al((size_t)g.total_tiles * g.parts * g.kbo * blk_f);
  w.dh_img = o, o += fused ? 0 : al((size_t)g.total_tiles * g.parts * g.kb2 * blk_f);
  w.d_slots = fused ? (p.src0_batch_sum ? tc_bwd_fused_grid_bsum(g) : tc_bwd_fused_grid(g))
              : use_dgrad_mc(p, g) ? tc_dgrad_mc_grid(g)
                                   : grid_for(g.smem_bytes, g.tmem_cols, g.total_tiles);
  w.w_slots = fused ? w.d_slots : wgrad_grid(g);
  w.partial = o, o += al((size_t)w.w_slots * p.d.n_chunks * g.p_total);
  w.vec_partial = o, o += al((size_t)w.d_slots * p.d.n_chunks * g.vec_len);
  w.total = o;
  return w;
}

}  // namespace tc

bool tc_rowmlp_bwd_is_fused(const nlam_rowmlp_bwd& bd) {
  KParams p{};
  if (fill_params(bd.fwd, p)) return false;
  p.split = bd.fwd.precision == NLAM_FP32;
  int need_dz = 0;
  for (int s = 0; s < bd.fwd.n_src; ++s) need_dz |= bd.d_src[s] != nullptr;
  return tc::use_bwd_fused(p, need_dz);
}

size_t tc_rowmlp_bwd_workspace(const nlam_rowmlp& d) {
  KParams p{};
  if (fill_params(d, p)) return 0;
  p.split = d.precision == NLAM_FP32;
  tc::BGeo g{};
  if (tc::make_bgeo(p, g)) return 0;
  // the kernel choice may depend on which gradients the run asks for: size for either
  size_t n = 0;
  if (!tc::use_bwd_fused(p, 1)) n = tc::tc_bwd_ws(p, g, false).total;
  if (tc::use_bwd_fused(p, 0)) {
    const size_t m = tc::tc_bwd_ws(p, g, true).total;
    n = m > n ? m : n;
  }
  return n;
}

int tc_rowmlp_bwd(const nlam_rowmlp_bwd& bd, cudaStream_t st) {
  const nlam_rowmlp& d = bd.fwd;
  KParams p{};
  if (fill_params(d, p)) return 1;
  p.split = d.precision == NLAM_FP32;  // fp32 operands: split bf16 tiles, 3 UMMAs per product
  NLAM_CHECK(bd.d_params, "rowmlp_bwd: d_params is NULL");
  if (d.rows == 0) {
    if (!bd.params_accumulate)
      NLAM_CUDA(cudaMemsetAsync(bd.d_params, 0, sizeof(float) * (size_t)d.n_chunks * p.lay.total(), st));
    return 0;
  }
  NLAM_CHECK(bd.g0 || bd.g1, "rowmlp_bwd: no output gradient given");
  NLAM_CHECK(!bd.g1 || bd.g1_idx, "rowmlp_bwd: g1 needs g1_idx");
  tc::BGeo g{};
  if (tc::make_bgeo(p, g)) return 1;
  int want_dz = 0;
  for (int s = 0; s < d.n_src; ++s) want_dz |= bd.d_src[s] != nullptr;
  const bool fused = tc::use_bwd_fused(p, want_dz);
  // sender pre-reduction / batch-summed source-0 gradient: fused 64-wide kernel only
  p.sp_src = bd.sp_src >= 0 && bd.sp_tile_ptr ? bd.sp_src : -1;
  p.n_sp = bd.n_sp, p.sp_tile_ptr = bd.sp_tile_ptr, p.sp_row_ptr = bd.sp_row_ptr, p.sp_rows = bd.sp_rows;
  p.src0_batch_sum = bd.src0_batch_sum;
  if (p.sp_src >= 0 || p.src0_batch_sum) {
    NLAM_CHECK(fused && tc_bwd_fused_kind(p) == 2 && d.tile_ptr && d.agg.tile_seg,
               "rowmlp_bwd: sp_src / src0_batch_sum need the fused 64-wide kernel on "
               "receiver-aligned tiles");
    NLAM_CHECK(p.sp_src < 0 || (p.sp_src == 1 && bd.sp_row_ptr && bd.sp_rows && bd.n_sp > 0 &&
                                bd.d_src[1] && !bd.d_src_idx[1] && !bd.d_src_idx[2] &&
                                bd.reduce_src != 1 && !bd.d_src_bf16[1]),
               "rowmlp_bwd: sp_src must be source 1 with dense, un-scattered partial rows");
    NLAM_CHECK(!p.src0_batch_sum || (d.src[0].batch_stride == 0 && d.batch > 1 && !bd.g0 &&
                                     d.residual_src < 0 && bd.d_src[0] && bd.reduce_src != 0 &&
                                     !bd.d_src_bf16[0]),
               "rowmlp_bwd: src0_batch_sum needs a batch-shared source 0 without residual");
  }
  NLAM_CHECK(!fused || (!bd.d_src_idx[1] && !bd.d_src_idx[2]),
             "rowmlp_bwd: the fused kernel scatters the rows of source 0 only (d_src_idx[1..2])");
  const tc::TcBwdWs ws = tc::tc_bwd_ws(p, g, fused);
  NLAM_CHECK(bd.workspace && bd.workspace_floats >= ws.total,
             "rowmlp_bwd: workspace too small (%zu < %zu floats)", bd.workspace_floats, ws.total);
  NLAM_CHECK(((uintptr_t)bd.workspace) % 16 == 0, "rowmlp_bwd: workspace must be 16B aligned");
  p.g0 = bd.g0, p.g1 = bd.g1, p.g1_idx = bd.g1_idx, p.g1_scale = bd.g1_scale;
  p.g1_batch_stride = bd.g1_batch_stride;
  p.g0_idx = bd.g0_idx;
  p.reduce_src = bd.reduce_src, p.reduce_accumulate = bd.reduce_accumulate;
  p.inputs_stable = (bd.stage_mask & 16) ? 1 : 0;
  p.g0_nsum = bd.g0_sum_count > 1 ? bd.g0_sum_count : 1;
  p.g0_sum_stride = bd.g0_sum_stride;
  NLAM_CHECK(p.g0_nsum == 1 || (fused && d.batch == 1 && d.d_out == 64 && bd.g0 && !bd.g0_idx &&
                                   d.residual_src < 0),
             "rowmlp_bwd: g0_sum_count needs the fused backward kernel, batch 1, dense g0 rows");
  {
    bool special = bd.g0_idx || bd.reduce_src >= 0;
    for (int s = 0; s < NLAM_MAX_SRC; ++s) {
      p.d_src_idx[s] = bd.d_src_idx[s];
      special |= bd.d_src_idx[s] != nullptr;
    }
    NLAM_CHECK(!special || tc::fast_gather(p),
               "rowmlp_bwd: row scatter / fused reduction need the square fast path "
               "(d_hidden == d_out == source widths in {64, 128})");
    NLAM_CHECK(bd.reduce_src < 0 || (d.agg.seg_ptr && d.agg.tile_seg && d.tile_ptr),
               "rowmlp_bwd: reduce_src needs fwd.agg.seg_ptr / tile_seg and a tile table");
  }
  g.need_dz = 0;
  for (int s = 0; s < NLAM_MAX_SRC; ++s) {
    p.d_src[s] = s < d.n_src ? bd.d_src[s] : nullptr;
    p.d_src_bf16[s] = s < d.n_src ? bd.d_src_bf16[s] : nullptr;
    if (p.d_src[s]) g.need_dz = 1;
  }
  g.a_img = reinterpret_cast<uint8_t*>(bd.workspace + ws.a_img);
  g.dy_img = reinterpret_cast<uint8_t*>(bd.workspace + ws.dy_img);
  g.dh_img = reinterpret_cast<uint8_t*>(bd.workspace + ws.dh_img);
  g.partial = bd.workspace + ws.partial;
  g.vec_partial = bd.workspace + ws.vec_partial;
  if (d.n_chunks > 1 && (!(bd.stage_mask & 7) || (bd.stage_mask & 1)))  // CTAs write visited chunks only
    NLAM_CUDA(cudaMemsetAsync(g.partial, 0,
                              sizeof(float) * (ws.vec_partial + (size_t)ws.d_slots * d.n_chunks *
                                                                    g.vec_len - ws.partial), st));
  const int fn = tc::fast_n(p);
  const int gd = ws.d_slots, gw = ws.w_slots;
  auto launch = [&](auto kern, int threads, int grid, uint32_t smem) -> int {
    NLAM_CUDA(ensure_dyn_smem((const void*)kern, (int)smem));
    kern<<<grid, threads, smem, st>>>(p, g);
    NLAM_CUDA(cudaGetLastError());
    count_launch();
    return 0;
  };
  const bool fg = tc::fast_gather(p);
  // d = 128: one CTA per SM anyway (shared memory) -> 512 threads = 16 warps, 32 columns of a
  // row per thread like the d = 64 kernels (option "wide128" = 0: 256 threads)
  const bool wide = (fn == 128 || (fn == 0 && g.n1 == 128)) && option_wide128() != 0 &&
                    g.smem_bytes > 113 * 1024 && g.w_smem_bytes > 113 * 1024;
  int rc;
  const int mask = (bd.stage_mask & 7) ? bd.stage_mask : (7 | (bd.stage_mask & 8));
  const bool dmc = tc::use_dgrad_mc(p, g);
  if (fused) {  // stage bit 1 covers input AND weight gradients
    if ((mask & 1) && tc_rowmlp_bwd_fused(p, g, st)) return 1;
    if (!(mask & 4)) return 0;
    return launch_reduce_params(g.partial, ws.w_slots, d.n_chunks, g.p_total, bd.d_params,
                                bd.params_accumulate, g.vec_partial, ws.d_slots, g.vec_len, p.lay,
                                st, (mask & 8) != 0);
  }
#define NLAM_BWD_PAIR(FNV, FGV, TN)                                                         \
  rc = 0;                                                                                   \
  if (mask & 1)                                                                             \
    rc = dmc ? tc_rowmlp_dgrad_mc(p, g, st)                                                 \
             : launch(tc::rowmlp_tc_dgrad_kernel<FNV, FGV, TN>, TN, gd, g.smem_bytes);      \
  if (!rc && (mask & 2))                                                                    \
    rc = launch(tc::rowmlp_tc_wgrad_kernel<FNV, FGV, TN>, TN, gw, g.w_smem_bytes);
  if (p.split) {
    rc = 0;
  } else if (fn == 64 && fg) {
    NLAM_BWD_PAIR(64, true, 256)
  } else if (fn == 64) {
    NLAM_BWD_PAIR(64, false, 256)
  } else if (fn == 128 && fg && wide) {
    NLAM_BWD_PAIR(128, true, 512)
  } else if (fn == 128 && fg) {
    NLAM_BWD_PAIR(128, true, 256)
  } else if (fn == 128 && wide) {
    NLAM_BWD_PAIR(128, false, 512)
  } else if (fn == 128) {
    NLAM_BWD_PAIR(128, false, 256)
  } else if (wide) {
    NLAM_BWD_PAIR(0, false, 512)
  } else {
    NLAM_BWD_PAIR(0, false, 256)
  }
#undef NLAM_BWD_PAIR
#define NLAM_BWD_PAIR_SP(FNV, FGV, TN)                                                          \
  rc = 0;                                                                                       \
  if (mask & 1) rc = launch(tc::rowmlp_tc_dgrad_kernel<FNV, FGV, TN, true>, TN, gd, g.smem_bytes); \
  if (!rc && (mask & 2))                                                                        \
    rc = launch(tc::rowmlp_tc_wgrad_kernel<FNV, FGV, TN, true>, TN, gw, g.w_smem_bytes);
  if (p.split) {  // same thread count for both kernels (the column-sum layout depends on it)
    const bool w = g.smem_bytes > 113 * 1024 && g.w_smem_bytes > 113 * 1024;
    if (fn == 64 && fg && w) {
      NLAM_BWD_PAIR_SP(64, true, 512)
    } else if (fn == 64 && fg) {
      NLAM_BWD_PAIR_SP(64, true, 256)
    } else if (fn == 64 && w) {
      NLAM_BWD_PAIR_SP(64, false, 512)
    } else if (fn == 64) {
      NLAM_BWD_PAIR_SP(64, false, 256)
    } else if (w) {
      NLAM_BWD_PAIR_SP(0, false, 512)
    } else {
      NLAM_BWD_PAIR_SP(0, false, 256)
    }
  }
#undef NLAM_BWD_PAIR_SP
  if (rc) return rc;
  if (!(mask & 4)) return 0;
  return launch_reduce_params(g.partial, ws.w_slots, d.n_chunks, g.p_total, bd.d_params,
                              bd.params_accumulate, g.vec_partial, ws.d_slots, g.vec_len, p.lay, st,
                              (mask & 8) != 0);
}

}  // namespace nlam
