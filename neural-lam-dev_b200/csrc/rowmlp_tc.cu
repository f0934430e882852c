// bf16 tensor-core (tcgen05) row-MLP forward for sm_100a.
//
//   out[b,r,:] = [src_0 +] LN( W2 . SiLU( W1 . concat_s src_s[b, idx_s[r], :] + b1 ) + b2 )
//
// A persistent CTA (256 threads) loops over tiles of 128 rows:
//   gather   fp32 rows (128-bit coalesced loads, per-source row index) -> bf16,
//            written straight into the UMMA A-operand layout (K-major, 128B swizzle)
//   GEMM 1   tcgen05.mma M=128 x N=d_hidden x K, accumulator in TMEM  (one thread issues)
//   epi 1    tcgen05.ld -> +b1 -> SiLU -> bf16 -> A operand of GEMM 2 (shared memory)
//   GEMM 2   tcgen05.mma into a second TMEM accumulator
//   epi 2    tcgen05.ld -> +b2 -> LayerNorm (thread owns a half row; two-pass,
//            partial sums exchanged through shared memory) -> fp32 staging tile
//   store    coalesced 128-bit row stores (+ fp32 residual re-read from global)
// W1/W2 are converted to bf16 once per CTA and stay resident in shared memory;
// the hidden activations never touch HBM.  Two CTAs per SM (d <= 64) overlap
// one tile's gather/store with the other's MMA/epilogue.
//
// Reference semantics: utils.make_mlp (utils.py:191-214), InteractionNet.message /
// aggr_mlp (interaction_net.py:106,117-121), SplitMLPs (:134-163).
#include <stdlib.h>

#include "rowmlp_tc.cuh"

namespace nlam {
namespace tc {


// TNT = threads per CTA: 256 (two CTAs per SM at d <= 64), or 512 for d = 128, where shared
// memory allows one CTA per SM only -- 16 warps instead of 8, and 32 instead of 64 columns of
// a row per thread (NG = TNT / 128 column groups per row).
// SP: fp32 operands as split bf16 tiles, three UMMAs per product (rowmlp_tc.cuh, put8).
template <int FN, bool FG, int TNT, bool SP = false>
__global__ void __launch_bounds__(TNT, TNT == 256 ? 2 : 1)
rowmlp_tc_fwd_kernel(const __grid_constant__ KParams p, const __grid_constant__ Geo g) {
  constexpr bool F = FN > 0;  // square fast path: sizes are compile-time constants
  constexpr int NG = TNT / 128;  // column groups (threads) per tile row
  const int n1 = F ? FN : g.n1, n2 = F ? FN : g.n2, k2 = F ? FN : g.k2;
  extern __shared__ __align__(1024) uint8_t sm[];
  if (smem_u32(sm) & 1023u) __trap();  // SWIZZLE_128B operands need 1024-byte alignment
  // region 0 is time-shared: A operand of GEMM 1 -> A operand of GEMM 2 ->
  // fp32 staging tile + LayerNorm exchange buffer
  uint8_t* sA = sm;
  uint8_t* sA2 = sm;
  uint8_t* sW1 = sm + g.off_w1;
  uint8_t* sW2 = sm + g.off_w2;
  float* sPar = reinterpret_cast<float*>(sm + g.off_par);  // b1[n1] b2[n2] gamma[n2] beta[n2]
  float* sLnx = reinterpret_cast<float*>(sm + g.off_lnx);  // [2][TM][2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + g.off_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  float* stg = reinterpret_cast<float*>(sA);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dh = F ? FN : p.d.d_hidden, dout = F ? FN : p.d.d_out;

  if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)g.tmem_cols);
  if (tid == 32) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tH = tmem_base, tY = tmem_base + (uint32_t)n1;
  uint32_t ph0 = 0, ph1 = 0;
  int loaded_chunk = -1;
  if (p.d.n_chunks == 1) {  // single weight set: staged while the previous kernel drains
    stage_weight<TNT, SP>(p.d.w.w1, dh, p.k_total, n1, g.k1, sW1);
    stage_weight<TNT, SP>(p.d.w.w2, dout, dh, n2, k2, sW2);
    stage_params<TNT>(p.d, 0, n1, n2, sPar);
    loaded_chunk = 0;
  }
  pdl_wait();
  // all CTAs of this persistent grid are resident: let the next kernel's CTAs take over
  // each SM (and run their prologue) as soon as this kernel's CTA there exits
  pdl_trigger();

  const uint32_t idesc1 = make_idesc_bf16(TM, n1);
  const uint32_t idesc2 = make_idesc_bf16(TM, n2);
  const uint32_t a_blk = TM * 128u;
  // split mode: byte offsets of the lo tiles behind their hi tiles
  const uint32_t a_lo = (uint32_t)g.kb1 * a_blk, a2_lo = (uint32_t)g.kb2 * a_blk;
  const uint32_t w1_lo = (uint32_t)g.kb1 * (uint32_t)n1 * 128u, w2_lo = (uint32_t)g.kb2 * (uint32_t)n2 * 128u;

  // epilogue ownership: TMEM lane quarter q, row r, column half hf
  const int q = warp & 3, hf = warp >> 2, r = q * 32 + lane;
  // columns per thread: a row is split over the NG column groups when every group gets at
  // least one 16-column chunk, else group 0 takes the whole row
  const int cp1 = n1 >= 16 * NG ? n1 / NG : n1, cp2 = n2 >= 16 * NG ? n2 / NG : n2;
  const bool act1 = n1 >= 16 * NG || hf == 0, act2 = n2 >= 16 * NG || hf == 0;
  const bool split2 = n2 >= 16 * NG;
  auto lnx_sum = [&](const float* base) {  // sum of the row's partial values (NG or 1)
    float s = base[0];
    if (split2)
#pragma unroll
      for (int h = 1; h < NG; ++h) s += base[h];
    return s;
  };
  const uint32_t lane_addr = (uint32_t)(q * 32) << 16;

  // pipelined gather (64-wide sources): row indices are fetched one tile ahead
  constexpr bool PIPE = false;  // no gain measured in the two-CTA forward kernel
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

    if (chunk != loaded_chunk) {  // (re)load the weight set of this chunk
      stage_weight<TNT, SP>(p.d.w.w1 + (size_t)chunk * dh * p.k_total, dh, p.k_total, n1, g.k1, sW1);
      stage_weight<TNT, SP>(p.d.w.w2 + (size_t)chunk * dout * dh, dout, dh, n2, k2, sW2);
      stage_params<TNT>(p.d, chunk, n1, n2, sPar);
      loaded_chunk = chunk;
    }

    // ---------------- gather: fp32 rows -> bf16 A operand
    if (PIPE) {
      const int cidx[NLAM_MAX_SRC] = {nidx[0], nidx[1], nidx[2]};
      gather_rows_pipe<64, TNT>(p, b, cidx, sA, tid);
    } else if (F && FG) {
      gather_rows_fast<(F ? FN : 64), TNT, SP>(p, b, row0, cnt, 0, p.d.n_src, sA, a_lo);
    } else {
      gather_rows<TNT, SP>(p, b, row0, cnt, 0, g.k1, sA, a_lo);
    }
    {  // next tile: row indices (pipelined gather) and L2 prefetch of its input rows
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
      }
    }
    fence_async_smem();
    __syncthreads();

    // ---------------- GEMM 1: H = A . W1^T
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a0 = smem_u32(sA), w0 = smem_u32(sW1);
      const uint32_t w_blk = (uint32_t)n1 * 128u;
      for (int ks = 0; ks < g.k1 / 16; ++ks) {
        const uint32_t kb = ks >> 2, kin = (ks & 3) * 32;
        umma_bf16(tH, make_desc_k_sw128(a0 + kb * a_blk + kin),
                  make_desc_k_sw128(w0 + kb * w_blk + kin), idesc1, ks > 0);
        if (SP) {
          umma_bf16(tH, make_desc_k_sw128(a0 + kb * a_blk + kin),
                    make_desc_k_sw128(w0 + w1_lo + kb * w_blk + kin), idesc1, 1u);
          umma_bf16(tH, make_desc_k_sw128(a0 + a_lo + kb * a_blk + kin),
                    make_desc_k_sw128(w0 + kb * w_blk + kin), idesc1, 1u);
        }
      }
      umma_commit(&bars[0]);
    }
    mbar_wait(&bars[0], ph0);
    ph0 ^= 1;
    tc_fence_after();

    // ---------------- epilogue 1: a = SiLU(H + b1) -> bf16 A2
    if (act1) {
      for (int cc = 0; cc < cp1; cc += 16) {
        const int c0 = hf * cp1 + cc;
        float v[16];
        tmem_ld16(tH + lane_addr + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = silu_sel<SP>(v[j] + sPar[c0 + j]);
        put16<SP>(sA2, a2_lo, r, c0, a_blk, v);
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();

    // ---------------- GEMM 2: Y = A2 . W2^T
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a0 = smem_u32(sA2), w0 = smem_u32(sW2);
      const uint32_t w_blk = (uint32_t)n2 * 128u;
      for (int ks = 0; ks < k2 / 16; ++ks) {
        const uint32_t kb = ks >> 2, kin = (ks & 3) * 32;
        umma_bf16(tY, make_desc_k_sw128(a0 + kb * a_blk + kin),
                  make_desc_k_sw128(w0 + kb * w_blk + kin), idesc2, ks > 0);
        if (SP) {
          umma_bf16(tY, make_desc_k_sw128(a0 + kb * a_blk + kin),
                    make_desc_k_sw128(w0 + w2_lo + kb * w_blk + kin), idesc2, 1u);
          umma_bf16(tY, make_desc_k_sw128(a0 + a2_lo + kb * a_blk + kin),
                    make_desc_k_sw128(w0 + kb * w_blk + kin), idesc2, 1u);
        }
      }
      umma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], ph1);
    ph1 ^= 1;
    tc_fence_after();

    // ---------------- epilogue 2: y + b2 -> LayerNorm -> fp32 staging tile
    const float* sB2 = sPar + n1;
    const float* sG = sB2 + n2;
    const float* sBe = sG + n2;
    float mean = 0.f, rstd = 1.f;
    if (p.d.w.ln_g) {
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
      sLnx[(0 * TM + r) * NG + hf] = s;
      __syncthreads();
      mean = lnx_sum(&sLnx[(0 * TM + r) * NG]) / (float)dout;
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
      sLnx[(1 * TM + r) * NG + hf] = qq;
      __syncthreads();
      const float var = lnx_sum(&sLnx[(1 * TM + r) * NG]) / (float)dout;
      rstd = rsqrtf(var + LN_EPS);
    }
    if (act2)
      for (int cc = 0; cc < cp2; cc += 16) {
        const int c0 = hf * cp2 + cc;
        float v[16];
        tmem_ld16(tY + lane_addr + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float y = v[j] + sB2[c0 + j];
          if (p.d.w.ln_g) y = (y - mean) * rstd * sG[c0 + j] + sBe[c0 + j];
          v[j] = y;
        }
        float4* dst = reinterpret_cast<float4*>(stg + (size_t)r * g.stg_ld + c0);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4)
          dst[j4] = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
      }
    tc_fence_before();
    __syncthreads();

    // ---------------- coalesced store (+ fp32 residual / second output src_0 + out),
    // optional row scatter (out_idx) and fused segment reduction (agg)
    if (p.d.out || p.d.out_res) {
      float* out = p.d.out ? p.d.out + (size_t)b * p.d.rows * dout : nullptr;
      float* out2 = p.d.out_res ? p.d.out_res + (size_t)b * p.d.rows * dout : nullptr;
      const int32_t* oidx = p.d.out_idx;
      const nlam_src& s0 = p.d.src[0];
      const bool res = p.d.residual_src == 0;
      const bool need0 = res || out2;
      if (p.out_vec_ok && (!need0 || p.vec_ok[0])) {
        const int w4 = dout >> 2;
        for (int u = tid; u < cnt * w4; u += TNT) {
          const int row = u / w4, c4 = u % w4;
          float4 v = *reinterpret_cast<const float4*>(stg + (size_t)row * g.stg_ld + c4 * 4);
          float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
          if (need0) {
            const int ridx = s0.idx ? __ldg(s0.idx + row0 + row) : row0 + row;
            e = __ldg(reinterpret_cast<const float4*>(s0.ptr + (long long)b * s0.batch_stride +
                                                      (long long)ridx * s0.ld) + c4);
          }
          if (res) v.x += e.x, v.y += e.y, v.z += e.z, v.w += e.w;
          const size_t orow = oidx ? (size_t)__ldg(oidx + row0 + row) : (size_t)(row0 + row);
          if (out) *reinterpret_cast<float4*>(out + orow * dout + c4 * 4) = v;
          if (p.d.out_bf16)
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.d.out_bf16) +
                                      ((size_t)b * p.d.rows + orow) * dout + c4 * 4) =
                make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
          if (out2)
            *reinterpret_cast<float4*>(out2 + orow * dout + c4 * 4) =
                make_float4(v.x + e.x, v.y + e.y, v.z + e.z, v.w + e.w);
          if (p.d.out_res_bf16)
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.d.out_res_bf16) +
                                      ((size_t)b * p.d.rows + orow) * dout + c4 * 4) =
                make_uint2(pack_bf16(v.x + e.x, v.y + e.y), pack_bf16(v.z + e.z, v.w + e.w));
        }
      } else {
        for (int u = tid; u < cnt * dout; u += TNT) {
          const int row = u / dout, c = u % dout;
          float v = stg[(size_t)row * g.stg_ld + c];
          float e = 0.f;
          if (need0) {
            const int ridx = s0.idx ? __ldg(s0.idx + row0 + row) : row0 + row;
            e = __ldg(s0.ptr + (long long)b * s0.batch_stride + (long long)ridx * s0.ld + c);
          }
          if (res) v += e;
          const size_t orow = oidx ? (size_t)__ldg(oidx + row0 + row) : (size_t)(row0 + row);
          if (out) out[orow * dout + c] = v;
          if (out2) out2[orow * dout + c] = v + e;
        }
      }
    }
    if (p.d.agg.out || p.d.agg.out_bf16) {  // receiver-aligned tile: every segment is complete
      const int seg_lo = __ldg(p.d.agg.tile_seg + tile), seg_hi = __ldg(p.d.agg.tile_seg + tile + 1);
      const int w4 = dout >> 2;
      float* ao = p.d.agg.out + (size_t)b * p.d.agg.n_seg * dout;
      for (int u = tid; u < (seg_hi - seg_lo) * w4; u += TNT) {
        const int seg = seg_lo + u / w4, c4 = u % w4;
        const int r0 = __ldg(p.d.agg.seg_ptr + seg) - row0, r1 = __ldg(p.d.agg.seg_ptr + seg + 1) - row0;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int rr = r0; rr < r1; ++rr) {
          const float4 v = *reinterpret_cast<const float4*>(stg + (size_t)rr * g.stg_ld + c4 * 4);
          acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
        }
        if (p.d.agg.scale) {
          const float sc = __ldg(p.d.agg.scale + seg);
          acc.x *= sc, acc.y *= sc, acc.z *= sc, acc.w *= sc;
        }
        if (p.d.agg.out) *reinterpret_cast<float4*>(ao + (size_t)seg * dout + c4 * 4) = acc;
        if (p.d.agg.out_bf16)
          *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.d.agg.out_bf16) +
                                    ((size_t)b * p.d.agg.n_seg + seg) * dout + c4 * 4) =
              make_uint2(pack_bf16(acc.x, acc.y), pack_bf16(acc.z, acc.w));
      }
    }
    __syncthreads();  // staging (aliases A) is free again
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
}


// Can the tensor-core path take this problem?  (else: fp32 FFMA path)
bool tc_supported(const nlam_rowmlp& d) {
  if (pad_n(d.d_hidden) < 0 || pad_n(d.d_out) < 0) return false;
  int k = 0;
  for (int s = 0; s < d.n_src; ++s) {
    k += d.src[s].width;
  }
  return k <= 384;
}

int make_geo(const KParams& p, Geo& g);
int make_bgeo_probe(const KParams& p);  // rowmlp_tc_bwd.cu

// fp32 on the tensor cores (split operands): the doubled tiles must fit shared memory in all
// three kernels, else the FFMA path takes the problem
bool tc_split_supported(const nlam_rowmlp& d) {
  if (option_fp32_split() == 0 || !tc_supported(d)) return false;
  KParams p{};
  if (fill_params(d, p)) return false;
  p.split = 1;
  Geo g{};
  return make_geo(p, g) == 0 && make_bgeo_probe(p) == 0;
}

int make_geo(const KParams& p, Geo& g) {
  const nlam_rowmlp& d = p.d;
  const uint32_t P = p.split ? 2u : 1u;  // operand tiles: hi (+ lo)
  g.parts = (int)P;
  g.n1 = pad_n(d.d_hidden), g.n2 = pad_n(d.d_out);
  g.k1 = (p.k_total + 15) / 16 * 16;
  g.k2 = (d.d_hidden + 15) / 16 * 16;
  g.kb1 = (g.k1 + 63) / 64, g.kb2 = (g.k2 + 63) / 64;
  int cols = g.n1 + g.n2;
  g.tmem_cols = 32;
  while (g.tmem_cols < cols) g.tmem_cols *= 2;
  g.stg_ld = g.n2 + 4;
  const uint32_t a_bytes = P * (uint32_t)g.kb1 * TM * 128u;
  const uint32_t a2_bytes = P * (uint32_t)g.kb2 * TM * 128u;
  const uint32_t stg_bytes = (uint32_t)TM * g.stg_ld * 4u;
  const uint32_t lnx_bytes = 2u * TM * 4u * 4u;  // [2][TM][<= 4 column groups]
  uint32_t r0 = a_bytes > a2_bytes ? a_bytes : a2_bytes;
  if (stg_bytes + lnx_bytes > r0) r0 = stg_bytes + lnx_bytes;
  auto al = [](uint32_t x) { return (x + 1023u) & ~1023u; };
  g.off_lnx = stg_bytes;
  uint32_t o = al(r0);
  g.off_w1 = o, o += al(P * (uint32_t)g.kb1 * g.n1 * 128u);
  g.off_w2 = o, o += al(P * (uint32_t)g.kb2 * g.n2 * 128u);
  g.off_par = o, o += (uint32_t)(g.n1 + 3 * g.n2) * 4u;
  g.off_bar = o, o += 64;
  g.smem_bytes = o;
  g.tiles_per_batch = n_tiles_of(d, TM);
  g.total_tiles = g.tiles_per_batch * d.batch;
  NLAM_CHECK(g.smem_bytes <= 232448, "rowmlp(bf16): needs %u bytes of shared memory", g.smem_bytes);
  return 0;
}

}  // namespace tc

int tc_rowmlp_fwd(const nlam_rowmlp& d, cudaStream_t st) {
  KParams p{};
  if (fill_params(d, p)) return 1;
  if (d.rows == 0) return 0;
  NLAM_CHECK(d.out || d.out_res || d.agg.out || d.agg.out_bf16, "rowmlp: no output requested");
  NLAM_CHECK(!(d.agg.out || d.agg.out_bf16) || d.d_out % 4 == 0, "rowmlp: agg needs d_out %% 4 == 0");
  NLAM_CHECK(!(d.out_bf16 || d.out_res_bf16 || d.agg.out_bf16) || (p.out_vec_ok && d.d_out % 4 == 0),
             "rowmlp: bf16 shadow outputs need 16-byte aligned fp32 outputs and d_out %% 4 == 0");
  NLAM_CHECK(!d.out_bf16 || d.out, "rowmlp: out_bf16 needs out");
  NLAM_CHECK(!d.out_res_bf16 || d.out_res, "rowmlp: out_res_bf16 needs out_res");
  p.split = d.precision == NLAM_FP32;  // fp32 operands: split bf16 tiles, 3 UMMAs per product
  tc::Geo g{};
  if (tc::make_geo(p, g)) return 1;
  if (!p.split) {  // four tiles in flight per SM with shared weights, when there is enough work
    const int mc_env = option_fwd_mc();
    // measured on MEPS shapes: +16 % on the 3-source edge MLPs (gather-latency bound),
    // -12 % on the 2-source node MLP, so only the former take this path by default
    const bool want = mc_env < 0 ? (g.total_tiles > 296 && d.n_src == 3) : mc_env != 0;
    if (want && option_tma() != 0 && tc_fwd_tma_supported(p)) return tc_rowmlp_fwd_tma(p, g, st);
    if (want && tc_fwd_mc_supported(p)) return tc_rowmlp_fwd_mc(p, g, st);
  }
  int per_sm = g.smem_bytes <= 113 * 1024 ? 2 : 1;
  if (g.tmem_cols * per_sm > 512) per_sm = 1;
  int grid = 148 * per_sm;
  if (grid > g.total_tiles) grid = g.total_tiles;
  const int fn = tc::fast_n(p);
  auto launch = [&](auto kern, int threads) -> int {
    NLAM_CUDA(ensure_dyn_smem((const void*)kern, (int)g.smem_bytes));
    NLAM_CUDA(launch_k(kern, grid, threads, g.smem_bytes, st, p, g));
    return 0;
  };
  const bool fg = tc::fast_gather(p);
  // d = 128: one CTA per SM (shared memory) -> 512 threads (option "wide128" = 0: 256)
  const bool wide = option_wide128() != 0 && per_sm == 1;
  if (p.split) {  // the hi/lo tiles of a d = 64 edge MLP take 160 KB: one wide CTA per SM
    NLAM_CHECK(!d.out_bf16 && !d.out_res_bf16 && !d.agg.out_bf16, "rowmlp(fp32): no bf16 shadow outputs");
    const bool w = per_sm == 1;
    int rc = fn == 64 ? (fg ? (w ? launch(tc::rowmlp_tc_fwd_kernel<64, true, 512, true>, 512)
                                 : launch(tc::rowmlp_tc_fwd_kernel<64, true, 256, true>, 256))
                            : (w ? launch(tc::rowmlp_tc_fwd_kernel<64, false, 512, true>, 512)
                                 : launch(tc::rowmlp_tc_fwd_kernel<64, false, 256, true>, 256)))
                      : (w ? launch(tc::rowmlp_tc_fwd_kernel<0, false, 512, true>, 512)
                           : launch(tc::rowmlp_tc_fwd_kernel<0, false, 256, true>, 256));
    if (rc) return rc;
    NLAM_CUDA(cudaGetLastError());
    count_launch();
    return 0;
  }
  // launches that cannot fill the SMs anyway (levels >= 1 of the hierarchical meshes): 16
  // warps per tile, 16 columns of a row per thread -> shorter epilogues on the one tile an
  // SM gets (option "small512")
  const bool small = option_small512() != 0 && g.total_tiles <= 148;
  int rc = fn == 64    ? (small ? (fg ? launch(tc::rowmlp_tc_fwd_kernel<64, true, 512>, 512)
                                      : launch(tc::rowmlp_tc_fwd_kernel<64, false, 512>, 512))
                          : fg  ? launch(tc::rowmlp_tc_fwd_kernel<64, true, 256>, 256)
                                : launch(tc::rowmlp_tc_fwd_kernel<64, false, 256>, 256))
           : fn == 128 ? (wide ? (fg ? launch(tc::rowmlp_tc_fwd_kernel<128, true, 512>, 512)
                                     : launch(tc::rowmlp_tc_fwd_kernel<128, false, 512>, 512))
                               : (fg ? launch(tc::rowmlp_tc_fwd_kernel<128, true, 256>, 256)
                                     : launch(tc::rowmlp_tc_fwd_kernel<128, false, 256>, 256)))
           // non-square shapes with a 128-wide hidden layer (d = 128 output map): one CTA per SM too
           : (wide && g.n1 == 128) ? launch(tc::rowmlp_tc_fwd_kernel<0, false, 512>, 512)
                                   : launch(tc::rowmlp_tc_fwd_kernel<0, false, 256>, 256);
  if (rc) return rc;
  NLAM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace nlam
