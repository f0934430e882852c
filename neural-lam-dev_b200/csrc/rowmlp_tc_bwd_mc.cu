// Multi-context bf16 tcgen05 row-MLP input-gradient kernel (d_hidden = d_out =
// source widths = 64, one weight set).
//
// One persistent CTA per SM, 384 threads = 3 warpgroups.  W1 / W2 (bf16 UMMA
// operands, 32 KB) are staged once per SM and shared; every warpgroup is an
// independent tile pipeline with its own 64 KB of shared memory (48 KB operand /
// staging region + 16 KB a -> dY -> dH tile), 128 TMEM columns, three mbarriers
// and a named barrier: three 128-row tiles are in flight per SM, each thread owns
// a whole row (LayerNorm backward without cross-thread exchange) and has up to
// 168 registers, which pays for the software-pipelined gather with row indices
// fetched one tile ahead.  TMEM columns are recycled: H -> dZ buffer 0,
// Y -> dA -> dZ buffer 1.
//
// Math, outputs, scratch images and partial layout are those of
// rowmlp_tc_dgrad_kernel (rowmlp_tc_bwd.cu); the weight-gradient kernel is shared.
#include "rowmlp_tc_bwd.cuh"

namespace nlam {
namespace tc {

constexpr int DM_WG = 3;
constexpr int DM_NT = 128 * DM_WG;
constexpr int DM_FN = 64;
constexpr uint32_t DM_R0 = 3u * TM * 128u;             // 48 KB
constexpr uint32_t DM_CTX = DM_R0 + TM * 128u;         // + 16 KB tile = 64 KB per context
constexpr uint32_t DM_OFF_W1 = DM_WG * DM_CTX;
constexpr uint32_t DM_OFF_W2 = DM_OFF_W1 + 3u * DM_FN * 128u;
constexpr uint32_t DM_OFF_PAR = DM_OFF_W2 + DM_FN * 128u;
constexpr uint32_t DM_OFF_BAR = DM_OFF_PAR + 3u * DM_FN * 4u;
constexpr uint32_t DM_SMEM = DM_OFF_BAR + 128u;

__device__ __forceinline__ void dm_wg_sync(int wg) {
  asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory");
}

__device__ __forceinline__ void copy_tile_out_wg(const uint8_t* s, uint8_t* g, int wtid) {
  // one 16 KB block, 128 threads: 8 x 128-bit per thread
  const uint4* src = reinterpret_cast<const uint4*>(s);
  uint4* dst = reinterpret_cast<uint4*>(g);
  uint4 v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = src[wtid + j * 128];
#pragma unroll
  for (int j = 0; j < 8; ++j) dst[wtid + j * 128] = v[j];
}

__global__ void __launch_bounds__(DM_NT, 1)
rowmlp_tc_dgrad_mc_kernel(const __grid_constant__ KParams p, const __grid_constant__ BGeo g) {
  extern __shared__ __align__(1024) uint8_t sm[];
  if (smem_u32(sm) & 1023u) __trap();
  constexpr int FN = DM_FN;
  const int tid = threadIdx.x, wg = tid >> 7, wtid = tid & 127;
  const int warp = tid >> 5, lane = tid & 31;
  uint8_t* sA = sm + (uint32_t)wg * DM_CTX;   // z blocks | fp32 staging
  float* stg = reinterpret_cast<float*>(sA);
  uint8_t* sT = sA + DM_R0;                    // a -> dY -> dH bf16 tile
  uint8_t* sW1 = sm + DM_OFF_W1;
  uint8_t* sW2 = sm + DM_OFF_W2;
  float* sPar = reinterpret_cast<float*>(sm + DM_OFF_PAR);  // b1 | b2 | gamma
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + DM_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * DM_WG);
  const bool has_ln = p.d.w.ln_g != nullptr;

  if (warp == 0) tmem_alloc(tmem_slot, 512u);
  if (tid == 32) {
    for (int i = 0; i < 3 * DM_WG; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  {  // weights: once per SM
    const int nch1 = (p.d.n_src * FN) >> 3;
    for (int u = tid; u < FN * nch1; u += DM_NT) {
      const int n = u / nch1, k0 = (u % nch1) * 8;
      const float4* q = reinterpret_cast<const float4*>(p.d.w.w1 + (size_t)n * p.k_total + k0);
      const float4 a = __ldg(q), c = __ldg(q + 1);
      *reinterpret_cast<uint4*>(sW1 + sw128_off(n, k0, FN * 128u)) =
          make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(c.x, c.y),
                     pack_bf16(c.z, c.w));
    }
    for (int u = tid; u < FN * (FN >> 3); u += DM_NT) {
      const int n = u / (FN >> 3), k0 = (u % (FN >> 3)) * 8;
      const float4* q = reinterpret_cast<const float4*>(p.d.w.w2 + (size_t)n * FN + k0);
      const float4 a = __ldg(q), c = __ldg(q + 1);
      *reinterpret_cast<uint4*>(sW2 + sw128_off(n, k0, FN * 128u)) =
          make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(c.x, c.y),
                     pack_bf16(c.z, c.w));
    }
    for (int i = tid; i < 3 * FN; i += DM_NT) {
      const int j = i % FN, which = i / FN;
      sPar[i] = which == 0 ? __ldg(p.d.w.b1 + j)
                : which == 1 ? __ldg(p.d.w.b2 + j)
                             : (has_ln ? __ldg(p.d.w.ln_g + j) : 1.f);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tH = tmem_base + (uint32_t)wg * 128u, tY = tH + 64u;  // also dZ buffers 0 / 1
  uint64_t* bar_m = &bars[3 * wg];
  uint64_t* bar_z = &bars[3 * wg + 1];  // [0], [1]
  uint32_t ph_m = 0, ph_z = 0;

  const uint32_t a_blk = TM * 128u;
  const uint32_t idesc = make_idesc_bf16(TM, FN);
  const uint32_t idesc_mn = make_idesc_bf16(TM, FN, 0, 1);  // B operand viewed MN-major
  const int r = wtid;
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
  const float* sB1 = sPar;
  const float* sB2 = sPar + FN;
  const float* sG = sPar + 2 * FN;
  const int n_src = p.d.n_src;
  const int k1steps = n_src * FN / 16;
  const int stride = gridDim.x * DM_WG;

  // column-sum accumulators: lane l owns column 16*i + (l & 15) of its warp's 32 rows
  float acc_db1[4], acc_db2[4], acc_dg[4], acc_dbt[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc_db1[i] = acc_db2[i] = acc_dg[i] = acc_dbt[i] = 0.f;

  int nidx[NLAM_MAX_SRC] = {-1, -1, -1};
  {
    const int t0 = blockIdx.x * DM_WG + wg;
    if (t0 < g.total_tiles) {
      int r0, c0, ch0;
      tile_range<TM>(p.d, t0 / p.d.batch, r0, c0, ch0);
      load_row_idx<128>(p, r0, c0, wtid, nidx);
    }
  }

  for (int t = blockIdx.x * DM_WG + wg; t < g.total_tiles; t += stride) {
    const int b = t % p.d.batch, tile = t / p.d.batch;
    int row0, cnt, chunk;
    tile_range<TM>(p.d, tile, row0, cnt, chunk);
    const size_t grow0 = (size_t)b * p.d.rows + row0;

    // dOut = g0 rows (+ scale * gathered g1 rows): 4 units (row, 16-byte column group) per call
    auto dm_load = [&](int base, float4 (&va)[4], float4 (&vb)[4], float (&gs)[4]) {
      const float* g0p[4];
      const float* g1p[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int u = base + j * 128, row = u >> 4, col = (u & 15) * 4;
        g0p[j] = g1p[j] = nullptr;
        gs[j] = 1.f;
        if (row < cnt) {
          if (p.g0) {
            const size_t gr = p.g0_idx ? (size_t)b * p.d.rows + __ldg(p.g0_idx + row0 + row)
                                       : grow0 + row;
            g0p[j] = p.g0 + gr * FN + col;
          }
          if (p.g1) {
            const int gi = __ldg(p.g1_idx + row0 + row);
            if (p.g1_scale) gs[j] = __ldg(p.g1_scale + gi);
            g1p[j] = p.g1 + (size_t)b * p.g1_batch_stride + (size_t)gi * FN + col;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        va[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        vb[j] = va[j];
        if (g0p[j]) va[j] = __ldg(reinterpret_cast<const float4*>(g0p[j]));
        if (g1p[j]) vb[j] = __ldg(reinterpret_cast<const float4*>(g1p[j]));
      }
    };
    auto dm_store = [&](int base, const float4 (&va)[4], const float4 (&vb)[4],
                        const float (&gs)[4]) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int u = base + j * 128, row = u >> 4, c4 = u & 15;
        *reinterpret_cast<float4*>(stg + stg_idx(row, c4, FN)) =
            make_float4(va[j].x + gs[j] * vb[j].x, va[j].y + gs[j] * vb[j].y,
                        va[j].z + gs[j] * vb[j].z, va[j].w + gs[j] * vb[j].w);
      }
    };

    // ---------------- gather (pipelined) + GEMM 1: H = z . W1^T
    {
      const int cidx[NLAM_MAX_SRC] = {nidx[0], nidx[1], nidx[2]};
      gather_rows_pipe<FN, 128>(p, b, cidx, sA, wtid);
    }
    fence_async_smem();
    dm_wg_sync(wg);
    if (wtid == 0) {
      tc_fence_after();
      const uint32_t a0 = smem_u32(sA), w0 = smem_u32(sW1);
      for (int ks = 0; ks < k1steps; ++ks) {
        const uint32_t kb = ks >> 2, kin = (ks & 3) * 32;
        umma_bf16(tH, make_desc_k_sw128(a0 + kb * a_blk + kin),
                  make_desc_k_sw128(w0 + kb * (FN * 128u) + kin), idesc, ks > 0);
      }
      umma_commit(bar_m);
    }
    // while GEMM 1 runs: first batch of dOut loads, next tile's indices + L2 prefetch
    float4 d0a[4], d0b[4];
    float d0s[4];
    dm_load(wtid, d0a, d0b, d0s);
    {
      const int tn = t + stride;
      if (tn < g.total_tiles) {
        int r0n, cn, chn;
        tile_range<TM>(p.d, tn / p.d.batch, r0n, cn, chn);
        const int bn = tn % p.d.batch;
        load_row_idx<128>(p, r0n, cn, wtid, nidx);
        prefetch_rows_of(p, bn, nidx, true);
        if (wtid < cn) {
          if (p.g0) {
            const size_t gr = p.g0_idx ? (size_t)bn * p.d.rows + __ldg(p.g0_idx + r0n + wtid)
                                       : (size_t)bn * p.d.rows + r0n + wtid;
            prefetch_l2(p.g0 + gr * FN);
            prefetch_l2(p.g0 + gr * FN + 32);
          }
          if (p.g1) {
            const int gi = __ldg(p.g1_idx + r0n + wtid);
            const float* q = p.g1 + (size_t)bn * p.g1_batch_stride + (size_t)gi * FN;
            prefetch_l2(q);
            prefetch_l2(q + 32);
          }
        }
      }
    }
    mbar_wait(bar_m, ph_m);
    ph_m ^= 1;
    tc_fence_after();

    // ---------------- dOut rows -> swizzled fp32 staging (the z blocks are dead now)
    dm_store(wtid, d0a, d0b, d0s);
#pragma unroll 1
    for (int base = wtid + 512; base < TM * 16; base += 512) {
      float4 va[4], vb[4];
      float gs[4];
      dm_load(base, va, vb, gs);
      dm_store(base, va, vb, gs);
    }

    // ---------------- epilogue 1: a = SiLU(H + b1) -> bf16 tile
#pragma unroll
    for (int cc = 0; cc < FN; cc += 16) {
      float v[16];
      tmem_ld16(tH + lane_addr + (uint32_t)cc, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = silu_fast(v[j] + sB1[cc + j]);
#pragma unroll
      for (int h8 = 0; h8 < 2; ++h8) {
        uint4 pk = make_uint4(pack_bf16(v[h8 * 8 + 0], v[h8 * 8 + 1]),
                              pack_bf16(v[h8 * 8 + 2], v[h8 * 8 + 3]),
                              pack_bf16(v[h8 * 8 + 4], v[h8 * 8 + 5]),
                              pack_bf16(v[h8 * 8 + 6], v[h8 * 8 + 7]));
        *reinterpret_cast<uint4*>(sT + sw128_off(r, cc + h8 * 8, a_blk)) = pk;
      }
    }
    fence_async_smem();
    tc_fence_before();
    dm_wg_sync(wg);

    // ---------------- GEMM 2: Y = a . W2^T ; the a tile goes to HBM meanwhile
    if (wtid == 0) {
      tc_fence_after();
      const uint32_t a0 = smem_u32(sT), w0 = smem_u32(sW2);
      for (int ks = 0; ks < FN / 16; ++ks)
        umma_bf16(tY, make_desc_k_sw128(a0 + ks * 32), make_desc_k_sw128(w0 + ks * 32), idesc,
                  ks > 0);
      umma_commit(bar_m);
    }
    copy_tile_out_wg(sT, g.a_img + (size_t)t * a_blk, wtid);
    mbar_wait(bar_m, ph_m);
    ph_m ^= 1;
    tc_fence_after();

    // ---------------- epilogue 2: LayerNorm backward on whole rows -> dY tile + column sums
    {
      float y[FN];
#pragma unroll
      for (int cc = 0; cc < FN; cc += 16) {
        float v[16];
        tmem_ld16(tY + lane_addr + (uint32_t)cc, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) y[cc + j] = v[j] + sB2[cc + j];
      }
      float rstd = 1.f, m1 = 0.f, m2 = 0.f;
      if (has_ln) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < FN; ++j) s += y[j];
        const float mean = s * (1.0f / FN);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < FN; ++j) {
          y[j] -= mean;
          q += y[j] * y[j];
        }
        rstd = rsqrtf(q * (1.0f / FN) + LN_EPS);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          float dmv[16], pv[16];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 d4 = *reinterpret_cast<const float4*>(stg + stg_idx(r, ci * 4 + j4, FN));
            dmv[j4 * 4] = d4.x, dmv[j4 * 4 + 1] = d4.y, dmv[j4 * 4 + 2] = d4.z, dmv[j4 * 4 + 3] = d4.w;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float yh = y[ci * 16 + j] * rstd;
            const float dyh = dmv[j] * sG[ci * 16 + j];
            y[ci * 16 + j] = yh;  // keep y_hat
            pv[j] = dmv[j] * yh;
            s1 += dyh;
            s2 += dyh * yh;
          }
          acc_dg[ci] += warp_colsum16(pv, lane);
          acc_dbt[ci] += warp_colsum16(dmv, lane);
        }
        m1 = s1 * (1.0f / FN);
        m2 = s2 * (1.0f / FN);
      }
      dm_wg_sync(wg);  // every thread has copied the a tile out of sT
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        float v[16];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 d4 = *reinterpret_cast<const float4*>(stg + stg_idx(r, ci * 4 + j4, FN));
          v[j4 * 4] = d4.x, v[j4 * 4 + 1] = d4.y, v[j4 * 4 + 2] = d4.z, v[j4 * 4 + 3] = d4.w;
        }
        if (has_ln) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float dyh = v[j] * sG[ci * 16 + j];
            v[j] = r < cnt ? rstd * (dyh - m1 - y[ci * 16 + j] * m2) : 0.f;
          }
        }
        acc_db2[ci] += warp_colsum16(v, lane);
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {
          uint4 pk = make_uint4(pack_bf16(v[h8 * 8 + 0], v[h8 * 8 + 1]),
                                pack_bf16(v[h8 * 8 + 2], v[h8 * 8 + 3]),
                                pack_bf16(v[h8 * 8 + 4], v[h8 * 8 + 5]),
                                pack_bf16(v[h8 * 8 + 6], v[h8 * 8 + 7]));
          *reinterpret_cast<uint4*>(sT + sw128_off(r, ci * 16 + h8 * 8, a_blk)) = pk;
        }
      }
    }
    fence_async_smem();
    tc_fence_before();
    dm_wg_sync(wg);

    // ---------------- GEMM 3: dA = dY . W2 (into Y's columns); dY tile -> HBM
    if (wtid == 0) {
      tc_fence_after();
      const uint32_t a0 = smem_u32(sT), w0 = smem_u32(sW2);
      for (int ks = 0; ks < FN / 16; ++ks)
        umma_bf16(tY, make_desc_k_sw128(a0 + ks * 32),
                  make_desc_mn_sw128(w0 + (uint32_t)ks * 2048u, FN * 128u), idesc_mn, ks > 0);
      umma_commit(bar_m);
    }
    copy_tile_out_wg(sT, g.dy_img + (size_t)t * a_blk, wtid);
    mbar_wait(bar_m, ph_m);
    ph_m ^= 1;
    tc_fence_after();
    dm_wg_sync(wg);  // dY tile fully copied before dH overwrites it

    // ---------------- epilogue 3: dH = dA * SiLU'(H + b1) -> bf16 tile
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
      float v[16], h[16];
      tmem_ld16(tY + lane_addr + (uint32_t)(ci * 16), v);
      tmem_ld16(tH + lane_addr + (uint32_t)(ci * 16), h);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] *= silu_grad_fast(h[j] + sB1[ci * 16 + j]);
      acc_db1[ci] += warp_colsum16(v, lane);
#pragma unroll
      for (int h8 = 0; h8 < 2; ++h8) {
        uint4 pk = make_uint4(pack_bf16(v[h8 * 8 + 0], v[h8 * 8 + 1]),
                              pack_bf16(v[h8 * 8 + 2], v[h8 * 8 + 3]),
                              pack_bf16(v[h8 * 8 + 4], v[h8 * 8 + 5]),
                              pack_bf16(v[h8 * 8 + 6], v[h8 * 8 + 7]));
        *reinterpret_cast<uint4*>(sT + sw128_off(r, ci * 16 + h8 * 8, a_blk)) = pk;
      }
    }
    fence_async_smem();
    tc_fence_before();
    dm_wg_sync(wg);
    copy_tile_out_wg(sT, g.dh_img + (size_t)t * a_blk, wtid);

    // ---------------- GEMM 4 + epilogue 4: dZ = dH . W1, one source (64 columns) at a time,
    // double-buffered in the recycled H / Y columns
    if (g.need_dz) {
      auto issue_dz = [&](int kb) {
        tc_fence_after();
        const uint32_t a0 = smem_u32(sT);
        const uint32_t w0 = smem_u32(sW1) + (uint32_t)kb * (FN * 128u);
        for (int ks = 0; ks < FN / 16; ++ks)
          umma_bf16((kb & 1) ? tY : tH, make_desc_k_sw128(a0 + ks * 32),
                    make_desc_mn_sw128(w0 + (uint32_t)ks * 2048u, 0), idesc_mn, ks > 0);
        umma_commit(&bar_z[kb & 1]);
      };
      if (wtid == 0) issue_dz(0);
      for (int kb = 0; kb < n_src; ++kb) {
        if (wtid == 0 && kb + 1 < n_src) issue_dz(kb + 1);
        float* fdst = p.d_src[kb];
        const bool fres = fdst && (kb == p.d.residual_src) && p.g0;
        const bool reduce = fdst && kb == p.reduce_src;
        const int32_t* didx = fdst ? p.d_src_idx[kb] : nullptr;
        // units of the store phase: u = wtid + 128 i -> row = (wtid >> 4) + 8 i, c4 = wtid & 15
        int orow_i[16];
        if (didx && !reduce) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int row = (wtid >> 4) + 8 * i;
            orow_i[i] = row < cnt ? __ldg(didx + row0 + row) : 0;
          }
        }
        mbar_wait(&bar_z[kb & 1], (ph_z >> (kb & 1)) & 1u);
        ph_z ^= 1u << (kb & 1);
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < FN; cc += 16) {
          float v[16];
          tmem_ld16(((kb & 1) ? tY : tH) + lane_addr + (uint32_t)cc, v);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4)
            *reinterpret_cast<float4*>(stg + stg_idx(r, (cc >> 2) + j4, FN)) =
                make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
        }
        tc_fence_before();
        dm_wg_sync(wg);
        if (reduce) {
          const int seg_lo = __ldg(p.d.agg.tile_seg + tile);
          const int seg_hi = __ldg(p.d.agg.tile_seg + tile + 1);
          float* ro = fdst + (size_t)b * p.d.agg.n_seg * FN + (wtid & 15) * 4;
          for (int seg = seg_lo + (wtid >> 4); seg < seg_hi; seg += 8) {
            const int r0 = __ldg(p.d.agg.seg_ptr + seg) - row0;
            const int r1 = __ldg(p.d.agg.seg_ptr + seg + 1) - row0;
            float4* o4 = reinterpret_cast<float4*>(ro + (size_t)seg * FN);
            float4 old = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.reduce_accumulate) old = *o4;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int rr = r0; rr < r1; ++rr) {
              const float4 v = *reinterpret_cast<const float4*>(stg + stg_idx(rr, wtid & 15, FN));
              acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
            }
            acc.x += old.x, acc.y += old.y, acc.z += old.z, acc.w += old.w;
            *o4 = acc;
          }
        } else if (fdst) {
          float* o = fdst + (wtid & 15) * 4;
#pragma unroll 4
          for (int i = 0; i < 16; ++i) {
            const int row = (wtid >> 4) + 8 * i;
            if (row < cnt) {
              float4 v = *reinterpret_cast<const float4*>(stg + stg_idx(row, wtid & 15, FN));
              if (fres) {
                const size_t gr = p.g0_idx ? (size_t)b * p.d.rows + __ldg(p.g0_idx + row0 + row)
                                           : grow0 + row;
                const float4 e =
                    __ldg(reinterpret_cast<const float4*>(p.g0 + gr * FN) + (wtid & 15));
                v.x += e.x, v.y += e.y, v.z += e.z, v.w += e.w;
              }
              const size_t orow = didx ? (size_t)b * p.d.rows + orow_i[i] : grow0 + row;
              *reinterpret_cast<float4*>(o + orow * FN) = v;
            }
          }
        }
        dm_wg_sync(wg);
      }
    }
    tc_fence_before();
    dm_wg_sync(wg);  // sT / staging free for the next tile
  }

  // ---------------- per-CTA column sums -> vec_partial[blockIdx.x]
  tc_fence_before();
  __syncthreads();
  {
    float* sRed = reinterpret_cast<float*>(sm);  // [12 warps][4][64], region 0 of context 0
    if (lane < 16) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        sRed[(warp * 4 + 0) * 64 + i * 16 + lane] = acc_db1[i];
        sRed[(warp * 4 + 1) * 64 + i * 16 + lane] = acc_db2[i];
        sRed[(warp * 4 + 2) * 64 + i * 16 + lane] = acc_dg[i];
        sRed[(warp * 4 + 3) * 64 + i * 16 + lane] = acc_dbt[i];
      }
    }
    __syncthreads();
    float* dst = g.vec_partial + (size_t)blockIdx.x * g.vec_len;
    for (int e = tid; e < 4 * FN; e += DM_NT) {
      const int which = e >> 6, col = e & 63;
      if (which >= 2 && !has_ln) continue;
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < DM_NT / 32; ++w) s += sRed[(w * 4 + which) * 64 + col];
      dst[which * FN + col] = s;
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512u);
}

}  // namespace tc

bool tc_dgrad_mc_supported(const KParams& p, const tc::BGeo& g) {
  const nlam_rowmlp& d = p.d;
  if (tc::fast_n(p) != tc::DM_FN || !tc::fast_gather(p) || d.n_chunks != 1) return false;
  for (const float* w : {d.w.w1, d.w.w2})
    if (((uintptr_t)w) % 16 != 0) return false;
  return true;
}

int tc_dgrad_mc_grid(const tc::BGeo& g) {
  int grid = (g.total_tiles + tc::DM_WG - 1) / tc::DM_WG;
  return grid > 148 ? 148 : grid;
}

int tc_rowmlp_dgrad_mc(const KParams& p, const tc::BGeo& g, cudaStream_t st) {
  NLAM_CUDA(ensure_dyn_smem((const void*)tc::rowmlp_tc_dgrad_mc_kernel, (int)tc::DM_SMEM));
  tc::rowmlp_tc_dgrad_mc_kernel<<<tc_dgrad_mc_grid(g), tc::DM_NT, tc::DM_SMEM, st>>>(p, g);
  NLAM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace nlam
