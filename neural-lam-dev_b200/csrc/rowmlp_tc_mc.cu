// Multi-context bf16 tcgen05 row-MLP forward (d_hidden = d_out = source widths = 64).
//
// One persistent CTA per SM, 512 threads = 4 warpgroups.  The weights (W1, W2 as
// bf16 UMMA B operands, 32 KB) are staged ONCE per SM and shared; every warpgroup
// is an independent tile pipeline with its own 48 KB operand/staging region,
// its own 128 TMEM columns, its own mbarriers and a named barrier -- so four
// 128-row tiles are in flight per SM (the two-CTA kernel in rowmlp_tc.cu has two)
// and each phase's latency (gather loads, MMA round trips, TMEM loads) is hidden
// by the other three.  A thread owns a whole row in the epilogues: LayerNorm
// needs no cross-thread exchange and no extra barrier.
//
// Same math and outputs as rowmlp_tc_fwd_kernel (incl. residual, second output,
// row scatter and the fused receiver-segment sum).
#include "rowmlp_tc.cuh"

namespace nlam {
namespace tc {

constexpr int MC_WG = 4;
constexpr int MC_NT = 128 * MC_WG;
constexpr int MC_FN = 64;
constexpr uint32_t MC_REGION = 3u * TM * 128u;  // 48 KB: A operand (3 sources) | A2 | staging
constexpr uint32_t MC_OFF_W1 = MC_WG * MC_REGION;
constexpr uint32_t MC_OFF_W2 = MC_OFF_W1 + 3u * MC_FN * 128u;
constexpr uint32_t MC_OFF_PAR = MC_OFF_W2 + MC_FN * 128u;
constexpr uint32_t MC_OFF_BAR = MC_OFF_PAR + 4u * MC_FN * 4u;
constexpr uint32_t MC_SMEM = MC_OFF_BAR + 128u;
constexpr int MC_STG_LD = MC_FN + 4;

__device__ __forceinline__ void wg_sync(int wg) {
  asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory");
}

__global__ void __launch_bounds__(MC_NT, 1)
rowmlp_tc_fwd_mc_kernel(const __grid_constant__ KParams p, const __grid_constant__ Geo g) {
  extern __shared__ __align__(1024) uint8_t sm[];
  if (smem_u32(sm) & 1023u) __trap();
  constexpr int FN = MC_FN;
  const int tid = threadIdx.x, wg = tid >> 7, wtid = tid & 127;
  const int warp = tid >> 5;
  uint8_t* sR = sm + (uint32_t)wg * MC_REGION;  // this warpgroup's region
  uint8_t* sW1 = sm + MC_OFF_W1;
  uint8_t* sW2 = sm + MC_OFF_W2;
  float* sPar = reinterpret_cast<float*>(sm + MC_OFF_PAR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + MC_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MC_WG);
  float* stg = reinterpret_cast<float*>(sR);

  if (warp == 0) tmem_alloc(tmem_slot, 512u);
  if (tid == 32) {
    for (int i = 0; i < 2 * MC_WG; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  // weights: once per SM, shared by the four warpgroups
  {
    const int nch1 = (p.d.n_src * FN) >> 3;
    for (int u = tid; u < FN * nch1; u += MC_NT) {
      const int n = u / nch1, k0 = (u % nch1) * 8;
      const float4* q = reinterpret_cast<const float4*>(p.d.w.w1 + (size_t)n * p.k_total + k0);
      const float4 a = __ldg(q), c = __ldg(q + 1);
      *reinterpret_cast<uint4*>(sW1 + sw128_off(n, k0, FN * 128u)) =
          make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(c.x, c.y),
                     pack_bf16(c.z, c.w));
    }
    for (int u = tid; u < FN * (FN >> 3); u += MC_NT) {
      const int n = u / (FN >> 3), k0 = (u % (FN >> 3)) * 8;
      const float4* q = reinterpret_cast<const float4*>(p.d.w.w2 + (size_t)n * FN + k0);
      const float4 a = __ldg(q), c = __ldg(q + 1);
      *reinterpret_cast<uint4*>(sW2 + sw128_off(n, k0, FN * 128u)) =
          make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(c.x, c.y),
                     pack_bf16(c.z, c.w));
    }
    for (int i = tid; i < 4 * FN; i += MC_NT) {
      const int j = i % FN, which = i / FN;
      float v;
      if (which == 0) v = __ldg(p.d.w.b1 + j);
      else if (which == 1) v = __ldg(p.d.w.b2 + j);
      else if (which == 2) v = p.d.w.ln_g ? __ldg(p.d.w.ln_g + j) : 1.f;
      else v = p.d.w.ln_g ? __ldg(p.d.w.ln_b + j) : 0.f;
      sPar[i] = v;
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tH = tmem_base + (uint32_t)wg * 128u, tY = tH + 64u;
  uint64_t* bar0 = &bars[2 * wg];
  uint64_t* bar1 = &bars[2 * wg + 1];
  uint32_t ph0 = 0, ph1 = 0;

  const uint32_t idesc = make_idesc_bf16(TM, FN);
  const uint32_t a_blk = TM * 128u;
  const int r = wtid;  // TMEM lane == tile row (warp % 4 selects the 32-lane quarter)
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
  const float* sB1 = sPar;
  const float* sB2 = sPar + FN;
  const float* sG = sPar + 2 * FN;
  const float* sBe = sPar + 3 * FN;
  const bool has_ln = p.d.w.ln_g != nullptr;
  const int k1steps = p.d.n_src * FN / 16;
  const int stride = gridDim.x * MC_WG;

  // source-row indices of this thread's row, loaded one tile ahead
  int nidx[NLAM_MAX_SRC] = {-1, -1, -1};
  {
    const int t0 = blockIdx.x + gridDim.x * wg;
    if (t0 < g.total_tiles) {
      int r0, c0, ch0;
      tile_range<TM>(p.d, t0 / p.d.batch, r0, c0, ch0);
      load_row_idx<128>(p, r0, c0, wtid, nidx);
    }
  }
  // Programmatic dependent launch: the prologue above (TMEM, weights, the first tile's
  // static row indices) overlapped the previous kernel's tail; its outputs are read from
  // here on.  All CTAs of this persistent grid are resident, so the next kernel's CTAs
  // may take over each SM as soon as this kernel's CTA there exits.
  pdl_wait();
  pdl_trigger();
  // SM-major round robin: tile counts differ by at most one across SMs
  for (int t = blockIdx.x + gridDim.x * wg; t < g.total_tiles; t += stride) {
    const int b = t % p.d.batch, tile = t / p.d.batch;  // batch innermost: shared rows hit L2
    int row0, cnt, chunk;
    tile_range<TM>(p.d, tile, row0, cnt, chunk);

    // ---------------- gather (bf16 A operand); indices of the next tile + L2 prefetch
    int cidx[NLAM_MAX_SRC] = {nidx[0], nidx[1], nidx[2]};
    gather_rows_pipe<FN, 128>(p, b, cidx, sR, wtid);
    {
      const int tn = t + stride;
      if (tn < g.total_tiles) {
        int r0n, cn, chn;
        tile_range<TM>(p.d, tn / p.d.batch, r0n, cn, chn);
        const int bn = tn % p.d.batch;
        load_row_idx<128>(p, r0n, cn, wtid, nidx);
#pragma unroll
        for (int s = 0; s < NLAM_MAX_SRC; ++s) {
          const int ri = s == 0 ? nidx[0] : s == 1 ? nidx[1] : nidx[2];
          if (s < p.d.n_src && ri >= 0) {
            const nlam_src& src = p.d.src[s];
            const char* q = reinterpret_cast<const char*>(
                src.ptr + (long long)bn * src.batch_stride + (long long)ri * src.ld);
            prefetch_l2(q);
            prefetch_l2(q + 128);
          }
        }
      }
    }
    fence_async_smem();
    wg_sync(wg);

    // ---------------- GEMM 1: H = A . W1^T
    if (wtid == 0) {
      tc_fence_after();
      const uint32_t a0 = smem_u32(sR), w0 = smem_u32(sW1);
      for (int ks = 0; ks < k1steps; ++ks) {
        const uint32_t kb = ks >> 2, kin = (ks & 3) * 32;
        umma_bf16(tH, make_desc_k_sw128(a0 + kb * a_blk + kin),
                  make_desc_k_sw128(w0 + kb * (FN * 128u) + kin), idesc, ks > 0);
      }
      umma_commit(bar0);
    }
    mbar_wait(bar0, ph0);
    ph0 ^= 1;
    tc_fence_after();

    // ---------------- epilogue 1: a = SiLU(H + b1) -> bf16 A2 (block 0 of the region)
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
        *reinterpret_cast<uint4*>(sR + sw128_off(r, cc + h8 * 8, a_blk)) = pk;
      }
    }
    fence_async_smem();
    tc_fence_before();
    wg_sync(wg);

    // ---------------- GEMM 2: Y = A2 . W2^T
    if (wtid == 0) {
      tc_fence_after();
      const uint32_t a0 = smem_u32(sR), w0 = smem_u32(sW2);
      for (int ks = 0; ks < FN / 16; ++ks)
        umma_bf16(tY, make_desc_k_sw128(a0 + ks * 32), make_desc_k_sw128(w0 + ks * 32), idesc,
                  ks > 0);
      umma_commit(bar1);
    }
    mbar_wait(bar1, ph1);
    ph1 ^= 1;
    tc_fence_after();

    // ---------------- epilogue 2: y + b2 -> LayerNorm (whole row in registers) -> staging
    {
      float y[FN];
#pragma unroll
      for (int cc = 0; cc < FN; cc += 16) {
        float v[16];
        tmem_ld16(tY + lane_addr + (uint32_t)cc, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) y[cc + j] = v[j] + sB2[cc + j];
      }
      if (has_ln) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < FN; ++j) s += y[j];
        const float mean = s * (1.0f / FN);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < FN; ++j) {
          const float dl = y[j] - mean;
          q += dl * dl;
        }
        const float rstd = rsqrtf(q * (1.0f / FN) + LN_EPS);
#pragma unroll
        for (int j = 0; j < FN; ++j) y[j] = (y[j] - mean) * rstd * sG[j] + sBe[j];
      }
      // GEMM 2 has completed (mbarrier): the region may now hold the staging tile
      float4* dst = reinterpret_cast<float4*>(stg + (size_t)r * MC_STG_LD);
#pragma unroll
      for (int j4 = 0; j4 < FN / 4; ++j4)
        dst[j4] = make_float4(y[j4 * 4], y[j4 * 4 + 1], y[j4 * 4 + 2], y[j4 * 4 + 3]);
    }
    tc_fence_before();
    wg_sync(wg);

    // ---------------- coalesced store / scatter / fused segment reduction
    if (p.d.out || p.d.out_res) {
      float* out = p.d.out ? p.d.out + (size_t)b * p.d.rows * FN : nullptr;
      float* out2 = p.d.out_res ? p.d.out_res + (size_t)b * p.d.rows * FN : nullptr;
      const int32_t* oidx = p.d.out_idx;
      const nlam_src& s0 = p.d.src[0];
      const bool res = p.d.residual_src == 0;
      const bool need0 = res || out2;
      for (int u = wtid; u < cnt * 16; u += 128) {
        const int row = u >> 4, c4 = u & 15;
        float4 v = *reinterpret_cast<const float4*>(stg + (size_t)row * MC_STG_LD + c4 * 4);
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
        if (need0) {
          const int ridx = s0.idx ? __ldg(s0.idx + row0 + row) : row0 + row;
          e = __ldg(reinterpret_cast<const float4*>(s0.ptr + (long long)b * s0.batch_stride +
                                                    (long long)ridx * s0.ld) + c4);
        }
        if (res) v.x += e.x, v.y += e.y, v.z += e.z, v.w += e.w;
        const size_t orow = oidx ? (size_t)__ldg(oidx + row0 + row) : (size_t)(row0 + row);
        if (out) *reinterpret_cast<float4*>(out + orow * FN + c4 * 4) = v;
        if (out2)
          *reinterpret_cast<float4*>(out2 + orow * FN + c4 * 4) =
              make_float4(v.x + e.x, v.y + e.y, v.z + e.z, v.w + e.w);
      }
    }
    if (p.d.agg.out) {
      const int seg_lo = __ldg(p.d.agg.tile_seg + tile), seg_hi = __ldg(p.d.agg.tile_seg + tile + 1);
      float* ao = p.d.agg.out + (size_t)b * p.d.agg.n_seg * FN;
      for (int u = wtid; u < (seg_hi - seg_lo) * 16; u += 128) {
        const int seg = seg_lo + (u >> 4), c4 = u & 15;
        const int r0 = __ldg(p.d.agg.seg_ptr + seg) - row0, r1 = __ldg(p.d.agg.seg_ptr + seg + 1) - row0;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int rr = r0; rr < r1; ++rr) {
          const float4 v = *reinterpret_cast<const float4*>(stg + (size_t)rr * MC_STG_LD + c4 * 4);
          acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
        }
        if (p.d.agg.scale) {
          const float sc = __ldg(p.d.agg.scale + seg);
          acc.x *= sc, acc.y *= sc, acc.z *= sc, acc.w *= sc;
        }
        *reinterpret_cast<float4*>(ao + (size_t)seg * FN + c4 * 4) = acc;
      }
    }
    wg_sync(wg);  // staging is free for the next gather
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512u);
}

}  // namespace tc

// Eligibility: square 64-wide fast path, single weight set, vectorisable I/O.
bool tc_fwd_mc_supported(const KParams& p) {
  const nlam_rowmlp& d = p.d;
  if (tc::fast_n(p) != tc::MC_FN || !tc::fast_gather(p) || d.n_chunks != 1) return false;
  if (!p.out_vec_ok) return false;
  if ((d.residual_src == 0 || d.out_res) && !p.vec_ok[0]) return false;
  for (const float* w : {d.w.w1, d.w.w2})
    if (((uintptr_t)w) % 16 != 0) return false;
  return true;
}

int tc_rowmlp_fwd_mc(const KParams& p, const tc::Geo& g, cudaStream_t st) {
  NLAM_CUDA(ensure_dyn_smem((const void*)tc::rowmlp_tc_fwd_mc_kernel, (int)tc::MC_SMEM));
  int grid = (g.total_tiles + tc::MC_WG - 1) / tc::MC_WG;
  if (grid > 148) grid = 148;
  NLAM_CUDA(launch_k(tc::rowmlp_tc_fwd_mc_kernel, grid, tc::MC_NT, tc::MC_SMEM, st, p, g));
  NLAM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace nlam
