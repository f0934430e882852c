// TMA row-gather variant of the multi-context bf16 tcgen05 row-MLP forward
// (d_hidden = d_out = source widths = 64, every source has a bf16 shadow).
//
// Same pipeline as rowmlp_tc_mc.cu -- one persistent CTA per SM, four warpgroups, each an
// independent tile pipeline over its own 48 KB region / 128 TMEM columns, ONE shared copy of
// the weights -- but the A operand of GEMM 1 no longer passes through registers: warp 0 of
// each warpgroup issues 3 x 32 cp.async.bulk.tensor tile::gather4 copies (4 gathered bf16
// rows of 128 bytes each; SWIZZLE_128B tensor map = the UMMA K-major SW128 layout) as soon
// as the region is free, i.e. right after the previous tile's stores; completion is an
// mbarrier transaction count.  The compute warps never execute a gather load, a conversion
// or an operand store.
//
// Reference: interaction_net.py:103-121 (propagate: x_j / x_i gather + message MLP) and
// :124-131 (aggregate).
#include "rowmlp_tc.cuh"
#include "tma.cuh"

namespace nlam {
namespace tc {

constexpr int TF_WG = 4;
constexpr int TF_NT = 128 * TF_WG;
constexpr int TF_FN = 64;
constexpr uint32_t TF_BLK = TM * 128u;             // one 64-column bf16 block: 16 KB
constexpr uint32_t TF_REGION = 3u * TF_BLK;        // z (3 sources) | A2 | staging
constexpr uint32_t TF_OFF_W1 = TF_WG * TF_REGION;
constexpr uint32_t TF_OFF_W2 = TF_OFF_W1 + 3u * TF_FN * 128u;
constexpr uint32_t TF_OFF_PAR = TF_OFF_W2 + TF_FN * 128u;
constexpr uint32_t TF_OFF_BAR = TF_OFF_PAR + 4u * TF_FN * 4u;
constexpr uint32_t TF_SMEM = TF_OFF_BAR + 128u;
constexpr int TF_STG_LD = TF_FN + 4;
constexpr int TF_NBAR = 3;  // per warpgroup: GEMM 1, GEMM 2, operand tile landed

struct alignas(64) TmaFwdMaps {
  CUtensorMap m[NLAM_MAX_SRC];
  int batch_rows[NLAM_MAX_SRC];  // rows between batch items of the shadow (0 = shared)
};

__device__ __forceinline__ void tf_sync(int wg) {
  asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory");
}

__global__ void __launch_bounds__(TF_NT, 1)
rowmlp_tc_fwd_tma_kernel(const __grid_constant__ KParams p, const __grid_constant__ Geo g,
                         const __grid_constant__ TmaFwdMaps tm) {
  extern __shared__ __align__(1024) uint8_t sm[];
  if (smem_u32(sm) & 1023u) __trap();
  constexpr int FN = TF_FN;
  const int tid = threadIdx.x, wg = tid >> 7, wtid = tid & 127;
  const int warp = tid >> 5;
  uint8_t* sR = sm + (uint32_t)wg * TF_REGION;
  uint8_t* sW1 = sm + TF_OFF_W1;
  uint8_t* sW2 = sm + TF_OFF_W2;
  float* sPar = reinterpret_cast<float*>(sm + TF_OFF_PAR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + TF_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + TF_NBAR * TF_WG);
  float* stg = reinterpret_cast<float*>(sR);

  if (warp == 0) tmem_alloc(tmem_slot, 512u);
  if (tid == 32) {
    for (int i = 0; i < TF_NBAR * TF_WG; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  if (tid == 64)
    for (int s = 0; s < 3; ++s) tma_prefetch_desc(&tm.m[s]);
  {  // weights: once per SM, shared by the four warpgroups
    for (int u = tid; u < FN * 24; u += TF_NT) {
      const int n = u / 24, k0 = (u % 24) * 8;
      const float4* q = reinterpret_cast<const float4*>(p.d.w.w1 + (size_t)n * 192 + k0);
      const float4 a = __ldg(q), c = __ldg(q + 1);
      *reinterpret_cast<uint4*>(sW1 + sw128_off(n, k0, FN * 128u)) =
          make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(c.x, c.y),
                     pack_bf16(c.z, c.w));
    }
    for (int u = tid; u < FN * 8; u += TF_NT) {
      const int n = u >> 3, k0 = (u & 7) * 8;
      const float4* q = reinterpret_cast<const float4*>(p.d.w.w2 + (size_t)n * FN + k0);
      const float4 a = __ldg(q), c = __ldg(q + 1);
      *reinterpret_cast<uint4*>(sW2 + sw128_off(n, k0, FN * 128u)) =
          make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(c.x, c.y),
                     pack_bf16(c.z, c.w));
    }
    for (int i = tid; i < 4 * FN; i += TF_NT) {
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
  uint64_t* bar0 = &bars[TF_NBAR * wg];
  uint64_t* bar1 = bar0 + 1;
  uint64_t* barz = bar0 + 2;
  uint32_t ph0 = 0, ph1 = 0, phz = 0;

  const uint32_t idesc = make_idesc_bf16(TM, FN);
  const int r = wtid;  // TMEM lane == tile row
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
  const float* sB1 = sPar;
  const float* sB2 = sPar + FN;
  const float* sG = sPar + 2 * FN;
  const float* sBe = sPar + 3 * FN;
  const bool has_ln = p.d.w.ln_g != nullptr;
  const int stride = gridDim.x * TF_WG;

  // warp 0 of the warpgroup: gather the three 128-row operand blocks of tile t.  Lane l owns
  // tile rows 4l .. 4l+3 (one gather4 per source); rows past the end of a short tile repeat
  // the tile's first row (their results are never stored).
  auto issue_gather = [&](int t) {
    const int b = t % p.d.batch;
    int row0, cnt, chunk;
    tile_range<TM>(p.d, t / p.d.batch, row0, cnt, chunk);
    const int lane = wtid;
    fence_async_smem();  // the region's last generic-proxy accesses precede the async writes
    if (lane == 0) mbar_arrive_expect_tx(barz, 3u * TF_BLK);
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int32_t* idx = p.d.src[s].idx;
      const int boff = b * tm.batch_rows[s];
      int rr[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int row = 4 * lane + j;
        const int gr = row0 + (row < cnt ? row : 0);
        rr[j] = (idx ? __ldg(idx + gr) : gr) + boff;
      }
      tma_gather4(sR + (uint32_t)s * TF_BLK + (uint32_t)lane * 512u, &tm.m[s], barz, 0, rr[0], rr[1],
                  rr[2], rr[3]);
    }
  };

  // Programmatic dependent launch: the prologue overlapped the previous kernel's tail; its
  // outputs (the shadows gathered below) are read from here on.
  pdl_wait();
  pdl_trigger();
  if (wtid < 32) {
    const int t0 = blockIdx.x + gridDim.x * wg;
    if (t0 < g.total_tiles) issue_gather(t0);
  }
  for (int t = blockIdx.x + gridDim.x * wg; t < g.total_tiles; t += stride) {
    const int b = t % p.d.batch, tile = t / p.d.batch;
    int row0, cnt, chunk;
    tile_range<TM>(p.d, tile, row0, cnt, chunk);

    // ---------------- GEMM 1: H = z . W1^T as soon as the operand tile has landed
    if (wtid == 0) {
      mbar_wait(barz, phz);
      tc_fence_after();
      const uint32_t a0 = smem_u32(sR), w0 = smem_u32(sW1);
      for (int ks = 0; ks < 12; ++ks) {
        const uint32_t kb = ks >> 2, kin = (ks & 3) * 32;
        umma_bf16(tH, make_desc_k_sw128(a0 + kb * TF_BLK + kin),
                  make_desc_k_sw128(w0 + kb * (FN * 128u) + kin), idesc, ks > 0);
      }
      umma_commit(bar0);
    }
    phz ^= 1;
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
        *reinterpret_cast<uint4*>(sR + sw128_off(r, cc + h8 * 8, TF_BLK)) = pk;
      }
    }
    fence_async_smem();
    tc_fence_before();
    tf_sync(wg);

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
      float4* dst = reinterpret_cast<float4*>(stg + (size_t)r * TF_STG_LD);
#pragma unroll
      for (int j4 = 0; j4 < FN / 4; ++j4)
        dst[j4] = make_float4(y[j4 * 4], y[j4 * 4 + 1], y[j4 * 4 + 2], y[j4 * 4 + 3]);
    }
    tc_fence_before();
    tf_sync(wg);

    // ---------------- coalesced store / scatter / fused segment reduction (+ bf16 shadows)
    if (p.d.out || p.d.out_res) {
      float* out = p.d.out ? p.d.out + (size_t)b * p.d.rows * FN : nullptr;
      float* out2 = p.d.out_res ? p.d.out_res + (size_t)b * p.d.rows * FN : nullptr;
      const int32_t* oidx = p.d.out_idx;
      const nlam_src& s0 = p.d.src[0];
      const bool res = p.d.residual_src == 0;
      const bool need0 = res || out2;
      for (int u = wtid; u < cnt * 16; u += 128) {
        const int row = u >> 4, c4 = u & 15;
        float4 v = *reinterpret_cast<const float4*>(stg + (size_t)row * TF_STG_LD + c4 * 4);
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
        if (need0) {
          const int ridx = s0.idx ? __ldg(s0.idx + row0 + row) : row0 + row;
          e = __ldg(reinterpret_cast<const float4*>(s0.ptr + (long long)b * s0.batch_stride +
                                                    (long long)ridx * s0.ld) + c4);
        }
        if (res) v.x += e.x, v.y += e.y, v.z += e.z, v.w += e.w;
        const size_t orow = oidx ? (size_t)__ldg(oidx + row0 + row) : (size_t)(row0 + row);
        if (out) *reinterpret_cast<float4*>(out + orow * FN + c4 * 4) = v;
        if (p.d.out_bf16)
          *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.d.out_bf16) +
                                    ((size_t)b * p.d.rows + orow) * FN + c4 * 4) =
              make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
        if (out2)
          *reinterpret_cast<float4*>(out2 + orow * FN + c4 * 4) =
              make_float4(v.x + e.x, v.y + e.y, v.z + e.z, v.w + e.w);
        if (p.d.out_res_bf16)
          *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.d.out_res_bf16) +
                                    ((size_t)b * p.d.rows + orow) * FN + c4 * 4) =
              make_uint2(pack_bf16(v.x + e.x, v.y + e.y), pack_bf16(v.z + e.z, v.w + e.w));
      }
    }
    if (p.d.agg.out || p.d.agg.out_bf16) {
      const int seg_lo = __ldg(p.d.agg.tile_seg + tile), seg_hi = __ldg(p.d.agg.tile_seg + tile + 1);
      for (int u = wtid; u < (seg_hi - seg_lo) * 16; u += 128) {
        const int seg = seg_lo + (u >> 4), c4 = u & 15;
        const int r0 = __ldg(p.d.agg.seg_ptr + seg) - row0, r1 = __ldg(p.d.agg.seg_ptr + seg + 1) - row0;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int rr = r0; rr < r1; ++rr) {
          const float4 v = *reinterpret_cast<const float4*>(stg + (size_t)rr * TF_STG_LD + c4 * 4);
          acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
        }
        if (p.d.agg.scale) {
          const float sc = __ldg(p.d.agg.scale + seg);
          acc.x *= sc, acc.y *= sc, acc.z *= sc, acc.w *= sc;
        }
        const size_t o = ((size_t)b * p.d.agg.n_seg + seg) * FN + c4 * 4;
        if (p.d.agg.out) *reinterpret_cast<float4*>(p.d.agg.out + o) = acc;
        if (p.d.agg.out_bf16)
          *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.d.agg.out_bf16) + o) =
              make_uint2(pack_bf16(acc.x, acc.y), pack_bf16(acc.z, acc.w));
      }
    }
    tf_sync(wg);  // staging is free: refill the region for the next tile right away
    if (wtid < 32 && t + stride < g.total_tiles) issue_gather(t + stride);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512u);
}

}  // namespace tc

// Eligibility: the multi-context forward's conditions + three 64-wide sources that all
// carry a dense bf16 shadow.
bool tc_fwd_tma_supported(const KParams& p) {
  const nlam_rowmlp& d = p.d;
  if (!tc_fwd_mc_supported(p) || d.n_src != 3) return false;
  for (int s = 0; s < 3; ++s) {
    const nlam_src& src = d.src[s];
    if (!src.shadow || ((uintptr_t)src.shadow) % 16 != 0 || src.shadow_rows <= 0) return false;
    if (src.shadow_batch_stride != 0 && src.shadow_batch_stride != src.shadow_rows * 64) return false;
    const long long total = src.shadow_batch_stride ? (long long)d.batch * src.shadow_rows : src.shadow_rows;
    if (total >= (1ll << 31)) return false;
  }
  return true;
}

int tc_rowmlp_fwd_tma(const KParams& p, const tc::Geo& g, cudaStream_t st) {
  tc::TmaFwdMaps tm;
  for (int s = 0; s < 3; ++s) {
    const nlam_src& src = p.d.src[s];
    const long long total = src.shadow_batch_stride ? (long long)p.d.batch * src.shadow_rows : src.shadow_rows;
    if (make_row_map_bf16(&tm.m[s], src.shadow, total, 1)) return 1;
    tm.batch_rows[s] = src.shadow_batch_stride ? (int)src.shadow_rows : 0;
  }
  NLAM_CUDA(ensure_dyn_smem((const void*)tc::rowmlp_tc_fwd_tma_kernel, (int)tc::TF_SMEM));
  int grid = (g.total_tiles + tc::TF_WG - 1) / tc::TF_WG;
  if (grid > 148) grid = 148;
  NLAM_CUDA(launch_k(tc::rowmlp_tc_fwd_tma_kernel, grid, tc::TF_NT, tc::TF_SMEM, st, p, g, tm));
  NLAM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace nlam
