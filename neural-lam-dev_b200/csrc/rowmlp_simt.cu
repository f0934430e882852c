// fp32 (FFMA) row-MLP kernels: the parity-mode data path and the path for
// widths the tensor-core kernels do not cover.
//
//   out[b,r,:] = [src_res +] LN( W2 . SiLU( W1 . concat_s src_s[b, idx_s[r], :] + b1 ) + b2 )
//
// One CTA (256 threads) owns a tile of 64 rows.  The gathered, concatenated
// input rows are staged once in shared memory (128-bit row loads), the two
// Linear layers run as register-blocked FFMA GEMMs (4 rows x DP/16 columns per
// thread) with the weights streamed through shared memory in 32-deep chunks,
// SiLU / bias / LayerNorm / residual are applied in registers, and only the
// final rows go back to HBM.  Backward recomputes the forward per tile, then
// produces per-row input gradients (dgrad) and saves SiLU output, dY and dH for
// a split-row weight-gradient kernel whose partials are summed in fixed order
// (deterministic, no float atomics).
//
// Reference semantics: utils.make_mlp (utils.py:191-214), InteractionNet
// .message / aggr_mlp (interaction_net.py:106,117-121), SplitMLPs (:134-163).
#include <math.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "rowmlp_common.cuh"

namespace nlam {

constexpr int TM = NLAM_TILE_ROWS;  // rows per tile
constexpr int NT = 256;             // threads per CTA


template <int DP>
struct Cfg {
  static constexpr int CPT = DP / 16;              // output columns per thread
  static constexpr int VEC = CPT >= 4 ? 4 : CPT;   // vector width of W reads
  static constexpr int NV = CPT / VEC;
  static constexpr int KC = DP >= 32 ? 32 : 16;    // reduction chunk
  static constexpr int HS = DP + 4;                // stride of [TM][DP] tiles
  static constexpr int AS = 3 * HS;                // stride of the input tile
  static constexpr int WS = DP + 4;                // stride of the weight chunk
  static constexpr int WSZ = (KC * WS > 32 * DP) ? KC * WS : 32 * DP;
  static constexpr size_t SMEM = sizeof(float) * (size_t)(TM * AS + TM * HS + WSZ);
};

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float silu_grad_f(float x) {
  float s = 1.0f / (1.0f + expf(-x));
  return s * (1.0f + x * (1.0f - s));
}

template <int DP>
__device__ __forceinline__ int col_of(int tx, int c) {
  using C = Cfg<DP>;
  return (tx + 16 * (c / C::VEC)) * C::VEC + (c % C::VEC);
}

// Stage the gathered + concatenated input rows of a tile in shared memory.
template <int DP>
__device__ void gather_tile(const KParams& p, int b, int row0, int cnt, float* As) {
  using C = Cfg<DP>;
  const int tid = threadIdx.x;
  const int kpad = ((p.k_total + C::KC - 1) / C::KC) * C::KC;
  // zero: padding columns of valid rows, everything of invalid rows
  const int padw = kpad - p.k_total;
  for (int e = tid; e < cnt * padw; e += NT) {
    int r = e / padw, c = e % padw;
    As[r * C::AS + p.k_total + c] = 0.f;
  }
  for (int e = tid; e < (TM - cnt) * kpad; e += NT) {
    int r = cnt + e / kpad, c = e % kpad;
    As[r * C::AS + c] = 0.f;
  }
  for (int s = 0; s < p.d.n_src; ++s) {
    const nlam_src& src = p.d.src[s];
    const float* base = src.ptr + (long long)b * src.batch_stride;
    const int w = src.width, ko = p.koff[s];
    if (p.vec_ok[s]) {
      const int w4 = w >> 2;
      for (int e = tid; e < cnt * w4; e += NT) {
        int r = e / w4, c4 = e % w4;
        int ridx = src.idx ? src.idx[row0 + r] : row0 + r;
        float4 v = __ldg(reinterpret_cast<const float4*>(base + (long long)ridx * src.ld) + c4);
        *reinterpret_cast<float4*>(As + r * C::AS + ko + c4 * 4) = v;
      }
    } else {
      for (int e = tid; e < cnt * w; e += NT) {
        int r = e / w, c = e % w;
        int ridx = src.idx ? src.idx[row0 + r] : row0 + r;
        As[r * C::AS + ko + c] = __ldg(base + (long long)ridx * src.ld + c);
      }
    }
  }
}

// acc[4][CPT] = A[rows ty*4..+3][0:kred] x Wm, streaming Wm through shared memory.
//   TRANS  : Wm(i, j) = W[j*ldw + i]        (forward: W is [N][K], reduce over K)
//   !TRANS : Wm(i, j) = W[i*ldw + joff + j] (backward: reduce over W's rows)
// i < kred, j < nvalid are real, the rest is zero-filled.
template <int DP, bool TRANS>
__device__ __forceinline__ void gemm_tile(float (&acc)[4][Cfg<DP>::CPT], const float* A, int lda,
                                          int kred, const float* __restrict__ W, int ldw,
                                          int nvalid, int joff, float* Ws) {
  using C = Cfg<DP>;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < C::CPT; ++c) acc[i][c] = 0.f;
  const float* Arow = A + (ty * 4) * lda;
  for (int k0 = 0; k0 < kred; k0 += C::KC) {
    __syncthreads();  // previous chunk consumed / A tile written
    for (int e = tid; e < C::KC * DP; e += NT) {
      int ii, j;
      float v = 0.f;
      if (TRANS) {
        ii = e % C::KC;
        j = e / C::KC;
        if (k0 + ii < kred && j < nvalid) v = __ldg(W + (long long)j * ldw + k0 + ii);
      } else {
        j = e % DP;
        ii = e / DP;
        if (k0 + ii < kred && j < nvalid) v = __ldg(W + (long long)(k0 + ii) * ldw + joff + j);
      }
      Ws[ii * C::WS + j] = v;
    }
    __syncthreads();
#pragma unroll 2
    for (int kk = 0; kk < C::KC; kk += 4) {
      float4 a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        a[i] = *reinterpret_cast<const float4*>(Arow + i * lda + k0 + kk);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float w[C::CPT];
        const float* wrow = Ws + (kk + q) * C::WS;
#pragma unroll
        for (int v = 0; v < C::NV; ++v) {
          const float* wp = wrow + (tx + 16 * v) * C::VEC;
          if (C::VEC == 4) {
            float4 t = *reinterpret_cast<const float4*>(wp);
            w[v * 4 + 0] = t.x, w[v * 4 + 1] = t.y, w[v * 4 + 2] = t.z, w[v * 4 + 3] = t.w;
          } else if (C::VEC == 2) {
            float2 t = *reinterpret_cast<const float2*>(wp);
            w[v * 2 + 0] = t.x, w[v * 2 + 1] = t.y;
          } else {
            w[v] = *wp;
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float av = q == 0 ? a[i].x : q == 1 ? a[i].y : q == 2 ? a[i].z : a[i].w;
#pragma unroll
          for (int c = 0; c < C::CPT; ++c) acc[i][c] = fmaf(av, w[c], acc[i][c]);
        }
      }
    }
  }
}

__device__ __forceinline__ float group16_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

// LayerNorm statistics of 4 rows held across the 16 lanes of a row group.
template <int DP>
__device__ __forceinline__ void ln_stats(const float (&y)[4][Cfg<DP>::CPT], int tx, int dout,
                                         float (&mean)[4], float (&rstd)[4]) {
  using C = Cfg<DP>;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C::CPT; ++c)
      if (col_of<DP>(tx, c) < dout) s += y[i][c];
    s = group16_sum(s);
    mean[i] = s / (float)dout;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < C::CPT; ++c)
      if (col_of<DP>(tx, c) < dout) {
        float dlt = y[i][c] - mean[i];
        q += dlt * dlt;
      }
    q = group16_sum(q);
    rstd[i] = 1.0f / sqrtf(q / (float)dout + LN_EPS);
  }
}

// -------------------------------------------------------------------- forward
template <int DP>
__global__ void __launch_bounds__(NT) rowmlp_fwd_kernel(const __grid_constant__ KParams p) {
  using C = Cfg<DP>;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Hs = As + TM * C::AS;
  float* Ws = Hs + TM * C::HS;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int b = blockIdx.y;
  int row0, cnt, chunk;
  tile_range<TM>(p.d, blockIdx.x, row0, cnt, chunk);
  if (cnt <= 0) return;
  const int dh = p.d.d_hidden, dout = p.d.d_out;
  const float* w1 = p.d.w.w1 + (size_t)chunk * dh * p.k_total;
  const float* b1 = p.d.w.b1 + (size_t)chunk * dh;
  const float* w2 = p.d.w.w2 + (size_t)chunk * dout * dh;
  const float* b2 = p.d.w.b2 + (size_t)chunk * dout;

  gather_tile<DP>(p, b, row0, cnt, As);

  float acc[4][C::CPT];
  gemm_tile<DP, true>(acc, As, C::AS, p.k_total, w1, p.k_total, dh, 0, Ws);
#pragma unroll
  for (int c = 0; c < C::CPT; ++c) {
    const int col = col_of<DP>(tx, c);
    const float bias = col < dh ? __ldg(b1 + col) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      Hs[(ty * 4 + i) * C::HS + col] = col < dh ? silu_f(acc[i][c] + bias) : 0.f;
  }
  gemm_tile<DP, true>(acc, Hs, C::HS, dh, w2, dh, dout, 0, Ws);
#pragma unroll
  for (int c = 0; c < C::CPT; ++c) {
    const int col = col_of<DP>(tx, c);
    const float bias = col < dout ? __ldg(b2 + col) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][c] += bias;
  }
  if (p.d.w.ln_g) {
    const float* g = p.d.w.ln_g + (size_t)chunk * dout;
    const float* be = p.d.w.ln_b + (size_t)chunk * dout;
    float mean[4], rstd[4];
    ln_stats<DP>(acc, tx, dout, mean, rstd);
#pragma unroll
    for (int c = 0; c < C::CPT; ++c) {
      const int col = col_of<DP>(tx, c);
      if (col < dout) {
        const float gg = __ldg(g + col), bb = __ldg(be + col);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][c] = (acc[i][c] - mean[i]) * rstd[i] * gg + bb;
      }
    }
  }
  if (p.d.residual_src >= 0) {
    const int ko = p.koff[p.d.residual_src];
#pragma unroll
    for (int c = 0; c < C::CPT; ++c) {
      const int col = col_of<DP>(tx, c);
      if (col < dout) {
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][c] += As[(ty * 4 + i) * C::AS + ko + col];
      }
    }
  }
  float* out = p.d.out + ((size_t)b * p.d.rows + row0) * dout;
  float* out2 = p.d.out_res ? p.d.out_res + ((size_t)b * p.d.rows + row0) * dout : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i;
    if (r >= cnt) continue;
    float* orow = out + (size_t)r * dout;
    const float* e0 = As + r * C::AS + p.koff[0];  // src_0 row (out_res = src_0 + out)
    if (C::VEC == 4 && p.out_vec_ok) {
#pragma unroll
      for (int v = 0; v < C::NV; ++v) {
        const int col = (tx + 16 * v) * 4;
        if (col < dout) {
          const float4 o =
              make_float4(acc[i][v * 4], acc[i][v * 4 + 1], acc[i][v * 4 + 2], acc[i][v * 4 + 3]);
          *reinterpret_cast<float4*>(orow + col) = o;
          if (out2)
            *reinterpret_cast<float4*>(out2 + (size_t)r * dout + col) =
                make_float4(o.x + e0[col], o.y + e0[col + 1], o.z + e0[col + 2], o.w + e0[col + 3]);
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < C::CPT; ++c) {
        const int col = col_of<DP>(tx, c);
        if (col < dout) {
          orow[col] = acc[i][c];
          if (out2) out2[(size_t)r * dout + col] = acc[i][c] + e0[col];
        }
      }
    }
  }
}

// -------------------------------------------------------------------- backward (dgrad)
template <int DP>
__global__ void __launch_bounds__(NT) rowmlp_bwd_kernel(const __grid_constant__ KParams p) {
  using C = Cfg<DP>;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;            // z, later reused as [a | dy | dh]
  float* Hs = As + TM * C::AS; // h_pre
  float* Ws = Hs + TM * C::HS;
  float* Aa = As;
  float* Dy = As + TM * C::HS;
  float* Dh = As + 2 * TM * C::HS;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int b = blockIdx.y;
  int row0, cnt, chunk;
  tile_range<TM>(p.d, blockIdx.x, row0, cnt, chunk);
  if (cnt <= 0) return;
  const int dh = p.d.d_hidden, dout = p.d.d_out;
  const float* w1 = p.d.w.w1 + (size_t)chunk * dh * p.k_total;
  const float* b1 = p.d.w.b1 + (size_t)chunk * dh;
  const float* w2 = p.d.w.w2 + (size_t)chunk * dout * dh;
  const float* b2 = p.d.w.b2 + (size_t)chunk * dout;
  const size_t grow0 = (size_t)b * p.d.rows + row0;  // first global row of the tile

  gather_tile<DP>(p, b, row0, cnt, As);

  float acc[4][C::CPT];
  // ---- recompute: h_pre, a
  gemm_tile<DP, true>(acc, As, C::AS, p.k_total, w1, p.k_total, dh, 0, Ws);
  __syncthreads();  // all reads of z done before As is reused
#pragma unroll
  for (int c = 0; c < C::CPT; ++c) {
    const int col = col_of<DP>(tx, c);
    const float bias = col < dh ? __ldg(b1 + col) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty * 4 + i;
      const float h = col < dh ? acc[i][c] + bias : 0.f;
      const float a = col < dh ? silu_f(h) : 0.f;
      Hs[r * C::HS + col] = h;
      Aa[r * C::HS + col] = a;
      if (r < cnt && col < dh) p.a_save[(grow0 + r) * dh + col] = a;
    }
  }
  // ---- recompute: y, LayerNorm statistics
  gemm_tile<DP, true>(acc, Aa, C::HS, dh, w2, dh, dout, 0, Ws);
#pragma unroll
  for (int c = 0; c < C::CPT; ++c) {
    const int col = col_of<DP>(tx, c);
    const float bias = col < dout ? __ldg(b2 + col) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][c] += bias;
  }
  // ---- dOut rows
  float dm[4][C::CPT];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i;
    const bool rv = r < cnt;
    int gi = 0;
    float gs = 1.f;
    if (rv && p.g1) {
      gi = p.g1_idx[row0 + r];
      if (p.g1_scale) gs = __ldg(p.g1_scale + gi);
    }
#pragma unroll
    for (int c = 0; c < C::CPT; ++c) {
      const int col = col_of<DP>(tx, c);
      float v = 0.f;
      if (rv && col < dout) {
        if (p.g0) v = __ldg(p.g0 + (grow0 + r) * dout + col);
        if (p.g1)
          v += gs * __ldg(p.g1 + (size_t)b * p.g1_batch_stride + (size_t)gi * dout + col);
      }
      dm[i][c] = v;
    }
  }
  // ---- LayerNorm backward -> dy (in acc)
  if (p.d.w.ln_g) {
    const float* g = p.d.w.ln_g + (size_t)chunk * dout;
    float mean[4], rstd[4];
    ln_stats<DP>(acc, tx, dout, mean, rstd);
    float pg[C::CPT], pb[C::CPT];
#pragma unroll
    for (int c = 0; c < C::CPT; ++c) pg[c] = 0.f, pb[c] = 0.f;
    float m1[4], m2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int c = 0; c < C::CPT; ++c) {
        const int col = col_of<DP>(tx, c);
        if (col < dout) {
          const float yh = (acc[i][c] - mean[i]) * rstd[i];
          const float dyh = dm[i][c] * __ldg(g + col);
          pg[c] += dm[i][c] * yh;
          pb[c] += dm[i][c];
          acc[i][c] = yh;  // keep y_hat
          s1 += dyh;
          s2 += dyh * yh;
        }
      }
      m1[i] = group16_sum(s1) / (float)dout;
      m2[i] = group16_sum(s2) / (float)dout;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < C::CPT; ++c) {
        const int col = col_of<DP>(tx, c);
        if (col < dout) {
          const float dyh = dm[i][c] * __ldg(g + col);
          acc[i][c] = rstd[i] * (dyh - m1[i] - acc[i][c] * m2[i]);
        } else {
          acc[i][c] = 0.f;
        }
      }
    // per-tile partial sums of dLN gamma / beta, reduced over the 16 row groups
    __syncthreads();  // Ws free (gemm2 done)
#pragma unroll
    for (int c = 0; c < C::CPT; ++c) {
      const int col = col_of<DP>(tx, c);
      Ws[ty * 2 * DP + col] = pg[c];
      Ws[ty * 2 * DP + DP + col] = pb[c];
    }
    __syncthreads();
    if (tid < 2 * DP) {
      const int which = tid / DP, col = tid % DP;
      if (col < dout) {
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < 16; ++t) s += Ws[t * 2 * DP + which * DP + col];
        p.ln_partial[(((size_t)b * gridDim.x + blockIdx.x) * 2 + which) * dout + col] = s;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < C::CPT; ++c) acc[i][c] = dm[i][c];
  }
  // rows beyond cnt carry dm == 0 -> dy == 0 (with or without LN)
#pragma unroll
  for (int c = 0; c < C::CPT; ++c) {
    const int col = col_of<DP>(tx, c);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty * 4 + i;
      const float v = (r < cnt && col < dout) ? acc[i][c] : 0.f;
      Dy[r * C::HS + col] = v;
      if (r < cnt && col < dout) p.dy_save[(grow0 + r) * dout + col] = v;
    }
  }
  // ---- da = dy . W2 ; dh = da * silu'(h_pre)
  gemm_tile<DP, false>(acc, Dy, C::HS, dout, w2, dh, dh, 0, Ws);
#pragma unroll
  for (int c = 0; c < C::CPT; ++c) {
    const int col = col_of<DP>(tx, c);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty * 4 + i;
      const float v =
          (r < cnt && col < dh) ? acc[i][c] * silu_grad_f(Hs[r * C::HS + col]) : 0.f;
      Dh[r * C::HS + col] = v;
      if (r < cnt && col < dh) p.dh_save[(grow0 + r) * dh + col] = v;
    }
  }
  // ---- dz = dh . W1, one DP-wide block of input columns at a time
  bool any = false;
  for (int s = 0; s < p.d.n_src; ++s) any |= (p.d_src[s] != nullptr);
  if (!any) return;
  for (int jb = 0; jb * DP < p.k_total; ++jb) {
    const int nval = min(DP, p.k_total - jb * DP);
    gemm_tile<DP, false>(acc, Dh, C::HS, dh, w1, p.k_total, nval, jb * DP, Ws);
#pragma unroll
    for (int c = 0; c < C::CPT; ++c) {
      const int kg = jb * DP + col_of<DP>(tx, c);
      if (kg >= p.k_total) continue;
      int s = 0;
      while (s + 1 < p.d.n_src && kg >= p.koff[s + 1]) ++s;
      float* dst = p.d_src[s];
      if (!dst) continue;
      const int w = p.d.src[s].width, cc = kg - p.koff[s];
      // residual path: out = src_s + mlp(...)  =>  d src_s += g0 (direct rows only)
      const bool res = (s == p.d.residual_src) && p.g0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = ty * 4 + i;
        if (r < cnt)
          dst[(grow0 + r) * w + cc] =
              acc[i][c] + (res ? __ldg(p.g0 + (grow0 + r) * dout + cc) : 0.f);
      }
    }
  }
}

// -------------------------------------------------------------------- wgrad
// partial[split][chunk][off + n*ldo + col0 + k] = sum_{rows of the split} G[row][n] * X[row][k]
struct WJob {
  const float* G;  // dense [batch*rows][ldg]
  int ldg, n;      // n = valid columns of G (output rows)
  nlam_src X;      // gathered (or dense) second operand
  int x_dense_rows;  // 1: X.ptr is dense [batch*rows][ld] (saved activations)
  int off, ldo, col0;
  int bias_off;    // >=0: also column sums of G
  int kblocks;     // ceil(width/64)
};
struct WParams {
  WJob job[NLAM_MAX_SRC + 1];
  int n_jobs;
  int batch, rows, n_chunks;
  const int32_t* chunk_ptr;
  int splits;
  int p_total;  // floats per (split, chunk)
  float* partial;
};

template <int NB>
__global__ void __launch_bounds__(NT) wgrad_kernel(const __grid_constant__ WParams p) {
  constexpr int RN = NB / 16;
  constexpr int GS = NB + 1, XS = 64 + 4;
  __shared__ __align__(16) float Gs[32 * GS];
  __shared__ __align__(16) float Xs[32 * XS];
  const int tid = threadIdx.x, tk = tid & 15, tn = tid >> 4;
  int jb = blockIdx.y, ji = 0;
  while (jb >= p.job[ji].kblocks) jb -= p.job[ji].kblocks, ++ji;
  const WJob& J = p.job[ji];
  const int chunk = blockIdx.z;
  const int c0 = p.chunk_ptr ? p.chunk_ptr[chunk] : 0;
  const int rc = (p.chunk_ptr ? p.chunk_ptr[chunk + 1] : p.rows) - c0;
  const long long total = (long long)p.batch * rc;
  long long per = (total + p.splits - 1) / p.splits;
  per = (per + 31) / 32 * 32;
  const long long f_begin = per * blockIdx.x;
  const long long f_end = min(total, f_begin + per);
  float acc[RN][4];
  float bsum[RN];
#pragma unroll
  for (int i = 0; i < RN; ++i) {
    bsum[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  }
  const int kbase = jb * 64;
  const int kw = min(64, J.X.width - kbase);
  for (long long f0 = f_begin; f0 < f_end; f0 += 32) {
    __syncthreads();
    for (int e = tid; e < 32 * NB; e += NT) {
      const int r = e / NB, n = e % NB;
      const long long f = f0 + r;
      float v = 0.f;
      if (f < f_end && n < J.n) {
        const long long bb = f / rc;
        const int row = c0 + (int)(f % rc);
        v = __ldg(J.G + (bb * p.rows + row) * J.ldg + n);
      }
      Gs[r * GS + n] = v;
    }
    for (int e = tid; e < 32 * 64; e += NT) {
      const int r = e >> 6, k = e & 63;
      const long long f = f0 + r;
      float v = 0.f;
      if (f < f_end && k < kw) {
        const long long bb = f / rc;
        const int row = c0 + (int)(f % rc);
        if (J.x_dense_rows) {
          v = __ldg(J.X.ptr + (bb * p.rows + row) * J.X.ld + kbase + k);
        } else {
          const int ridx = J.X.idx ? J.X.idx[row] : row;
          v = __ldg(J.X.ptr + bb * J.X.batch_stride + (long long)ridx * J.X.ld + kbase + k);
        }
      }
      Xs[r * XS + k] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
      const float4 x = *reinterpret_cast<const float4*>(Xs + r * XS + tk * 4);
#pragma unroll
      for (int i = 0; i < RN; ++i) {
        const float g = Gs[r * GS + tn + 16 * i];
        bsum[i] += g;
        acc[i][0] = fmaf(g, x.x, acc[i][0]);
        acc[i][1] = fmaf(g, x.y, acc[i][1]);
        acc[i][2] = fmaf(g, x.z, acc[i][2]);
        acc[i][3] = fmaf(g, x.w, acc[i][3]);
      }
    }
  }
  float* dst = p.partial + ((size_t)blockIdx.x * p.n_chunks + chunk) * p.p_total;
#pragma unroll
  for (int i = 0; i < RN; ++i) {
    const int n = tn + 16 * i;
    if (n >= J.n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = tk * 4 + j;
      if (k < kw) dst[J.off + (size_t)n * J.ldo + J.col0 + kbase + k] = acc[i][j];
    }
    if (J.bias_off >= 0 && jb == 0 && tk == 0) dst[J.bias_off + n] = bsum[i];
  }
}

// d_params[chunk][j] = sum over splits (fixed order); LN parts = sum over tiles.
struct RParams {
  const float* partial;
  int splits, n_chunks, p_total, p_main;  // p_main = floats covered by wgrad partials
  const float* ln_partial;
  int batch, n_tiles, d_out;
  const int32_t* tile_chunk;
  float* out;
  int accumulate;
  // optional second partial buffer holding only the bias / LayerNorm-affine
  // gradients: [vec_slots][n_chunks][vec_len], laid out [db1 | db2 | dLNg | dLNb]
  const float* vec_partial;
  int vec_slots, vec_len;
  ParamLayout lay;
};
// Few partial slots (the small launches of the hierarchical models): one thread per output
// element sums all slots itself, 256 outputs per block -- 8x fewer blocks, no shared memory.
constexpr int RP_WIDE_MAX = 32;
__host__ __device__ inline bool rp_wide(const RParams& p) {
  return p.splits <= RP_WIDE_MAX && p.vec_slots <= RP_WIDE_MAX && p.p_main == p.p_total;
}
__host__ __device__ inline int rp_blocks(const RParams& p) {
  return rp_wide(p) ? (p.p_total + 255) / 256 : (p.p_total + 31) / 32;
}
__device__ __forceinline__ int rp_vec_index(const RParams& p, int j) {
  int v = -1;
  if (p.vec_partial) {  // is j a bias / LayerNorm-affine entry?
    const ParamLayout& L = p.lay;
    if (j >= L.off_b1() && j < L.off_w2()) v = j - L.off_b1();
    else if (j >= L.off_b2() && j < L.off_b2() + L.d_out) v = L.d_hidden + j - L.off_b2();
    else if (L.has_ln && j >= L.off_lng()) v = L.d_hidden + L.d_out + j - L.off_lng();
  }
  return v;
}
__device__ __forceinline__ void reduce_params_block(const RParams& p, int bx, int chunk) {
  if (rp_wide(p)) {
    const int j = bx * 256 + threadIdx.x;
    if (j >= p.p_total) return;
    const int v = rp_vec_index(p, j);
    const float* src = v >= 0 ? p.vec_partial + (size_t)chunk * p.vec_len + v
                              : p.partial + (size_t)chunk * p.p_total + j;
    const size_t step = (size_t)p.n_chunks * (v >= 0 ? p.vec_len : p.p_total);
    const int n = v >= 0 ? p.vec_slots : p.splits;
    float x[RP_WIDE_MAX];
#pragma unroll
    for (int i = 0; i < RP_WIDE_MAX; ++i) x[i] = i < n ? src[(size_t)i * step] : 0.f;  // all in flight
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < RP_WIDE_MAX; ++i) t += x[i];  // slot order (deterministic)
    float* o = p.out + (size_t)chunk * p.p_total + j;
    *o = p.accumulate ? *o + t : t;
    return;
  }
  // Block = 32 consecutive output elements (coalesced 128-byte rows of the partial
  // matrix) x 8 warps; warp w sums partials w, w+8, ... (four independent sub-sums, so four
  // loads are in flight); the 8 sub-sums are then combined in a fixed order (deterministic).
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int j = bx * 32 + lane;
  float s = 0.f;
  if (j < p.p_total) {
    int v = -1;
    if (p.vec_partial) {  // is j a bias / LayerNorm-affine entry?
      const ParamLayout& L = p.lay;
      if (j >= L.off_b1() && j < L.off_w2()) v = j - L.off_b1();
      else if (j >= L.off_b2() && j < L.off_b2() + L.d_out) v = L.d_hidden + j - L.off_b2();
      else if (L.has_ln && j >= L.off_lng()) v = L.d_hidden + L.d_out + j - L.off_lng();
    }
    if (v >= 0) {
      for (int sp = w; sp < p.vec_slots; sp += 8)
        s += p.vec_partial[((size_t)sp * p.n_chunks + chunk) * p.vec_len + v];
    } else if (j < p.p_main) {
      const float* src = p.partial + (size_t)chunk * p.p_total + j;
      const size_t step = (size_t)p.n_chunks * p.p_total;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      int sp = w;
      for (; sp + 24 < p.splits; sp += 32) {
        s0 += src[(size_t)sp * step], s1 += src[(size_t)(sp + 8) * step];
        s2 += src[(size_t)(sp + 16) * step], s3 += src[(size_t)(sp + 24) * step];
      }
      for (; sp < p.splits; sp += 8) s0 += src[(size_t)sp * step];
      s = (s0 + s1) + (s2 + s3);
    } else {
      const int q = j - p.p_main;  // [2][d_out]
      const int n = p.batch * p.n_tiles;
      for (int i = w; i < n; i += 8) {
        const int t = i % p.n_tiles;
        if (p.tile_chunk && p.tile_chunk[t] != chunk) continue;
        s += p.ln_partial[(size_t)i * 2 * p.d_out + q];
      }
    }
  }
  red[w][lane] = s;
  __syncthreads();
  if (w == 0 && j < p.p_total) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][lane];
    float* o = p.out + (size_t)chunk * p.p_total + j;
    *o = p.accumulate ? *o + t : t;
  }
}
__global__ void __launch_bounds__(256) reduce_params_kernel(const __grid_constant__ RParams p) {
  pdl_wait();
  reduce_params_block(p, blockIdx.x, blockIdx.y);
}

// Deferred reductions (stage_mask bit 8): the jobs of many backward calls are queued on
// the host and run by ONE launch (nlam_rowmlp_bwd_flush) -- the per-MLP reductions are
// ~7 us kernels whose launch gaps and tails would otherwise add up over 19 MLPs.
constexpr int RB_MAX = 28;
struct RBatch {
  int n;
  int blk_off[RB_MAX + 1];
  RParams job[RB_MAX];
};
static_assert(sizeof(RBatch) <= 4000, "RBatch must fit the kernel parameter space");
__global__ void __launch_bounds__(256) reduce_params_batch_kernel(const __grid_constant__ RBatch b) {
  pdl_wait();
  int j = 0;
  while (j + 1 < b.n && (int)blockIdx.x >= b.blk_off[j + 1]) ++j;
  const RParams& p = b.job[j];
  const int local = blockIdx.x - b.blk_off[j];
  const int bpc = rp_blocks(p);
  reduce_params_block(p, local % bpc, local / bpc);
}

// g_rq[w] = wave w: reductions into distinct outputs (one launch); a reduction that
// accumulates into an output already queued (weights used several times in a step,
// e.g. an unrolled rollout) goes to the next wave, so the order of accumulation is
// that of the immediate mode
// One queue per (device, stream): a flush on one stream never runs (or drops) reductions
// whose producers were enqueued on another stream or GPU.
typedef std::vector<std::vector<RParams>> RWaves;
static std::mutex g_rq_mutex;
static std::map<std::pair<int, cudaStream_t>, RWaves> g_rqs;
static std::pair<int, cudaStream_t> rq_key(cudaStream_t st) {
  int dev = 0;
  cudaGetDevice(&dev);
  return {dev, st};
}
static int queue_or_launch_reduce(const RParams& rp, bool defer, cudaStream_t st) {
  if (defer) {
    std::lock_guard<std::mutex> lk(g_rq_mutex);
    RWaves& g_rq = g_rqs[rq_key(st)];
    size_t wave = 0;
    for (size_t w = 0; w < g_rq.size(); ++w)
      for (const RParams& q : g_rq[w])
        if (q.out == rp.out) wave = w + 1;
    if (wave >= g_rq.size()) g_rq.resize(wave + 1);
    g_rq[wave].push_back(rp);
    return 0;
  }
  dim3 rgrid(rp_blocks(rp), rp.n_chunks);
  NLAM_CUDA(launch_k(reduce_params_kernel, rgrid, 256, 0, st, rp));
  NLAM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int reduce_params_flush(cudaStream_t st) {
  RWaves waves;
  {
    std::lock_guard<std::mutex> lk(g_rq_mutex);
    auto it = g_rqs.find(rq_key(st));
    if (it == g_rqs.end()) return 0;
    waves.swap(it->second);
    g_rqs.erase(it);
  }
  for (const std::vector<RParams>& jobs : waves)
  for (size_t i = 0; i < jobs.size(); i += RB_MAX) {
    RBatch b{};
    b.n = (int)std::min<size_t>(RB_MAX, jobs.size() - i);
    int off = 0;
    for (int j = 0; j < b.n; ++j) {
      b.job[j] = jobs[i + j];
      b.blk_off[j] = off;
      off += rp_blocks(b.job[j]) * b.job[j].n_chunks;
    }
    b.blk_off[b.n] = off;
    if (off == 0) continue;
    NLAM_CUDA(launch_k(reduce_params_batch_kernel, off, 256, 0, st, b));
    NLAM_CUDA(cudaGetLastError());
    count_launch();
  }
  return 0;
}
int reduce_params_pending() {
  std::lock_guard<std::mutex> lk(g_rq_mutex);
  size_t n = 0;
  for (const auto& kv : g_rqs)
    for (const auto& w : kv.second) n += w.size();
  return (int)n;
}
int reduce_params_discard(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_rq_mutex);
  g_rqs.erase(rq_key(st));
  return 0;
}

int launch_reduce_params(const float* partial, int splits, int n_chunks, int p_total, float* out,
                         int accumulate, const float* vec_partial, int vec_slots, int vec_len,
                         ParamLayout lay, cudaStream_t st, bool defer) {
  RParams rp{};
  rp.accumulate = accumulate;
  rp.vec_partial = vec_partial, rp.vec_slots = vec_slots, rp.vec_len = vec_len, rp.lay = lay;
  rp.partial = partial, rp.splits = splits, rp.n_chunks = n_chunks;
  rp.p_total = p_total, rp.p_main = p_total;
  rp.out = out;
  return queue_or_launch_reduce(rp, defer, st);
}

// -------------------------------------------------------------------- host side
template <int DP>
static int launch_fwd(const KParams& p, cudaStream_t st) {
  using C = Cfg<DP>;
  NLAM_CUDA(ensure_dyn_smem((const void*)rowmlp_fwd_kernel<DP>, (int)C::SMEM));
  dim3 grid(n_tiles_of(p.d), p.d.batch);
  rowmlp_fwd_kernel<DP><<<grid, NT, C::SMEM, st>>>(p);
  NLAM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int simt_rowmlp_fwd(const nlam_rowmlp& d, cudaStream_t st) {
  KParams p{};
  if (fill_params(d, p)) return 1;
  if (d.rows == 0) return 0;
  NLAM_CHECK(d.out, "rowmlp: out is NULL");
  NLAM_CHECK(!d.out_bf16 && !d.out_res_bf16 && !d.agg.out_bf16,
             "rowmlp(fp32 path): bf16 shadow outputs need the tensor-core path");
  NLAM_CHECK(!d.agg.out && !d.out_idx,
             "rowmlp: fused aggregation / row scatter exist only on the bf16 tensor-core path");
  const int dp = pick_dp(d);
  switch (dp) {
    case 16: return launch_fwd<16>(p, st);
    case 32: return launch_fwd<32>(p, st);
    case 64: return launch_fwd<64>(p, st);
    case 128: return launch_fwd<128>(p, st);
  }
  set_error("rowmlp: widths above 128 are not supported (d_hidden=%d d_out=%d K=%d)", d.d_hidden,
            d.d_out, p.k_total);
  return 1;
}

static int wgrad_splits(const nlam_rowmlp& d, int k_total) {
  // enough CTAs for ~2 waves of 148 SMs, at least 64 rows per split
  int kblocks = 0;
  for (int s = 0; s < d.n_src; ++s) kblocks += (d.src[s].width + 63) / 64;
  kblocks += (d.d_hidden + 63) / 64;
  long long total = (long long)d.batch * d.rows;
  long long by_rows = (total + 63) / 64;
  long long want = (296 + kblocks * d.n_chunks - 1) / (kblocks * d.n_chunks);
  long long s = want < by_rows ? want : by_rows;
  if (s < 1) s = 1;
  if (s > 128) s = 128;
  return (int)s;
}

struct BwdWs {
  size_t a_save, dy_save, dh_save, ln_partial, partial, total;
};
static BwdWs bwd_ws(const nlam_rowmlp& d) {
  const size_t rows = (size_t)d.batch * d.rows;
  const int k = k_total_of(d);
  ParamLayout lay{k, d.d_hidden, d.d_out, d.w.ln_g != nullptr};
  auto al = [](size_t x) { return (x + 3) / 4 * 4; };
  BwdWs w;
  size_t o = 0;
  w.a_save = o, o += al(rows * d.d_hidden);
  w.dy_save = o, o += al(rows * d.d_out);
  w.dh_save = o, o += al(rows * d.d_hidden);
  w.ln_partial = o, o += al(lay.has_ln ? (size_t)d.batch * n_tiles_of(d) * 2 * d.d_out : 0);
  w.partial = o, o += al((size_t)wgrad_splits(d, k) * d.n_chunks * lay.total());
  w.total = o;
  return w;
}
size_t simt_rowmlp_bwd_workspace(const nlam_rowmlp& d) { return bwd_ws(d).total; }

template <int DP>
static int launch_bwd(const KParams& p, const WParams& wp, const RParams& rp, int mask,
                      cudaStream_t st) {
  using C = Cfg<DP>;
  NLAM_CUDA(ensure_dyn_smem((const void*)rowmlp_bwd_kernel<DP>, (int)C::SMEM));
  if (mask & 1) {
    dim3 grid(n_tiles_of(p.d), p.d.batch);
    rowmlp_bwd_kernel<DP><<<grid, NT, C::SMEM, st>>>(p);
    NLAM_CUDA(cudaGetLastError());
    count_launch();
  }
  if (mask & 2) {
    int kb = 0;
    for (int j = 0; j < wp.n_jobs; ++j) kb += wp.job[j].kblocks;
    dim3 wgrid(wp.splits, kb, wp.n_chunks);
    wgrad_kernel<DP><<<wgrid, NT, 0, st>>>(wp);
    NLAM_CUDA(cudaGetLastError());
    count_launch();
  }
  if (mask & 4) return queue_or_launch_reduce(rp, (mask & 8) != 0, st);
  return 0;
}

int simt_rowmlp_bwd(const nlam_rowmlp_bwd& bd, cudaStream_t st) {
  const nlam_rowmlp& d = bd.fwd;
  KParams p{};
  if (fill_params(d, p)) return 1;
  NLAM_CHECK(bd.d_params, "rowmlp_bwd: d_params is NULL");
  NLAM_CHECK(bd.g0_sum_count <= 1, "rowmlp_bwd: g0_sum_count exists only on the fused bf16 path");
  const ParamLayout lay = p.lay;
  if (d.rows == 0) {
    if (!bd.params_accumulate)
      NLAM_CUDA(cudaMemsetAsync(bd.d_params, 0, sizeof(float) * (size_t)d.n_chunks * lay.total(), st));
    return 0;
  }
  NLAM_CHECK(bd.g0 || bd.g1, "rowmlp_bwd: no output gradient given");
  NLAM_CHECK(!bd.g1 || bd.g1_idx, "rowmlp_bwd: g1 needs g1_idx");
  NLAM_CHECK(!bd.g0_idx && bd.reduce_src < 0 && !bd.d_src_idx[0] && !bd.d_src_idx[1] &&
                 !bd.d_src_idx[2],
             "rowmlp_bwd: row scatter / fused reduction exist only on the bf16 tensor-core path");
  const BwdWs ws = bwd_ws(d);
  NLAM_CHECK(bd.workspace && bd.workspace_floats >= ws.total,
             "rowmlp_bwd: workspace too small (%zu < %zu floats)", bd.workspace_floats, ws.total);
  NLAM_CHECK(((uintptr_t)bd.workspace) % 16 == 0, "rowmlp_bwd: workspace must be 16B aligned");
  p.g0 = bd.g0;
  p.g1 = bd.g1;
  p.g1_idx = bd.g1_idx;
  p.g1_scale = bd.g1_scale;
  p.g1_batch_stride = bd.g1_batch_stride;
  for (int s = 0; s < NLAM_MAX_SRC; ++s) p.d_src[s] = s < d.n_src ? bd.d_src[s] : nullptr;
  p.a_save = bd.workspace + ws.a_save;
  p.dy_save = bd.workspace + ws.dy_save;
  p.dh_save = bd.workspace + ws.dh_save;
  p.ln_partial = bd.workspace + ws.ln_partial;

  WParams wp{};
  wp.batch = d.batch, wp.rows = d.rows, wp.n_chunks = d.n_chunks, wp.chunk_ptr = d.chunk_ptr;
  wp.splits = wgrad_splits(d, p.k_total);
  wp.p_total = lay.total();
  wp.partial = bd.workspace + ws.partial;
  int nj = 0;
  for (int s = 0; s < d.n_src; ++s) {
    WJob& J = wp.job[nj++];
    J.G = p.dh_save, J.ldg = d.d_hidden, J.n = d.d_hidden;
    J.X = d.src[s], J.x_dense_rows = 0;
    J.off = lay.off_w1(), J.ldo = p.k_total, J.col0 = p.koff[s];
    J.bias_off = s == 0 ? lay.off_b1() : -1;
    J.kblocks = (d.src[s].width + 63) / 64;
  }
  {
    WJob& J = wp.job[nj++];
    J.G = p.dy_save, J.ldg = d.d_out, J.n = d.d_out;
    J.X = nlam_src{p.a_save, nullptr, 0, d.d_hidden, d.d_hidden};
    J.x_dense_rows = 1;
    J.off = lay.off_w2(), J.ldo = d.d_hidden, J.col0 = 0;
    J.bias_off = lay.off_b2();
    J.kblocks = (d.d_hidden + 63) / 64;
  }
  wp.n_jobs = nj;

  RParams rp{};
  rp.partial = wp.partial, rp.splits = wp.splits, rp.n_chunks = d.n_chunks;
  rp.p_total = lay.total(), rp.p_main = lay.off_b2() + d.d_out;
  rp.ln_partial = p.ln_partial, rp.batch = d.batch, rp.n_tiles = n_tiles_of(d);
  rp.d_out = d.d_out, rp.tile_chunk = d.tile_chunk, rp.out = bd.d_params;
  rp.accumulate = bd.params_accumulate;

  const int dp = pick_dp(d);
  const int mask = (bd.stage_mask & 7) ? bd.stage_mask : (7 | (bd.stage_mask & 8));
  switch (dp) {
    case 16: return launch_bwd<16>(p, wp, rp, mask, st);
    case 32: return launch_bwd<32>(p, wp, rp, mask, st);
    case 64: return launch_bwd<64>(p, wp, rp, mask, st);
    case 128: return launch_bwd<128>(p, wp, rp, mask, st);
  }
  set_error("rowmlp_bwd: widths above 128 are not supported");
  return 1;
}

}  // namespace nlam
