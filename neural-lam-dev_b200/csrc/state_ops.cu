// Auto-regressive state update fused with its loss term, forward and backward.
//   pred = prev + net_out * diff_std + diff_mean          (base_graph_model.py:174-177)
//   new  = interior ? pred : truth                        (ar_model.py:244-247)
//   loss_sum = sum_{rows, f} interior * ((new - truth) * inv_std_f)^2   (metrics.py:21-84;
//              the caller divides by batch * steps * #interior nodes, ar_model.py:294-298)
// One pass over the (batch*nodes, F) tensors instead of ~11 elementwise/reduction
// launches; per-block partial sums are combined in a fixed order (deterministic).
#include "common.cuh"

namespace nlam {

constexpr int SS_THREADS = 256;
constexpr int SS_ROWS_PER_BLOCK = 256;  // one row per thread

__global__ void __launch_bounds__(SS_THREADS) state_step_fwd_kernel(const nlam_state_step p) {
  __shared__ float red[SS_THREADS / 32];
  pdl_wait();
  const long long row = (long long)blockIdx.x * SS_ROWS_PER_BLOCK + threadIdx.x;
  float acc = 0.f;
  if (row < p.rows) {
    const int node = (int)(row % p.nodes);
    const float in = __ldg(p.interior + node);
    const float* no = p.net_out + row * p.features;
    const long long bi = row / p.nodes, dense = (long long)p.nodes * p.features;
    const float* pv = p.prev + bi * (p.prev_batch_stride ? p.prev_batch_stride : dense) +
                      (long long)node * p.features;
    const float* tr = p.truth + bi * (p.truth_batch_stride ? p.truth_batch_stride : dense) +
                      (long long)node * p.features;
    float* ns = p.new_state + row * p.features;
    for (int f = 0; f < p.features; ++f) {
      const float t = __ldg(tr + f);
      const float pred = __ldg(pv + f) + __ldg(no + f) * __ldg(p.diff_std + f) + __ldg(p.diff_mean + f);
      const float nw = in != 0.f ? pred : t;
      ns[f] = nw;
      const float e = (nw - t) * (p.inv_std ? __ldg(p.inv_std + f) : 1.f);
      acc += in * e * e;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < SS_THREADS / 32; ++i) s += red[i];
    p.loss_partial[blockIdx.x] = s;
  }
}

// loss_sum[0] = sum of the per-block partials, fixed order (single block)
__global__ void __launch_bounds__(256) state_step_sum_kernel(const float* partial, int n, float* out) {
  __shared__ float red[256];
  pdl_wait();
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}

// d_pred = interior * (d_new + d_loss * 2 * (new - truth) * inv_std^2)
// d_net_out = d_pred * diff_std ; d_prev = d_pred
__global__ void __launch_bounds__(SS_THREADS) state_step_bwd_kernel(const nlam_state_step_bwd p) {
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = p.fwd.rows * p.fwd.features;
  if (i >= total) return;
  const long long row = i / p.fwd.features;
  const int f = (int)(i % p.fwd.features);
  const float in = __ldg(p.fwd.interior + (int)(row % p.fwd.nodes));
  float d = 0.f;
  if (in != 0.f) {
    if (p.d_new) d = __ldg(p.d_new + i);
    if (p.d_loss) {
      const float is = p.fwd.inv_std ? __ldg(p.fwd.inv_std + f) : 1.f;
      const long long bi = row / p.fwd.nodes, node = row % p.fwd.nodes;
      const long long ts = p.fwd.truth_batch_stride ? p.fwd.truth_batch_stride
                                                    : (long long)p.fwd.nodes * p.fwd.features;
      const float tr = __ldg(p.fwd.truth + bi * ts + node * p.fwd.features + f);
      d += __ldg(p.d_loss) * 2.f * (__ldg(p.fwd.new_state + i) - tr) * is * is;
    }
  }
  if (p.d_net_out) p.d_net_out[i] = d * __ldg(p.fwd.diff_std + f);
  if (p.d_prev) p.d_prev[i] = d;
}

}  // namespace nlam

using namespace nlam;

extern "C" int64_t nlam_state_step_partials(int64_t rows) {
  return (rows + SS_ROWS_PER_BLOCK - 1) / SS_ROWS_PER_BLOCK;
}

extern "C" int nlam_state_step_fwd(const nlam_state_step* d, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  NLAM_CHECK(d && d->net_out && d->prev && d->truth && d->diff_std && d->diff_mean && d->interior &&
                 d->new_state && d->loss_partial && d->loss_sum,
             "state_step_fwd: NULL argument");
  NLAM_CHECK(d->rows > 0 && d->nodes > 0 && d->features > 0 && d->rows % d->nodes == 0,
             "state_step_fwd: bad sizes");
  const int nb = (int)nlam_state_step_partials(d->rows);
  NLAM_CUDA(launch_k(state_step_fwd_kernel, nb, SS_THREADS, 0, st, *d));
  NLAM_CUDA(cudaGetLastError());
  count_launch();
  NLAM_CUDA(launch_k(state_step_sum_kernel, 1, 256, 0, st, (const float*)d->loss_partial, nb,
                     d->loss_sum));
  NLAM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int nlam_state_step_bwd_run(const nlam_state_step_bwd* d, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  NLAM_CHECK(d && d->fwd.new_state && d->fwd.truth && d->fwd.interior && d->fwd.diff_std,
             "state_step_bwd: NULL argument");
  NLAM_CHECK(d->d_net_out || d->d_prev, "state_step_bwd: nothing to compute");
  const long long total = d->fwd.rows * d->fwd.features;
  NLAM_CUDA(launch_k(state_step_bwd_kernel, (int)((total + SS_THREADS - 1) / SS_THREADS),
                     SS_THREADS, 0, st, *d));
  NLAM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
