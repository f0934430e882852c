// Device-side data feed: standardisation, init/target slicing, forcing windowing and batch
// assembly of /root/reference/neural_lam/weather_dataset.py:163-496 (WeatherDataset
// ._slice_state_time, ._slice_forcing_time, ._build_item_dataarrays, .__getitem__ for analysis
// data) on a time series that stays resident in HBM -- the reference does this per sample
// with xarray on CPU workers and ships every batch through pinned host memory.
//
// Both kernels are pure data movement (HBM-bound): coalesced row copies for the state part,
// a small (window x feature) transposition per grid node for the forcing part.
#include "common.cuh"

namespace nlam {

__global__ void __launch_bounds__(256)
feed_standardize_kernel(const float* __restrict__ src, const float* __restrict__ mean,
                        const float* __restrict__ std, float* __restrict__ dst, long long total,
                        int d) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int c = (int)(i % d);
    dst[i] = (src[i] - __ldg(mean + c)) / __ldg(std + c);
  }
}

__device__ __forceinline__ long long slot_of(long long t, int ring) {
  return ring > 0 ? t % ring : t;
}

// VEC: the state time slices are 16-byte aligned and a multiple of 4 floats long -> 128-bit
// copies.
template <bool VEC>
__global__ void __launch_bounds__(256) feed_batch_kernel(const __grid_constant__ nlam_feed_batch p) {
  const int T = 2 + p.ar_steps;
  const int W = p.past + p.future + 1;
  const long long row_s = (long long)p.n_grid * p.d_state;  // floats of one state time slice
  const long long stride = (long long)gridDim.x * 256;
  const int p2 = p.past > 2 ? p.past - 2 : 0;   // max(0, past - init_steps)   (:219-222)
  const int o2 = p.past > 2 ? p.past : 2;       // max(init_steps, past)       (:289)
  // ---- init / target states: time slices start .. start + 2 + ar_steps, contiguous copies
  if (VEC) {
    const long long row4 = row_s >> 2, n4 = (long long)p.batch * T * row4;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += stride) {
      const long long bt = i / row4, r = i - bt * row4;
      const int b = (int)(bt / T), t = (int)(bt - (long long)b * T);
      const long long ts = p.sample_idx[b] + p2 + t;
      const float4 v = __ldg(reinterpret_cast<const float4*>(p.state + slot_of(ts, p.ring_cap) * row_s) + r);
      float* dst = t < 2 ? p.init_states + ((long long)b * 2 + t) * row_s
                         : p.target_states + ((long long)b * p.ar_steps + (t - 2)) * row_s;
      reinterpret_cast<float4*>(dst)[r] = v;
    }
  } else {
    const long long n_state = (long long)p.batch * T * row_s;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n_state; i += stride) {
      const long long bt = i / row_s, r = i - bt * row_s;
      const int b = (int)(bt / T), t = (int)(bt - (long long)b * T);
      const long long ts = p.sample_idx[b] + p2 + t;
      const float v = __ldg(p.state + slot_of(ts, p.ring_cap) * row_s + r);
      if (t < 2) p.init_states[((long long)b * 2 + t) * row_s + r] = v;
      else p.target_states[((long long)b * p.ar_steps + (t - 2)) * row_s + r] = v;
    }
  }
  // ---- forcing, windowed: out[b, s, n, f * W + w] = F[offset + s - past + w, n, f]
  //      (stack(forcing_feature_windowed=("forcing_feature", "window")), :417-420)
  if (p.forcing) {  // one thread per output element: coalesced writes, L1-served reads
    const long long row_f = (long long)p.n_grid * p.d_forcing;
    const long long dfw = (long long)p.d_forcing * W;
    const long long n_forc = (long long)p.batch * p.ar_steps * p.n_grid * dfw;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n_forc; i += stride) {
      const long long node_i = i / dfw;
      const int c = (int)(i - node_i * dfw);
      const int f = c / W, w = c - f * W;
      const long long bs = node_i / p.n_grid;
      const int n = (int)(node_i - bs * p.n_grid);
      const int b = (int)(bs / p.ar_steps), s = (int)(bs - (long long)b * p.ar_steps);
      const long long tf = p.sample_idx[b] + o2 + s - p.past + w;
      p.forcing_out[i] = __ldg(p.forcing + slot_of(tf, p.ring_cap) * row_f + (long long)n * p.d_forcing + f);
    }
  }
  // ---- target times (:478-481)
  if (p.times && p.target_times)
    for (int i = blockIdx.x * 256 + threadIdx.x; i < p.batch * p.ar_steps; i += (int)stride) {
      const int b = i / p.ar_steps, s = i - b * p.ar_steps;
      p.target_times[i] = p.times[slot_of(p.sample_idx[b] + p2 + 2 + s, p.ring_cap)];
    }
}

}  // namespace nlam

using namespace nlam;

extern "C" int nlam_feed_standardize(const float* src, const float* mean, const float* std,
                                     float* dst, int64_t rows, int32_t d, void* stream) {
  NLAM_CHECK(src && mean && std && dst && rows >= 0 && d > 0, "feed_standardize: bad arguments");
  const long long total = (long long)rows * d;
  if (total == 0) return 0;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  feed_standardize_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, mean, std, dst, total, d);
  NLAM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int nlam_feed_batch_run(const nlam_feed_batch* d, void* stream) {
  NLAM_CHECK(d, "feed_batch: NULL descriptor");
  NLAM_CHECK(d->state && d->init_states && d->target_states, "feed_batch: missing state buffers");
  NLAM_CHECK(d->batch >= 1 && d->batch <= NLAM_FEED_MAX_BATCH, "feed_batch: batch %d not in 1..%d",
             d->batch, NLAM_FEED_MAX_BATCH);
  NLAM_CHECK(d->ar_steps >= 1 && d->past >= 0 && d->future >= 0 && d->n_grid > 0 && d->d_state > 0,
             "feed_batch: bad sizes");
  NLAM_CHECK(!d->forcing || (d->forcing_out && d->d_forcing > 0), "feed_batch: forcing needs forcing_out");
  // the window of every sample must lie inside the resident time range [t_lo, t_hi)
  const int o2 = d->past > 2 ? d->past : 2;
  for (int b = 0; b < d->batch; ++b) {
    const long long i = d->sample_idx[b];
    NLAM_CHECK(i >= d->t_lo && i + o2 + d->ar_steps + d->future <= d->t_hi,
               "feed_batch: sample %lld needs time steps [%lld, %lld), resident are [%lld, %lld)", i, i,
               i + o2 + d->ar_steps + d->future, (long long)d->t_lo, (long long)d->t_hi);
  }
  const long long total = (long long)d->batch * (2 + d->ar_steps) * d->n_grid *
                          (d->d_state + (long long)d->d_forcing * (d->past + d->future + 1));
  long long blocks = (total + 255) / 256 / 4 + 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
  const long long row_s = (long long)d->n_grid * d->d_state;
  const bool vec = row_s % 4 == 0 && ((uintptr_t)d->state) % 16 == 0 &&
                   ((uintptr_t)d->init_states) % 16 == 0 && ((uintptr_t)d->target_states) % 16 == 0;
  if (vec) feed_batch_kernel<true><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(*d);
  else feed_batch_kernel<false><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(*d);
  NLAM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
