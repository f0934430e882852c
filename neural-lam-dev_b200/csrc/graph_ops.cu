// Integer graph plumbing and the deterministic segment sum.
//
// nlam_csr_build: stable counting sort of edge ids by receiver (or sender)
//   id -- the CSR / transposed CSR that replaces PyG's scatter index handling
//   for the `edge_index` of interaction_net.py:55-61.  Result is bit-identical
//   to numpy argsort(kind="stable") + bincount/cumsum.
// nlam_segsum: out[b,i,:] (+)= scale[i] * sum_{p in seg(i)} src[b, idx[p], :]
//   in list order (== ascending edge id), one thread per 128-bit column group,
//   no float atomics: PyG scatter sum/mean (interaction_net.py:124-131) and the
//   backward of the x_j / x_i gathers.
#include "common.cuh"

namespace nlam {

__global__ void csr_count_kernel(const int32_t* __restrict__ key, long long m, int32_t* cnt) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) atomicAdd(cnt + key[i], 1);
}

// single CTA exclusive scan: ptr[0..n] from cnt[0..n-1]; cnt is then reused as cursor (zeroed)
__global__ void csr_scan_kernel(int32_t* cnt, int32_t n, int32_t* ptr, float* inv_deg) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + tid;
    const int32_t v = i < n ? cnt[i] : 0;
    int32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int32_t t = lane < (int)(blockDim.x >> 5) ? warp_tot[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int32_t y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      warp_tot[lane] = t;
    }
    __syncthreads();
    const int32_t carry = carry_s;
    const int32_t incl = x + (wid > 0 ? warp_tot[wid - 1] : 0) + carry;
    if (i < n) {
      ptr[i] = incl - v;
      cnt[i] = 0;
      if (inv_deg) inv_deg[i] = 1.0f / (float)(v > 1 ? v : 1);
    }
    __syncthreads();
    if (tid == blockDim.x - 1) carry_s = incl;
    __syncthreads();
  }
  if (tid == 0) ptr[n] = carry_s;
}

__global__ void csr_fill_kernel(const int32_t* __restrict__ key, long long m,
                                const int32_t* __restrict__ ptr, int32_t* cursor, int32_t* perm) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) {
    const int32_t k = key[i];
    const int32_t pos = atomicAdd(cursor + k, 1);
    perm[ptr[k] + pos] = (int32_t)i;
  }
}

// restore ascending edge-id order inside every segment (insertion sort; segments are short)
__global__ void csr_sort_kernel(const int32_t* __restrict__ ptr, int32_t n, int32_t* perm) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = ptr[i], e = ptr[i + 1];
  for (int a = b + 1; a < e; ++a) {
    const int32_t v = perm[a];
    int c = a - 1;
    while (c >= b && perm[c] > v) {
      perm[c + 1] = perm[c];
      --c;
    }
    perm[c + 1] = v;
  }
}

template <int VEC>
__global__ void segsum_kernel(const __grid_constant__ nlam_segsum p) {
  pdl_wait();
  const int wv = p.width / VEC;  // column groups per row
  const long long total = (long long)p.batch * p.n_out * wv;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int cg = (int)(e % wv);
  const long long row = e / wv;
  const int i = (int)(row % p.n_out);
  const int b = (int)(row / p.n_out);
  const float* src = p.src + (long long)b * p.src_batch_stride + cg * VEC;
  const int beg = p.ptr[i], end = p.ptr[i + 1];
  float acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
  for (int q = beg; q < end; ++q) {
    const float* s = src + (long long)p.idx[q] * p.width;
    if (VEC == 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(s));
      acc[0] += t.x, acc[1] += t.y, acc[2] += t.z, acc[3] += t.w;
    } else {
      acc[0] += __ldg(s);
    }
  }
  const float sc = p.scale ? p.scale[i] : 1.f;
  float* o = p.out + ((long long)b * p.n_out + i) * p.width + cg * VEC;
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    float r = p.scale ? acc[v] * sc : acc[v];
    o[v] = p.accumulate ? o[v] + r : r;
  }
}

}  // namespace nlam

using namespace nlam;

extern "C" int nlam_csr_build(const int32_t* key, int64_t n_edges, int32_t n_keys, int32_t* ptr,
                              int32_t* perm, float* inv_deg, int32_t* workspace, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  NLAM_CHECK(n_keys >= 0 && n_edges >= 0 && n_edges < (1ll << 31), "csr_build: bad sizes");
  NLAM_CHECK(ptr && workspace && (n_edges == 0 || (key && perm)), "csr_build: NULL argument");
  NLAM_CUDA(cudaMemsetAsync(workspace, 0, sizeof(int32_t) * (size_t)(n_keys + 1), st));
  const int nb = (int)((n_edges + 255) / 256);
  if (n_edges > 0) {
    csr_count_kernel<<<nb, 256, 0, st>>>(key, n_edges, workspace);
    NLAM_CUDA(cudaGetLastError());
    count_launch();
  }
  csr_scan_kernel<<<1, 1024, 0, st>>>(workspace, n_keys, ptr, inv_deg);
  NLAM_CUDA(cudaGetLastError());
  count_launch();
  if (n_edges > 0) {
    csr_fill_kernel<<<nb, 256, 0, st>>>(key, n_edges, ptr, workspace, perm);
    NLAM_CUDA(cudaGetLastError());
    count_launch();
    csr_sort_kernel<<<(n_keys + 127) / 128, 128, 0, st>>>(ptr, n_keys, perm);
    NLAM_CUDA(cudaGetLastError());
    count_launch();
  }
  return 0;
}

extern "C" int nlam_segsum_run(const nlam_segsum* d, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  NLAM_CHECK(d && d->out && d->ptr && d->width > 0 && d->batch >= 1, "segsum: bad arguments");
  if (d->n_out == 0) return 0;
  NLAM_CHECK(d->src && d->idx, "segsum: NULL src/idx");
  const bool vec = (d->width % 4 == 0) && (((uintptr_t)d->src) % 16 == 0) &&
                   (d->src_batch_stride % 4 == 0);
  const long long total = (long long)d->batch * d->n_out * (vec ? d->width / 4 : d->width);
  const int nb = (int)((total + 255) / 256);
  if (vec)
    NLAM_CUDA(launch_k(segsum_kernel<4>, nb, 256, 0, st, *d));
  else
    NLAM_CUDA(launch_k(segsum_kernel<1>, nb, 256, 0, st, *d));
  NLAM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
