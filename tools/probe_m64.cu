// Probe: where does a cta_group::1 M=64 tcgen05.mma put its accumulator rows in TMEM,
// and can two M=64 accumulators share columns at lane offsets 0 and 16?
// nvcc -gencode arch=compute_100a,code=sm_100a -I neural-lam-dev_b200/csrc -o /tmp/probe tools/probe_m64.cu
#include <cstdio>
#include <vector>
#include "rowmlp_tc_bwd.cuh"
using namespace nlam::tc;

__global__ void probe(float* out, int lane_off2) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* sA = sm;            // [128 K rows][64 MN] bf16, SW128
  uint8_t* sB = sm + 16384;
  uint8_t* sA2 = sm + 32768;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 49152);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 49152 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0;
  __syncthreads();
  if (tid < 64) {  // row k = 0: A[0][m] = m + 1, B[0][n] = n + 1, A2[0][m] = 101 + m
    const int m = tid;
    reinterpret_cast<__nv_bfloat16*>(sA + sw128_off(0, m & ~7, 16384))[m & 7] = __float2bfloat16(m + 1);
    reinterpret_cast<__nv_bfloat16*>(sB + sw128_off(0, m & ~7, 16384))[m & 7] = __float2bfloat16(m + 1);
    reinterpret_cast<__nv_bfloat16*>(sA2 + sw128_off(0, m & ~7, 16384))[m & 7] = __float2bfloat16(101 + m);
  }
  if (warp == 0) tmem_alloc(slot, 64u);
  if (tid == 32) { mbar_init(bar, 1); mbar_fence_init(); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  // clear TMEM with an M=128 MMA of zeros (A = sA2 rows beyond... use B zero): D = 0
  if (tid == 0) {
    const uint32_t idesc128 = make_idesc_bf16(128, 64, 1, 1);
    // zero tile: sm + 32768 + 8192.. is zero except row 0 of sA2; use K step 1 (rows 16..31: all zero)
    umma_bf16(tb, make_desc_mn_sw128(smem_u32(sA) + 2048u, 16384), make_desc_mn_sw128(smem_u32(sB) + 2048u, 16384), idesc128, 0);
    const uint32_t idesc64 = make_idesc_bf16(64, 64, 1, 1);
    umma_bf16(tb, make_desc_mn_sw128(smem_u32(sA), 16384), make_desc_mn_sw128(smem_u32(sB), 16384), idesc64, 1);
    if (lane_off2 >= 0)
      umma_bf16(tb + ((uint32_t)lane_off2 << 16), make_desc_mn_sw128(smem_u32(sA2), 16384),
                make_desc_mn_sw128(smem_u32(sB), 16384), idesc64, 1);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  for (int c = 0; c < 64; c += 16) {
    float v[16];
    tmem_ld16(tb + lane_addr + c, v);
    for (int j = 0; j < 16; ++j) out[tid * 64 + c + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 64u);
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 64 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 50000);
  for (int off : {-1, 16}) {
    probe<<<1, 128, 50000>>>(d, off);
    cudaError_t e = cudaDeviceSynchronize();
    printf("lane_off2=%d: %s\n", off, cudaGetErrorString(e));
    std::vector<float> h(128 * 64);
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    for (int l = 0; l < 128; ++l) {
      // D[m][n] = (m+1)(n+1) -> col0 = m+1, col1 = 2(m+1)
      printf("%d:%g/%g ", l, h[l * 64], h[l * 64 + 1]);
      if (l % 8 == 7) printf("\n");
    }
  }
  return 0;
}
