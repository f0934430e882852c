// Probe (sm_100a): layout facts of the TMA row gather / scatter used by the v2 row-MLP kernels.
//  1. cp.async.bulk.tensor.2d tile::gather4 of 4 rows x 64 bf16 (box {64,1}, SWIZZLE_128B) into
//     consecutive 128-byte shared-memory rows: is the result the UMMA K-major SW128 tile
//     (16-byte chunk c of tile row r at r*128 + ((c ^ (r & 7)) << 4))?
//  2. tile::scatter4 store of an fp32 [rows][32] SW128 shared-memory tile to rows of a global
//     fp32 [R][64] matrix (box {32,1}) at a column offset.
//  3. plain tile store (box {32,128}) and the reduce-add tile store.
// nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/probe_g4 tools/probe_gather4.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out,
                      const __grid_constant__ CUtensorMap map_tile, const int* __restrict__ idx,
                      const int* __restrict__ oidx, uint8_t* raw, int mode) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 32768);
  const int tid = threadIdx.x;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid < 32) {
    if (tid == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(16384u)
                   : "memory");
    __syncwarp();
    // lane l gathers tile rows 4l .. 4l+3
    const int r0 = idx[4 * tid], r1 = idx[4 * tid + 1], r2 = idx[4 * tid + 2], r3 = idx[4 * tid + 3];
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(sm + tid * 512)),
        "l"(&map_in), "r"(smem_u32(bar)), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
        : "memory");
  }
  // wait
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(ok)
          : "r"(smem_u32(bar)), "r"(0u)
          : "memory");
  }
  for (int i = tid; i < 16384; i += blockDim.x) raw[i] = sm[i];
  __syncthreads();

  // ---- store side: fp32 [128][32] SW128 tile, value = 1000 * row + col (+ 32 * half)
  float* st = reinterpret_cast<float*>(sm);
  for (int half = 0; half < 2; ++half) {
    for (int u = tid; u < 128 * 8; u += blockDim.x) {
      const int r = u >> 3, c = u & 7;  // 16-byte chunk c = columns 4c..4c+3
      float4 v = make_float4(1000.f * r + 32 * half + 4 * c, 1000.f * r + 32 * half + 4 * c + 1,
                             1000.f * r + 32 * half + 4 * c + 2, 1000.f * r + 32 * half + 4 * c + 3);
      *reinterpret_cast<float4*>(sm + half * 16384 + r * 128 + ((c ^ (r & 7)) << 4)) = v;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (mode == 0 && tid < 32) {
    const int r0 = oidx[4 * tid], r1 = oidx[4 * tid + 1], r2 = oidx[4 * tid + 2], r3 = oidx[4 * tid + 3];
    for (int half = 0; half < 2; ++half)
      asm volatile(
          "cp.async.bulk.tensor.2d.global.shared::cta.tile::scatter4.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
              &map_out),
          "r"(smem_u32(sm + half * 16384 + tid * 512)), "r"(32 * half), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
          : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  if (mode >= 1 && tid == 0) {
    // plain tile store of rows [256, 384) (mode 1) or reduce-add (mode 2)
    for (int half = 0; half < 2; ++half) {
      if (mode == 1)
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                         &map_tile),
                     "r"(smem_u32(sm + half * 16384)), "r"(32 * half), "r"(256)
                     : "memory");
      else
        asm volatile(
            "cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                &map_tile),
            "r"(smem_u32(sm + half * 16384)), "r"(32 * half), "r"(256)
            : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main() {
  const int R = 1000;
  EncodeTiled enc = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres) != cudaSuccess || !enc) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  std::vector<uint16_t> h((size_t)R * 64);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < 64; ++c) h[(size_t)r * 64 + c] = (uint16_t)((r * 64 + c) & 0xffff);
  uint16_t* d_in;
  float* d_out;
  cudaMalloc(&d_in, h.size() * 2);
  cudaMemcpy(d_in, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  cudaMalloc(&d_out, (size_t)R * 64 * 4);
  std::vector<int> idx(128), oidx(128);
  for (int i = 0; i < 128; ++i) idx[i] = (i * 37 + 11) % R, oidx[i] = (i * 7 + 3) % R;
  int *d_idx, *d_oidx;
  cudaMalloc(&d_idx, 512);
  cudaMalloc(&d_oidx, 512);
  cudaMemcpy(d_idx, idx.data(), 512, cudaMemcpyHostToDevice);
  cudaMemcpy(d_oidx, oidx.data(), 512, cudaMemcpyHostToDevice);
  uint8_t* d_raw;
  cudaMalloc(&d_raw, 16384);

  CUtensorMap map_in, map_out, map_tile;
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)R}, strides[1] = {128};
    cuuint32_t box[2] = {64, 1}, es[2] = {1, 1};
    CUresult r = enc(&map_in, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_in, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode in: %d\n", (int)r);
  }
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)R}, strides[1] = {256};
    cuuint32_t box[2] = {32, 1}, es[2] = {1, 1};
    CUresult r = enc(&map_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_out, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode out: %d\n", (int)r);
    cuuint32_t box2[2] = {32, 128};
    r = enc(&map_tile, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_out, dims, strides, box2, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode tile: %d\n", (int)r);
  }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 34000);
  int bad_total = 0;
  for (int mode = 0; mode < 3; ++mode) {
    if (mode < 2) cudaMemset(d_out, 0, (size_t)R * 64 * 4);
    probe<<<1, 128, 34000>>>(map_in, map_out, map_tile, d_idx, d_oidx, d_raw, mode);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %d: %s\n", mode, cudaGetErrorString(e));
    if (e != cudaSuccess) return 2;
    if (mode == 0) {
      std::vector<uint8_t> raw(16384);
      cudaMemcpy(raw.data(), d_raw, 16384, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r = 0; r < 128; ++r)
        for (int c = 0; c < 8; ++c) {
          const uint16_t* got = reinterpret_cast<const uint16_t*>(raw.data() + r * 128 + ((c ^ (r & 7)) << 4));
          for (int j = 0; j < 8; ++j)
            if (got[j] != h[(size_t)idx[r] * 64 + c * 8 + j]) ++bad;
        }
      printf("gather4 SW128 layout mismatches: %d of 8192\n", bad);
      if (bad) {  // show what row 1 looks like
        const uint16_t* g = reinterpret_cast<const uint16_t*>(raw.data() + 128);
        for (int j = 0; j < 64; ++j) printf("%d ", (int)g[j] - (idx[1] * 64 & 0xffff));
        printf("\n");
      }
      bad_total += bad;
    }
    std::vector<float> o((size_t)R * 64);
    cudaMemcpy(o.data(), d_out, o.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    if (mode == 0) {
      for (int i = 0; i < 128; ++i)
        for (int c = 0; c < 64; ++c)
          if (o[(size_t)oidx[i] * 64 + c] != 1000.f * i + c) ++bad;
      printf("scatter4 mismatches: %d of 8192\n", bad);
    } else {
      const float mul = mode == 1 ? 1.f : 2.f;  // mode 2 adds the tile to what mode 1 stored
      for (int i = 0; i < 128; ++i)
        for (int c = 0; c < 64; ++c)
          if (o[(size_t)(256 + i) * 64 + c] != mul * (1000.f * i + c)) ++bad;
      printf("%s mismatches: %d of 8192\n", mode == 1 ? "tile store" : "reduce-add store", bad);
    }
    bad_total += bad;
  }
  printf(bad_total ? "PROBE FAILED\n" : "PROBE OK\n");
  return bad_total ? 3 : 0;
}
