// Shared declarations for libnlam_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/nlam_b200.h"

namespace nlam {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// Opt a kernel in for `bytes` of dynamic shared memory on the CURRENT device (cached per
// kernel and device: a process may drive several GPUs).
cudaError_t ensure_dyn_smem(const void* kern, int bytes);
int option_fwd_mc();    // -1 auto, 0 off, 1 force
int option_dgrad_mc();
int option_bwd_fused();
int option_tma();        // 1: TMA row-gather kernels where every operand has a bf16 shadow (default 0:
                         // measured 5 % slower than the register gather of the shadows)
int option_bwd_nh();     // fused backward: threads per tile row, 2 (default) or 4 (measured 7 % slower)
int option_wide128();    // 1 (default): d = 128 kernels with 512 threads per CTA
int option_rb128();      // z blocks per gather round of the d = 128 input-gradient kernel: 4 (default) or 2
int option_bwd_spread();  // fused backward launches of <= 148 tiles: 1 = one tile per CTA (HiLAM +3.2 %),
                          // 2 (default) = and on the single-context kernel with 4 threads per row (+4.9 %)
int option_small512();    // 1 (default): 512-thread forward CTAs for d = 64 launches of <= 148 tiles (HiLAM +2.6 %)
int option_fp32_split();  // 1 (default): fp32 mode on the tensor cores (split bf16 operands) where the tiles fit
int option_pdl();        // 1: launch with programmatic dependent launch (see launch_k)  // -1 / 1: fused dgrad + wgrad kernel where eligible, 0: two kernels

// Programmatic dependent launch: a kernel launched with this attribute may start while
// its predecessor in the stream is still running (once every CTA of the predecessor has
// executed griddepcontrol.launch_dependents or exited); it must execute pdl_wait() before
// it touches anything the predecessor writes.  Used to overlap a kernel's prologue (TMEM
// allocation, weights -> bf16 shared-memory operands) with the tail of the previous one.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                            cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = option_pdl() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

#define NLAM_CHECK(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      nlam::set_error(__VA_ARGS__);  \
      return 1;                      \
    }                                \
  } while (0)

#define NLAM_CUDA(expr)                                                        \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) {                                                   \
      nlam::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                      __FILE__, __LINE__);                                     \
      return 1;                                                                \
    }                                                                          \
  } while (0)

// ---- layout of the flat per-chunk parameter(-gradient) block
struct ParamLayout {
  int k_total, d_hidden, d_out, has_ln;
  __host__ __device__ int off_w1() const { return 0; }
  __host__ __device__ int off_b1() const { return d_hidden * k_total; }
  __host__ __device__ int off_w2() const { return off_b1() + d_hidden; }
  __host__ __device__ int off_b2() const { return off_w2() + d_out * d_hidden; }
  __host__ __device__ int off_lng() const { return off_b2() + d_out; }
  __host__ __device__ int off_lnb() const { return off_lng() + d_out; }
  __host__ __device__ int total() const { return off_b2() + d_out + (has_ln ? 2 * d_out : 0); }
};

inline int k_total_of(const nlam_rowmlp& d) {
  int k = 0;
  for (int s = 0; s < d.n_src; ++s) k += d.src[s].width;
  return k;
}

inline int pick_dp(const nlam_rowmlp& d) {
  int need = d.d_hidden > d.d_out ? d.d_hidden : d.d_out;
  int k3 = (k_total_of(d) + 2) / 3;
  if (k3 > need) need = k3;
  for (int s = 0; s < d.n_src; ++s)
    if (d.src[s].width > need) need = d.src[s].width;
  if (need <= 16) return 16;
  if (need <= 32) return 32;
  if (need <= 64) return 64;
  if (need <= 128) return 128;
  return -1;
}

// fp32 SIMT path
int simt_rowmlp_fwd(const nlam_rowmlp& d, cudaStream_t st);
int simt_rowmlp_bwd(const nlam_rowmlp_bwd& d, cudaStream_t st);
size_t simt_rowmlp_bwd_workspace(const nlam_rowmlp& d);

// bf16 tcgen05 path
namespace tc {
bool tc_supported(const nlam_rowmlp& d);
bool tc_split_supported(const nlam_rowmlp& d);  // precision fp32 on the tensor cores
}
int tc_rowmlp_fwd(const nlam_rowmlp& d, cudaStream_t st);
int tc_rowmlp_bwd(const nlam_rowmlp_bwd& d, cudaStream_t st);
size_t tc_rowmlp_bwd_workspace(const nlam_rowmlp& d);
bool tc_rowmlp_bwd_is_fused(const nlam_rowmlp_bwd& d);
int reduce_params_flush(cudaStream_t st);  // rowmlp_simt.cu: deferred partial reductions
int reduce_params_pending();
int reduce_params_discard(cudaStream_t st);  // drop queued jobs (a backward pass failed half-way)

}  // namespace nlam
