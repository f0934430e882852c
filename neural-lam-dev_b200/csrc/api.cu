// C ABI entry points (include/nlam_b200.h) -> kernel launchers.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#include <map>
#include <mutex>
#include <utility>
#include <string.h>

#include "common.cuh"

namespace nlam {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::atomic<long long> g_launches{0};
static int env_or(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
static std::atomic<int> g_fwd_mc{env_or("NLAM_FWD_MC", -1)};
static std::atomic<int> g_dgrad_mc{env_or("NLAM_DGRAD_MC", 0)};
static std::atomic<int> g_bwd_fused{env_or("NLAM_BWD_FUSED", -1)};
static std::atomic<int> g_pdl{env_or("NLAM_PDL", 1)};
static std::atomic<int> g_tma{env_or("NLAM_TMA", 0)};
int option_tma() { return g_tma.load(); }
static std::atomic<int> g_wide128{env_or("NLAM_WIDE128", 1)};
int option_wide128() { return g_wide128.load(); }
static std::atomic<int> g_rb128{env_or("NLAM_RB128", 4)};
int option_rb128() { return g_rb128.load(); }
static std::atomic<int> g_bwd_nh{env_or("NLAM_BWD_NH", 2)};
int option_bwd_nh() { return g_bwd_nh.load(); }
static std::atomic<int> g_bwd_spread{env_or("NLAM_BWD_SPREAD", 2)};
int option_bwd_spread() { return g_bwd_spread.load(); }
static std::atomic<int> g_small512{env_or("NLAM_SMALL512", 1)};
int option_small512() { return g_small512.load(); }
static std::atomic<int> g_fp32_split{env_or("NLAM_FP32_SPLIT", 1)};
int option_fp32_split() { return g_fp32_split.load(); }
int option_fwd_mc() { return g_fwd_mc.load(); }
int option_dgrad_mc() { return g_dgrad_mc.load(); }
int option_bwd_fused() { return g_bwd_fused.load(); }
int option_pdl() { return g_pdl.load(); }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// largest opt-in so far per (kernel, device)
static std::mutex g_smem_mutex;
static std::map<std::pair<const void*, int>, int> g_smem_set;
cudaError_t ensure_dyn_smem(const void* kern, int bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(g_smem_mutex);
  int& have = g_smem_set[{kern, dev}];
  if (bytes > have) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    have = bytes;
  }
  return cudaSuccess;
}
}  // namespace nlam

using namespace nlam;

extern "C" int nlam_set_option(const char* name, int value) {
  if (name && !strcmp(name, "fwd_mc")) return nlam::g_fwd_mc.store(value), 0;
  if (name && !strcmp(name, "dgrad_mc")) return nlam::g_dgrad_mc.store(value), 0;
  if (name && !strcmp(name, "bwd_fused")) return nlam::g_bwd_fused.store(value), 0;
  if (name && !strcmp(name, "pdl")) return nlam::g_pdl.store(value), 0;
  if (name && !strcmp(name, "tma")) return nlam::g_tma.store(value), 0;
  if (name && !strcmp(name, "bwd_nh")) return nlam::g_bwd_nh.store(value), 0;
  if (name && !strcmp(name, "wide128")) return nlam::g_wide128.store(value), 0;
  if (name && !strcmp(name, "fp32_split")) return nlam::g_fp32_split.store(value), 0;
  if (name && !strcmp(name, "small512")) return nlam::g_small512.store(value), 0;
  if (name && !strcmp(name, "bwd_spread")) return nlam::g_bwd_spread.store(value), 0;
  nlam::set_error("nlam_set_option: unknown option");
  return 1;
}

extern "C" int64_t nlam_launch_count(void) { return (int64_t)nlam::g_launches.load(); }

extern "C" const char* nlam_last_error(void) { return g_err; }
extern "C" int nlam_version(void) { return 1; }

extern "C" int nlam_rowmlp_fwd(const nlam_rowmlp* d, void* stream) {
  NLAM_CHECK(d, "rowmlp_fwd: NULL descriptor");
  if (d->precision == NLAM_FP32) {
    if (tc::tc_split_supported(*d)) return tc_rowmlp_fwd(*d, (cudaStream_t)stream);
    return simt_rowmlp_fwd(*d, (cudaStream_t)stream);
  }
  if (d->precision == NLAM_BF16) {
    if (tc::tc_supported(*d)) return tc_rowmlp_fwd(*d, (cudaStream_t)stream);
    return simt_rowmlp_fwd(*d, (cudaStream_t)stream);  // widths the MMA path does not take
  }
  set_error("rowmlp_fwd: unknown precision mode %d", d->precision);
  return 1;
}

extern "C" int nlam_rowmlp_path(const nlam_rowmlp* d) {
  if (!d) return 0;
  if (d->precision == NLAM_BF16) return tc::tc_supported(*d) ? 1 : 0;
  if (d->precision == NLAM_FP32) return tc::tc_split_supported(*d) ? 2 : 0;
  return 0;
}

extern "C" size_t nlam_rowmlp_bwd_workspace(const nlam_rowmlp* d) {
  if (!d) return 0;
  if (d->precision == NLAM_BF16 && tc::tc_supported(*d)) return tc_rowmlp_bwd_workspace(*d);
  if (d->precision == NLAM_FP32 && tc::tc_split_supported(*d)) return tc_rowmlp_bwd_workspace(*d);
  return simt_rowmlp_bwd_workspace(*d);
}

extern "C" int nlam_rowmlp_bwd_stages(const nlam_rowmlp_bwd* d) {
  if (!d) return 0;
  if (d->fwd.precision == NLAM_BF16 && tc::tc_supported(d->fwd) && tc_rowmlp_bwd_is_fused(*d))
    return 2;
  return 3;
}

extern "C" size_t nlam_rowmlp_param_floats(const nlam_rowmlp* d) {
  if (!d) return 0;
  ParamLayout lay{k_total_of(*d), d->d_hidden, d->d_out, d->w.ln_g != nullptr};
  return (size_t)lay.total();
}

extern "C" int nlam_rowmlp_bwd_run(const nlam_rowmlp_bwd* d, void* stream) {
  NLAM_CHECK(d, "rowmlp_bwd: NULL descriptor");
  if (d->fwd.precision == NLAM_BF16 && tc::tc_supported(d->fwd))
    return tc_rowmlp_bwd(*d, (cudaStream_t)stream);
  if (d->fwd.precision == NLAM_FP32 && tc::tc_split_supported(d->fwd))
    return tc_rowmlp_bwd(*d, (cudaStream_t)stream);
  if (d->fwd.precision == NLAM_FP32 || d->fwd.precision == NLAM_BF16)
    return simt_rowmlp_bwd(*d, (cudaStream_t)stream);
  set_error("rowmlp_bwd: unknown precision mode %d", d->fwd.precision);
  return 1;
}

extern "C" int nlam_rowmlp_bwd_flush(void* stream) {
  return reduce_params_flush((cudaStream_t)stream);
}
extern "C" int nlam_rowmlp_bwd_pending(void) { return reduce_params_pending(); }
extern "C" int nlam_rowmlp_bwd_discard(void* stream) {
  return reduce_params_discard((cudaStream_t)stream);
}
