/*
 * nlam_b200.h -- C ABI of libnlam_b200.so: the B200 (sm_100a) kernels behind
 * Neural-LAM's InteractionNet message-passing hot path.
 *
 * The reference has NO native interface for this path (it is pure Python on
 * top of PyG/ATen); every entry point below replaces a piece
 * of Python in /root/reference and cites it.  All pointers are DEVICE
 * pointers unless said otherwise; sizes are element counts; `stream` is a
 * cudaStream_t passed as void*.  No compute function allocates device memory or
 * synchronises (re-entrant; CUDA-graph capturable); workspaces are caller-provided.
 * Process-wide state is limited to three things, all mutex/atomic protected: the
 * tuning switches of nlam_set_option, a per-(kernel, device) cache of shared-memory
 * opt-ins, and the per-(device, stream) queues of DEFERRED parameter-gradient
 * reductions (stage_mask bit 8; see nlam_rowmlp_bwd_flush).  Every function returns 0
 * on success; otherwise nlam_last_error() (thread-local) describes the failure.
 *
 * Row-MLP: out[b,r,:] = [residual +] LN( W2 * SiLU( W1 * concat_s src_s[b, idx_s[r], :] + b1 ) + b2 )
 * which covers, with different gather descriptors,
 *   - InteractionNet.message  (interaction_net.py:117-121; PyG gather of x_j/x_i)
 *   - the aggregation MLP     (interaction_net.py:106-109)
 *   - utils.make_mlp modules  (utils.py:191-214): embedders, grid MLP, output map
 *   - SplitMLPs               (interaction_net.py:134-163): per-tile weight set.
 */
#ifndef NLAM_B200_H
#define NLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NLAM_MAX_SRC 3
#define NLAM_TILE_ROWS 64 /* rows of a row-MLP tile (chunk tables use it) */

/* precision modes */
#define NLAM_FP32 0 /* fp32 parity mode (rtol 1e-4): tcgen05 with split bf16 operands where the
                       tiles fit (nlam_rowmlp_path == 2), fp32 FFMA kernels elsewhere */
#define NLAM_BF16 1 /* bf16 tcgen05 MMA, fp32 accumulate/storage (2e-2)  */

/* One gathered input of a row-MLP: row r of batch b is
 *   ptr + b*batch_stride + (idx ? idx[r] : r)*ld,  `width` floats. */
typedef struct nlam_src {
  const float* ptr;
  const int32_t* idx;   /* [rows] or NULL */
  int64_t batch_stride; /* 0 = shared by all batch items (stride-0 expand,
                           ar_model.py:204-209) */
  int32_t ld;
  int32_t width;
  /* Optional bf16 SHADOW of the same rows (NLAM_BF16 mode): dense rows of `width` bf16
   * values, row i of batch b at shadow + (b*shadow_batch_stride + i*width) elements,
   * written by the kernel that produced `ptr` (out_bf16 & co. below).  The tensor-core
   * kernels read their MMA operands from it (half the gather bytes, no conversion; the
   * TMA row-gather kernels need it); `ptr` stays the fp32 master for residuals. */
  const void* shadow;
  int64_t shadow_batch_stride; /* elements; 0 = shared */
  int64_t shadow_rows;         /* rows per batch item of the shadow tensor */
} nlam_src;

/* Weights of make_mlp([K, d_hidden, d_out]) (+LayerNorm), nn.Linear layout
 * (out, in).  With n_chunks > 1 (SplitMLPs) each pointer addresses n_chunks
 * stacked, contiguous copies and tile t uses set tile_chunk[t]. */
typedef struct nlam_mlp_weights {
  const float* w1; /* [d_hidden, K] */
  const float* b1; /* [d_hidden]    */
  const float* w2; /* [d_out, d_hidden] */
  const float* b2; /* [d_out]       */
  const float* ln_g; /* [d_out] or NULL (no LayerNorm) */
  const float* ln_b;
} nlam_mlp_weights;

/* Optional in-kernel deterministic segment reduction over the rows of a tile
 * (receiver-aligned tiles: every segment lies inside one tile).  Forward:
 * out[b, i, :] = scale[i] * sum_{r in [seg_ptr[i], seg_ptr[i+1])} y[b, r, :]
 * = PyG scatter sum/mean (interaction_net.py:124-131) fused into the edge
 * kernel, rows summed in tile-table order (= ascending edge id).  Needs the
 * bf16 path with d_hidden == d_out == source widths in {64, 128}. */
typedef struct nlam_agg {
  const int32_t* seg_ptr;  /* [n_seg+1] */
  const int32_t* tile_seg; /* [n_tiles+1]: tile t owns segments [tile_seg[t], tile_seg[t+1]) */
  const float* scale;      /* [n_seg] or NULL */
  float* out;              /* [batch, n_seg, d_out]; NULL = no reduction */
  int32_t n_seg;
  void* out_bf16;          /* optional bf16 shadow of `out` (may be given WITHOUT out: the
                              aggregated messages only feed the node MLP's MMA) */
} nlam_agg;

typedef struct nlam_rowmlp {
  int32_t n_src;
  nlam_src src[NLAM_MAX_SRC];
  int32_t batch;
  int32_t rows; /* rows per batch item */
  int32_t d_hidden;
  int32_t d_out;
  nlam_mlp_weights w;
  int32_t n_chunks;           /* 1 = single weight set */
  const int32_t* tile_ptr;    /* [n_tiles+1] row ranges (<= NLAM_TILE_ROWS rows
                                 each, never straddling a chunk) or NULL =
                                 uniform tiles */
  const int32_t* tile_chunk;  /* [n_tiles] or NULL */
  const int32_t* chunk_ptr;   /* [n_chunks+1] row offsets of the chunks or NULL */
  int32_t n_tiles;
  int32_t residual_src;       /* -1, or s: out += src_s row (needs width==d_out) */
  float* out;                 /* [batch, rows, d_out] contiguous */
  float* out_res;             /* optional second output: src_0 row + (the value
                                 written to out); InteractionNet needs both the
                                 message m_k and e_k + m_k (interaction_net.py:
                                 112,131).  NULL = not wanted */
  const int32_t* out_idx;     /* optional: row r of out / out_res is written to row
                                 out_idx[r] (scatter back to original edge order) */
  nlam_agg agg;               /* optional fused segment reduction (out may then be NULL) */
  int32_t precision;          /* NLAM_FP32 | NLAM_BF16 */
  void* out_bf16;             /* optional bf16 shadows of out / out_res (same row mapping), */
  void* out_res_bf16;         /* NLAM_BF16 mode, d_out == 64 */
} nlam_rowmlp;

/* Backward of a row-MLP (recomputes the forward from the same inputs).
 *   dOut[b,r,:] = (g0 ? g0[b,r,:] : 0) + (g1 ? g1_scale[g1_idx[r]] * g1[b, g1_idx[r], :] : 0)
 * d_src[s] (may be NULL) receives the un-reduced per-row gradient of source s,
 * dense [batch, rows, width_s]; for s == residual_src dOut is added.
 * d_params is one flat buffer per chunk: [dW1 | db1 | dW2 | db2 | dLNg | dLNb]
 * (LN parts absent without LayerNorm), overwritten unless params_accumulate. */
typedef struct nlam_rowmlp_bwd {
  nlam_rowmlp fwd;
  const float* g0;       /* [batch, rows, d_out] or NULL */
  const float* g1;       /* [batch, n1, d_out] or NULL  */
  const int32_t* g1_idx; /* [rows] */
  const float* g1_scale; /* [n1] or NULL (mean aggregation: 1/max(deg,1)) */
  int64_t g1_batch_stride;
  float* d_src[NLAM_MAX_SRC];
  const int32_t* g0_idx;                   /* optional: dOut row r = g0[b, g0_idx[r], :] */
  const int32_t* d_src_idx[NLAM_MAX_SRC];  /* optional: row r of d_src[s] goes to row idx[r] */
  int32_t reduce_src;        /* -1, or s: the gradient rows of source s are segment-summed
                                (tables of fwd.agg) into d_src[s] = [batch, n_seg, width] */
  int32_t reduce_accumulate; /* add to the existing contents of d_src[reduce_src] */
  float* d_params;
  int32_t params_accumulate; /* 0: d_params is overwritten; 1: gradients are ADDED to it
                                (lets the caller pass its flat gradient buffer) */
  float* workspace;       /* nlam_rowmlp_bwd_workspace() floats */
  size_t workspace_floats;
  int32_t stage_mask;     /* 0 = everything; else bit 0: input-gradient kernel, bit 1:
                             weight-gradient kernel, bit 2: partial reduction (lets a
                             profiler time the three launches separately; the stages
                             must run in this order on the same workspace); bit 3
                             (value 8, alone = everything): queue the partial reduction
                             for nlam_rowmlp_bwd_flush instead of launching it; bit 4
                             (value 16): the rows of fwd.src were NOT written by the
                             kernel launched just before this call on the stream (true
                             under autograd: they were saved by the forward pass) -- the
                             fused kernel then gathers its first tile before it waits
                             for that kernel (programmatic dependent launch) */
  int32_t g0_sum_count;   /* > 1: dOut row r = sum over k < g0_sum_count of the g0 rows at    */
  int64_t g0_sum_stride;  /* g0 + k * g0_sum_stride floats (fwd.batch must be 1): the backward */
                          /* of an output that was expand()-ed over a batch, without the       */
                          /* summed copy; only where nlam_rowmlp_bwd_stages() == 2             */
  /* bf16 shadows of the upstream gradients (dense [batch, rows|n1, 64]; written by the
   * backward kernel that produced g0 / g1 through d_src_bf16) and of the per-row source
   * gradients this run writes.  The TMA kernel (nlam_rowmlp_bwd_stages() == 1) needs the
   * shadows of every gradient it reads. */
  const void* g0_bf16;
  const void* g1_bf16;
  void* d_src_bf16[NLAM_MAX_SRC];
  /* Sender pre-reduction (TMA kernel): instead of one gradient row per edge, source
   * `sp_src` gets one PARTIAL row per (tile, distinct sender in the tile): partial q =
   * sum of the tile rows sp_rows[sp_row_ptr[q] .. sp_row_ptr[q+1]) (tile-local row ids,
   * ascending), tile t owns partials [sp_tile_ptr[t], sp_tile_ptr[t+1]); d_src[sp_src] is
   * then [batch, n_sp, width].  A CSR over the partials (by sender) finishes the sum
   * (nlam_segsum).  sp_src = -1: off. */
  int32_t sp_src;
  int32_t n_sp;
  const int32_t* sp_tile_ptr;
  const int32_t* sp_row_ptr;
  const int32_t* sp_rows;
  /* src0_batch_sum = 1: source 0 is shared by the batch (batch_stride 0) and d_src[0] is
   * [1, rows, width] = the gradient SUMMED over the batch on chip (TMEM accumulation over
   * the batch tiles of one row range) instead of [batch, rows, width]. */
  int32_t src0_batch_sum;
} nlam_rowmlp_bwd;

/* out[b,i,:] (+)= scale[i] * sum_{p in [ptr[i],ptr[i+1])} src[b, idx[p], :]
 * Deterministic segment sum in list order: PyG scatter-sum/mean
 * (interaction_net.py:124-131) with a receiver-sorted CSR, and the transposed
 * (sender-sorted) CSR for the backward of the x_j gather. */
typedef struct nlam_segsum {
  const float* src;
  int64_t src_batch_stride;
  const int32_t* ptr; /* [n_out+1] */
  const int32_t* idx; /* [ptr[n_out]] */
  const float* scale; /* [n_out] or NULL */
  float* out;         /* [batch, n_out, width] contiguous */
  int32_t batch, n_out, width;
  int32_t accumulate; /* 0: overwrite, 1: add to out */
} nlam_segsum;

/* One auto-regressive state update fused with its loss term:
 *   pred = prev + net_out*diff_std + diff_mean            (base_graph_model.py:174-177)
 *   new  = interior ? pred : truth                        (ar_model.py:244-247)
 *   loss_sum = sum_{rows,f} interior * ((new - truth) * inv_std[f])^2
 * (metrics.py:21-84 wmse/mse with the interior mask; the caller divides by
 * batch * steps * #interior, ar_model.py:294-298).  Tensors are [rows, features]
 * contiguous with rows = batch*nodes; inv_std NULL = unweighted (mse). */
typedef struct nlam_state_step {
  const float* net_out;
  const float* prev;
  const float* truth;
  const float* diff_std;  /* [features] */
  const float* diff_mean; /* [features] */
  const float* inv_std;   /* [features] = 1/per_var_std, or NULL */
  const float* interior;  /* [nodes] 1 = interior, 0 = boundary */
  float* new_state;       /* [rows, features] */
  float* loss_partial;    /* [nlam_state_step_partials(rows)] scratch */
  float* loss_sum;        /* [1] */
  int64_t rows;
  int32_t nodes;
  int32_t features;
  int64_t prev_batch_stride;  /* floats between batch items of prev / truth (rows of a batch */
  int64_t truth_batch_stride; /* item stay dense); 0 = nodes * features.  Lets the caller  */
                              /* pass slices like init_states[:, 1] without a copy        */
} nlam_state_step;

/* Backward: d_pred = interior * (d_new + d_loss * 2 (new - truth) inv_std^2);
 * d_net_out = d_pred * diff_std; d_prev = d_pred.  d_new / d_loss (device scalar)
 * may be NULL; fwd.new_state is the saved forward output. */
typedef struct nlam_state_step_bwd {
  nlam_state_step fwd;
  const float* d_new;
  const float* d_loss;
  float* d_net_out;
  float* d_prev;
} nlam_state_step_bwd;

/* Device-side data feed (replaces the per-sample xarray work of
 * neural_lam/weather_dataset.py:163-496, analysis data).  The standardised time series stays
 * in HBM: state [slots, n_grid, d_state], forcing [slots, n_grid, d_forcing] (or NULL), times
 * [slots]; logical time step t lives in slot (ring_cap ? t % ring_cap : t).  Sample i uses
 *   states   t = i + max(0, past - 2) + {0, 1}            -> init_states   [B, 2, N, d_state]
 *            t = i + max(0, past - 2) + 2 + s             -> target_states [B, ar, N, d_state]
 *   forcing  t = i + max(2, past) + s - past + w, w < past + future + 1
 *            -> forcing_out[b, s, n, f * (past+future+1) + w]  (feature-major, window inner)
 *   target_times[b, s] = times of the target steps.
 * sample_idx is passed by value (host side validates every window against [t_lo, t_hi)). */
#define NLAM_FEED_MAX_BATCH 64
typedef struct nlam_feed_batch {
  const float* state;
  const float* forcing;
  const int64_t* times;
  int64_t sample_idx[NLAM_FEED_MAX_BATCH];
  int32_t batch, n_grid, d_state, d_forcing;
  int32_t ar_steps, past, future;
  int32_t ring_cap;     /* 0 = the series is stored linearly from time step 0 */
  int64_t t_lo, t_hi;   /* resident logical time steps [t_lo, t_hi) */
  float* init_states;
  float* target_states;
  float* forcing_out;   /* [B, ar, N, d_forcing * (past + future + 1)] */
  int64_t* target_times; /* [B, ar] or NULL */
} nlam_feed_batch;

const char* nlam_last_error(void);
int nlam_version(void);
/* Kernel-selection knobs (process-wide; -1 = automatic, 0 = off, 1 = force when
 * eligible): "fwd_mc" = 4-pipeline shared-weight forward kernel, "dgrad_mc" =
 * 3-pipeline shared-weight input-gradient kernel.  Returns 0, or 1 for an
 * unknown name.  "bwd_fused" (default on) = single input + weight gradient kernel
 * for square 64-wide MLPs.  Environment NLAM_FWD_MC / NLAM_DGRAD_MC /
 * NLAM_BWD_FUSED give the initial values.  "pdl" (default 1, env NLAM_PDL): launch with
 * programmatic dependent launch so that a kernel's prologue overlaps the tail of the
 * previous one (every kernel waits with griddepcontrol.wait before reading its inputs).
 * "tma" (default 0, env NLAM_TMA): forward edge kernel whose operands arrive by
 * cp.async.bulk.tensor tile::gather4 from the bf16 shadows; "bwd_nh" (default 2, env
 * NLAM_BWD_NH): threads per tile row of the fused backward kernel (2 or 4).  "wide128"
 * (default 1, env NLAM_WIDE128): 512-thread CTAs for the d = 128 kernels.  "fp32_split"
 * (default 1, env NLAM_FP32_SPLIT): precision NLAM_FP32 on the tcgen05 kernels with split
 * bf16 operands where the tiles fit (see nlam_rowmlp_path); 0 = fp32 FFMA kernels.
 * "small512" (default 1, env NLAM_SMALL512): 512-thread forward CTAs for d = 64 launches
 * of at most 148 tiles (the small levels of the hierarchical meshes); "bwd_spread" (default
 * 2, env NLAM_BWD_SPREAD): fused backward launches of at most 148 tiles run one tile per CTA
 * instead of two (1), on the single-context variant of the kernel with four threads per
 * tile row (2); 0 = as every other launch. */
int nlam_set_option(const char* name, int value);
/* Number of kernels this library has launched in this process (monotonic;
 * bench.py reports the delta over its timed region as "gpu_launches"). */
int64_t nlam_launch_count(void);

/* Stable counting sort of the M edges by `key` (receiver or sender id):
 * ptr[n_keys+1], perm[M] = edge ids grouped by key, ascending within a key
 * (== numpy argsort(kind="stable")); inv_deg[n_keys] = 1/max(deg,1) (may be
 * NULL).  workspace: n_keys+1 int32.  Replaces the scatter index handling of
 * PyG for the edge_index produced by interaction_net.py:55-61. */
int nlam_csr_build(const int32_t* key, int64_t n_edges, int32_t n_keys, int32_t* ptr,
                   int32_t* perm, float* inv_deg, int32_t* workspace, void* stream);

int nlam_rowmlp_fwd(const nlam_rowmlp* desc, void* stream);
/* Kernel family nlam_rowmlp_fwd / nlam_rowmlp_bwd_run take for this descriptor: 0 = fp32
 * FFMA kernels; 1 = bf16 tcgen05 kernels (precision NLAM_BF16); 2 = precision NLAM_FP32 on
 * the tcgen05 kernels -- every fp32 operand is held as two bf16 tiles (hi = bf16(x), lo =
 * bf16(x - hi)) and every product is three UMMAs into one fp32 accumulator, ~16 mantissa
 * bits; taken when the doubled tiles fit shared memory (all d_hidden <= 64 shapes of the
 * models) and option "fp32_split" (default 1, env NLAM_FP32_SPLIT) is on.  Families 1 and 2
 * support fused aggregation / row scatter (agg, out_idx, reduce_src ...), family 0 does not.
 * Only widths, alignment and precision of the descriptor are inspected. */
int nlam_rowmlp_path(const nlam_rowmlp* desc);
size_t nlam_rowmlp_bwd_workspace(const nlam_rowmlp* desc);
size_t nlam_rowmlp_param_floats(const nlam_rowmlp* desc); /* per chunk */
/* Kernels nlam_rowmlp_bwd_run launches for this descriptor: 3 = input gradients,
 * weight gradients, partial reduction (stage_mask bits 1, 2, 4); 2 = one fused
 * input + weight gradient kernel (bit 1; bit 2 is a no-op) and the reduction; 1 = the
 * TMA row-gather variant of the fused kernel (operands and upstream gradients read
 * from bf16 shadows by cp.async.bulk.tensor tile::gather4) and the reduction.  The
 * choice can depend on which source gradients (d_src) the descriptor asks for. */
int nlam_rowmlp_bwd_stages(const nlam_rowmlp_bwd* desc);
int nlam_rowmlp_bwd_run(const nlam_rowmlp_bwd* desc, void* stream);
/* Deferred parameter-gradient reductions.  A run with stage_mask bit 3 (value 8) set
 * does not launch its own partial reduction but queues it on the queue of (current
 * device, `stream`) (thread safe); nlam_rowmlp_bwd_flush runs every reduction queued
 * for (current device, `stream`) in ONE launch on that stream.  The caller keeps the
 * runs' workspaces and d_params alive until the flush.  nlam_rowmlp_bwd_pending =
 * queued reductions over all queues; nlam_rowmlp_bwd_discard drops the queue of
 * (current device, `stream`) without running it (a backward pass that failed half-way
 * must not leave pointers to freed workspaces behind). */
int nlam_rowmlp_bwd_flush(void* stream);
int nlam_rowmlp_bwd_pending(void);
int nlam_rowmlp_bwd_discard(void* stream);
int nlam_segsum_run(const nlam_segsum* desc, void* stream);
/* dst[r, c] = (src[r, c] - mean[c]) / std[c]  (weather_dataset.py:399-420), in place allowed */
int nlam_feed_standardize(const float* src, const float* mean, const float* std, float* dst,
                          int64_t rows, int32_t d, void* stream);
int nlam_feed_batch_run(const nlam_feed_batch* desc, void* stream);
int64_t nlam_state_step_partials(int64_t rows);
int nlam_state_step_fwd(const nlam_state_step* desc, void* stream);
int nlam_state_step_bwd_run(const nlam_state_step_bwd* desc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NLAM_B200_H */
