"""Host side of the custom ops: builds C-ABI descriptors from tensors and wires
them into autograd.  PyTorch is only the tensor / autograd carrier; every
FLOP of the hot path runs in csrc/ kernels.

Ops
---
mlp_forward(module, x)         utils.make_mlp modules (utils.py:191-214)
split_mlp_forward(module, x)   SplitMLPs (interaction_net.py:134-163)
interaction_net(...)           InteractionNet.forward (interaction_net.py:86-131)
"""
import ctypes

import numpy as np
import torch
from torch import nn

from . import lib as L

import os as _os

_state = {"precision": "fp32",
          "disable_aligned": _os.environ.get("NLAM_FUSED_AGG", "1") == "0",
          # on-chip reductions of the fused edge backward (set_backward_reductions):
          # sender partials: measured -3 % on the m2g layer (1274 -> 1235 us fwd+bwd), on;
          # batch-summed static edge gradient: measured +29 % (1274 -> 1646 us: the batch
          # items of a row tile must then run one after the other in ONE context and
          # read-modify-write their gradient rows), off
          "no_sender_partials": _os.environ.get("NLAM_SENDER_PARTIALS", "1") == "0",
          "no_batch_sum": _os.environ.get("NLAM_BATCH_SUM", "0") != "1"}
_PREC = {"fp32": L.FP32, "bf16": L.BF16}


def set_precision(mode):
    """'fp32' (parity mode: tcgen05 with split bf16 operands where the tiles fit, FFMA kernels
    elsewhere -- see kernel_family) or 'bf16' (tcgen05 MMA, fp32 accumulate)."""
    assert mode in _PREC
    _state["precision"] = mode


def get_precision():
    return _state["precision"]


def set_param_grad_sink(enabled):
    """When the six gradient tensors (.grad) of an MLP already exist and are
    adjacent views of one flat buffer (train.FlatGradBuckets), let the kernels
    ADD the weight gradients straight into that buffer instead of returning them
    to autograd (which would launch one accumulation kernel per parameter)."""
    _state["grad_sink"] = bool(enabled)


def set_deferred_param_reduce(enabled):
    """Queue the per-MLP partial reductions of the parameter gradients and run them in
    one launch when `flush_param_grads()` is called (the trainer does, right after
    `loss.backward()`).  Off by default: with plain `loss.backward()` +
    `optimizer.step()` nobody would call the flush."""
    if not enabled:
        flush_param_grads()
    _state["defer_reduce"] = bool(enabled)


_deferred_keep = []  # workspaces / gradient buffers of queued reductions


def scoped_trainer_flags(sink, defer):
    """Set (grad_sink, defer_reduce) and return the previous pair (train.py scopes both to
    its own forward/backward).  Does not flush: the caller does."""
    prev = (_state.get("grad_sink", False), _state.get("defer_reduce", False))
    _state["grad_sink"], _state["defer_reduce"] = bool(sink), bool(defer)
    return prev


def discard_param_grads():
    """Drop the queued reductions of the current stream without running them (error path)."""
    if torch.cuda.is_available():
        L.load().nlam_rowmlp_bwd_discard(_stream())
    _deferred_keep.clear()


def flush_param_grads():
    """Run every queued parameter-gradient reduction (one kernel) on the current stream."""
    lib = L.load()
    if lib.nlam_rowmlp_bwd_pending() > 0:
        L.check(lib.nlam_rowmlp_bwd_flush(_stream()), "nlam_rowmlp_bwd_flush")
    _deferred_keep.clear()


def _grad_sink(params):
    """Flat destination tensor for an MLP's parameter gradients, or None."""
    if not _state.get("grad_sink", False):
        return None
    ps = [q for q in params if q is not None]
    if any((not isinstance(q, nn.Parameter)) or q.grad is None or not q.grad.is_contiguous()
           or q.grad.dtype != torch.float32 for q in ps):
        return None
    base = ps[0].grad
    off = base.data_ptr()
    total = 0
    for q in ps:
        if q.grad.data_ptr() != off:
            return None
        off += 4 * q.numel()
        total += q.numel()
    if base.storage_offset() + total > base.untyped_storage().nbytes() // 4:
        return None
    return torch.as_strided(base, (1, total), (total, 1))


def set_backward_reductions(sender_partials=True, batch_sum=False):
    """On-chip reductions of the fused edge backward (A/B switch for tests and benchmarks):
    per-(tile, sender) partial gradient rows instead of per-edge rows, and the gradient of a
    batch-shared edge embedding summed over the batch inside the kernel."""
    _state["no_sender_partials"] = not sender_partials
    _state["no_batch_sum"] = not batch_sum


def set_fused_aggregation(enabled):
    """Receiver-aligned tiles with the segment sum inside the edge kernel (bf16
    path, d in {64,128}) vs. messages written once and summed by nlam_segsum."""
    _state["disable_aligned"] = not enabled


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class KernelTimer:
    """Optional per-launch CUDA-event timing (on the launching stream) used by
    bench.py for the roofline of the dominant kernel.  `only`: restrict to one
    tag so the timed region is not perturbed by the others."""

    def __init__(self, only=None):
        self.only = only
        self.records = {}  # tag -> {"events": [(start, end)], "bytes": int, "flops": int}

    def want(self, tag):
        return self.only is None or tag == self.only

    def start(self, tag, nbytes, flops, impl_bytes=None):
        rec = self.records.setdefault(tag, {"events": [], "bytes": nbytes, "flops": flops,
                                            "impl_bytes": impl_bytes if impl_bytes is not None
                                            else nbytes})
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
        rec["events"].append(ev)
        return ev[1]

    def summary(self):
        """tag -> (launches, total_ms, bytes_per_launch, flops_per_launch); call after a sync."""
        out = {}
        for tag, rec in self.records.items():
            ms = sum(a.elapsed_time(b) for a, b in rec["events"])
            out[tag] = (len(rec["events"]), ms, rec["bytes"], rec["flops"])
        return out

    def impl_bytes(self, tag):
        return self.records[tag]["impl_bytes"]


_timer = {"t": None}


def set_timer(timer):
    _timer["t"] = timer


def _src_bytes(t, idx, batch):
    """Algorithmic bytes of one source: every distinct row once."""
    nb = 1 if (t.shape[0] == 1 or t.stride(0) == 0) else batch
    return 4 * nb * t.shape[1] * t.shape[2] + (4 * idx.numel() if idx is not None else 0)


def _rowmlp_cost(kind, srcs, W, batch, rows, extra_row_floats):
    """(tag, algorithmic bytes, flops) of one row-MLP launch.  bytes: distinct
    input rows + gather indices + weights + rows written/read besides the
    inputs; flops: 2*rows*(K*dh + dh*dout) forward, 3x for dgrad+recompute."""
    k = sum(it[0].shape[2] for it in srcs)
    tag = f"{kind}|rows={rows}|K={k}|dh={W.d_hidden}|dout={W.d_out}|B={batch}"
    nbytes = sum(_src_bytes(it[0], it[1], batch) for it in srcs)
    nbytes += 4 * W.n_chunks * W.param_floats()
    nbytes += 4 * batch * rows * extra_row_floats
    flops = 2 * batch * rows * (k * W.d_hidden + W.d_hidden * W.d_out)
    return tag, nbytes, flops


def _check_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(
            f"{what}: tensor is on {t.device}; the neural_lam_b200 kernels run on CUDA "
            "only (there is no CPU fallback)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{what}: expected float32, got {t.dtype}")


def _rows3d(t, what):
    """Normalise to a 3-D (B, n, w) tensor whose rows are contiguous.  A
    stride-0 batch dim (expand_to_batch) is kept as is."""
    _check_cuda(t, what)
    if t.dim() != 3:
        raise ValueError(f"{what}: expected (B, N, d), got {tuple(t.shape)}")
    w = t.shape[2]
    ok = (t.stride(2) == 1 or w == 1) and t.stride(1) >= w and t.stride(0) >= 0
    if t.shape[1] == 1 and not ok:
        ok = t.stride(2) == 1
    if not ok:
        t = t.contiguous()
    return t


def _src(t, idx, sh=None):
    s = L.Src()
    s.ptr = t.data_ptr()
    s.idx = idx.data_ptr() if idx is not None else None
    s.batch_stride = t.stride(0) if t.shape[0] > 1 else 0
    s.ld = t.stride(1) if t.shape[1] > 1 else t.shape[2]
    s.width = t.shape[2]
    if sh is not None:
        assert sh.dtype == torch.bfloat16 and sh.shape == t.shape and sh.stride(2) == 1 \
            and (sh.shape[1] == 1 or sh.stride(1) == sh.shape[2])
        s.shadow = sh.data_ptr()
        s.shadow_batch_stride = sh.stride(0) if (sh.shape[0] > 1 and s.batch_stride != 0) else 0
        s.shadow_rows = sh.shape[1]
    return s


# ---- bf16 shadows -----------------------------------------------------------------------
# In bf16 mode every kernel that produces a 64-wide representation also writes a bf16 copy
# of it (the "shadow"); the consuming kernels build their MMA operands from the shadow (half
# the gather bytes, no fp32 -> bf16 conversion; the TMA row-gather kernels need it), while
# the fp32 tensor stays the master for residual streams and for autograd.  The shadow
# travels as an attribute of the fp32 tensor object, together with the tensor's version
# counter at the time it was written (an in-place update of the master invalidates it).
SHADOW_WIDTHS = (64, 128)


def attach_shadow(t, sh):
    if sh is not None:
        t._nlam_sh = (sh, t._version)
    return t


def shadow_of(t):
    rec = getattr(t, "_nlam_sh", None)
    if rec is None or _state.get("no_shadows", False):
        return None
    sh, ver = rec
    if t._version != ver or sh.shape != t.shape or sh.device != t.device:
        return None
    return sh


def make_shadow(t):
    """Give a tensor that did not come out of one of this library's kernels (a leaf, user
    data) a bf16 shadow, so that consumers can take the shadow / TMA paths."""
    if t.is_cuda and t.dtype == torch.float32 and t.shape[-1] in SHADOW_WIDTHS:
        attach_shadow(t, t.detach().to(torch.bfloat16).contiguous()
                      if t.is_contiguous() else t.detach().to(torch.bfloat16))
    return t


def set_shadows(enabled):
    """bf16 shadow activations on / off (bf16 mode only; off = every kernel gathers fp32)."""
    _state["no_shadows"] = not enabled


def _want_shadow(W, precision, batch_rows_ok=True):
    return (precision == "bf16" and not _state.get("no_shadows", False)
            and W.d_out in SHADOW_WIDTHS and W.d_hidden <= 128 and W.k <= 384 and batch_rows_ok)


def expand_with_shadow(x, batch_size):
    """x (N, d) -> stride-0 (B, N, d) view; carries the shadow along."""
    out = x.unsqueeze(0).expand(batch_size, -1, -1)
    sh = shadow_of(x)
    if sh is not None:
        attach_shadow(out, sh.unsqueeze(0).expand(batch_size, -1, -1))
    return out


class TileTable:
    """Row tiles that never straddle a chunk (SplitMLPs weight sets)."""

    def __init__(self, chunk_sizes, device):
        tp, tc = [0], []
        cp = [0]
        for c, n in enumerate(chunk_sizes):
            start = cp[-1]
            for r0 in range(0, n, L.TILE_ROWS):
                tp.append(start + min(n, r0 + L.TILE_ROWS))
                tc.append(c)
            cp.append(start + n)
        self.n_tiles = len(tc)
        self.n_chunks = len(chunk_sizes)
        self.rows = cp[-1]
        mk = lambda a: torch.tensor(np.asarray(a, dtype=np.int32), device=device)
        self.tile_ptr, self.tile_chunk, self.chunk_ptr = mk(tp), mk(tc), mk(cp)


class Weights:
    """The six tensors of make_mlp([K, dh, dout]) (+LN); stacked on a leading
    chunk dim for SplitMLPs."""

    def __init__(self, w1, b1, w2, b2, ln_g, ln_b, n_chunks):
        self.t = (w1, b1, w2, b2, ln_g, ln_b)
        self.n_chunks = n_chunks
        self.d_hidden, self.k = w1.shape[-2], w1.shape[-1]
        self.d_out = w2.shape[-2]
        self.has_ln = ln_g is not None

    def fill(self, desc):
        for name, t in zip(("w1", "b1", "w2", "b2", "ln_g", "ln_b"), self.t):
            setattr(desc.w, name, t.data_ptr() if t is not None else None)

    def param_floats(self):
        n = self.d_hidden * self.k + self.d_hidden + self.d_out * self.d_hidden + self.d_out
        return n + (2 * self.d_out if self.has_ln else 0)

    def split_grads(self, flat):
        """flat (n_chunks, P) -> grads shaped like self.t (all None if the
        gradients were accumulated in place, see set_param_grad_sink)."""
        if flat is None:
            return [None] * 6
        C, dh, k, do = self.n_chunks, self.d_hidden, self.k, self.d_out
        sizes = [dh * k, dh, do * dh, do] + ([do, do] if self.has_ln else [])
        parts = torch.split(flat, sizes, dim=1)
        shapes = [(dh, k), (dh,), (do, dh), (do,), (do,), (do,)]
        out = []
        for i, t in enumerate(self.t):
            if t is None:
                out.append(None)
            else:
                g = parts[i].reshape((C,) + shapes[i])
                out.append(g if t.dim() == len(shapes[i]) + 1 else g[0])
        return out


def _mlp_layers(mlp):
    layers = list(mlp)
    ok = (len(layers) in (3, 4) and isinstance(layers[0], nn.Linear)
          and isinstance(layers[1], nn.SiLU) and isinstance(layers[2], nn.Linear)
          and (len(layers) == 3 or isinstance(layers[3], nn.LayerNorm)))
    if not ok:
        raise NotImplementedError(
            "fused MLP kernels cover make_mlp blueprints with one hidden layer "
            "(Linear, SiLU, Linear[, LayerNorm]); got " + repr(mlp))
    ln = layers[3] if len(layers) == 4 else None
    return layers[0], layers[2], ln


def _linears_of(mlp):
    """(list of nn.Linear, LayerNorm | None) of a make_mlp Sequential with any number of hidden
    layers: Linear, (SiLU, Linear)*, [LayerNorm]  (utils.py:191-214)."""
    layers = list(mlp)
    ln = layers.pop() if layers and isinstance(layers[-1], nn.LayerNorm) else None
    lin = layers[0::2]
    ok = (len(layers) % 2 == 1 and all(isinstance(m, nn.Linear) for m in lin)
          and all(isinstance(m, nn.SiLU) for m in layers[1::2]))
    if not ok or len(lin) < 2:
        raise NotImplementedError(
            "fused MLP kernels cover make_mlp blueprints with at least one hidden layer "
            "(Linear, (SiLU, Linear)+[, LayerNorm]); got " + repr(mlp))
    return lin, ln


_identity_cache = {}


def _identity(d, device):
    key = (d, str(device))
    if key not in _identity_cache:
        _identity_cache[key] = (torch.eye(d, device=device), torch.zeros(d, device=device))
    return _identity_cache[key]


def blocks_of(module):
    """Kernel-sized pieces of a make_mlp Sequential / SplitMLPs with h >= 1 hidden layers
    (`--hidden_layers`, train_model.py:94): every launch of the fused kernel computes
    [LN](W_b . SiLU(W_a . x + b_a) + b_b).  Block 0 = the first two Linears; every further
    Linear W_k rides on an identity first layer, [LN](W_k . SiLU(I . y + 0) + b_k) -- exactly
    the SiLU that precedes it in the blueprint.  LayerNorm belongs to the last block.  With
    one hidden layer (every BASELINE config) this is the single block of weights_of()."""
    mlps = getattr(module, "mlps", None)
    per_chunk = [_linears_of(m) for m in mlps] if mlps is not None else [_linears_of(module)]
    n_lin = len(per_chunk[0][0])
    if any(len(lin) != n_lin for lin, _ in per_chunk):
        raise NotImplementedError("SplitMLPs chunks with different depths")
    C = len(per_chunk)
    st = (lambda ts: torch.stack(list(ts))) if mlps is not None else (lambda ts: list(ts)[0])
    blocks = []
    for k in range(1, n_lin):
        last = k == n_lin - 1
        ln_g = st(ln.weight for _, ln in per_chunk) if (last and per_chunk[0][1] is not None) else None
        ln_b = st(ln.bias for _, ln in per_chunk) if (last and per_chunk[0][1] is not None) else None
        w2, b2 = st(lin[k].weight for lin, _ in per_chunk), st(lin[k].bias for lin, _ in per_chunk)
        if k == 1:
            w1, b1 = st(lin[0].weight for lin, _ in per_chunk), st(lin[0].bias for lin, _ in per_chunk)
        else:
            d = per_chunk[0][0][k].in_features
            eye, zero = _identity(d, w2.device)
            w1 = eye.expand(C, d, d).contiguous() if mlps is not None else eye
            b1 = zero.expand(C, d).contiguous() if mlps is not None else zero
        blocks.append(Weights(w1, b1, w2, b2, ln_g, ln_b, C if mlps is not None else 1))
    return blocks


def weights_of(module):
    """Weights of a make_mlp Sequential or a SplitMLPs."""
    mlps = getattr(module, "mlps", None)
    if mlps is None:
        l1, l2, ln = _mlp_layers(module)
        return Weights(l1.weight, l1.bias, l2.weight, l2.bias,
                       ln.weight if ln is not None else None,
                       ln.bias if ln is not None else None, 1)
    trip = [_mlp_layers(m) for m in mlps]
    st = lambda ts: torch.stack(list(ts))
    has_ln = trip[0][2] is not None
    return Weights(st(t[0].weight for t in trip), st(t[0].bias for t in trip),
                   st(t[1].weight for t in trip), st(t[1].bias for t in trip),
                   st(t[2].weight for t in trip) if has_ln else None,
                   st(t[2].bias for t in trip) if has_ln else None, len(trip))


def _fill_desc(desc, srcs, W, batch, rows, residual, tiles, out, precision, aligned=None):
    """aligned: optional dict with the receiver-aligned tile table and segment
    tables of a _GraphPlan (fused aggregation path)."""
    desc.n_src = len(srcs)
    for i, item in enumerate(srcs):
        desc.src[i] = _src(*item)
    desc.batch, desc.rows = batch, rows
    desc.d_hidden, desc.d_out = W.d_hidden, W.d_out
    W.fill(desc)
    desc.n_chunks = W.n_chunks
    if tiles is not None:
        assert tiles.rows == rows and tiles.n_chunks == W.n_chunks
        desc.tile_ptr = tiles.tile_ptr.data_ptr()
        desc.tile_chunk = tiles.tile_chunk.data_ptr()
        desc.chunk_ptr = tiles.chunk_ptr.data_ptr()
        desc.n_tiles = tiles.n_tiles
    elif aligned is not None:
        assert W.n_chunks == 1
        desc.tile_ptr = aligned["tile_ptr"].data_ptr()
        desc.n_tiles = aligned["n_tiles"]
        desc.agg.seg_ptr = aligned["seg_ptr"].data_ptr()
        desc.agg.tile_seg = aligned["tile_seg"].data_ptr()
        desc.agg.n_seg = aligned["n_seg"]
    else:
        assert W.n_chunks == 1
        desc.n_tiles = (rows + L.TILE_ROWS - 1) // L.TILE_ROWS
    desc.residual_src = 0 if residual else -1
    desc.out = out.data_ptr() if out is not None else None
    desc.precision = _PREC[precision]


def rowmlp_fwd_raw(srcs, W, batch, rows, residual, tiles, precision, want_res=False,
                   aligned=None, sh_box=None):
    """srcs: list of (3-D tensor, int32 row index or None).  want_res: also
    return src_0 + out (second output of the same launch).
    aligned (fused aggregation, see _GraphPlan): rows are processed in
    receiver-sorted order; returns (None, src_0 + out scattered back through
    aligned['out_idx'] if want_res else None, aggregated [batch, n_seg, d_out])."""
    lib = L.load()
    dev = srcs[0][0].device
    desc = L.RowMlp()
    out = out_res = agg_out = None
    if aligned is None:
        out = torch.empty((batch, rows, W.d_out), device=dev, dtype=torch.float32)
    _fill_desc(desc, srcs, W, batch, rows, residual, tiles, out, precision, aligned)
    if want_res:
        out_res = torch.empty((batch, rows, W.d_out), device=dev, dtype=torch.float32)
        desc.out_res = out_res.data_ptr()
    if aligned is not None:
        agg_out = torch.empty((batch, aligned["n_seg"], W.d_out), device=dev, dtype=torch.float32)
        desc.agg.out = agg_out.data_ptr()
        desc.agg.scale = aligned["scale"].data_ptr() if aligned.get("scale") is not None else None
        desc.out_idx = aligned["out_idx"].data_ptr()
    if sh_box is not None and _want_shadow(W, precision):
        # bf16 shadows of everything this launch produces (see attach_shadow)
        mk = lambda t: torch.empty(t.shape, device=dev, dtype=torch.bfloat16)
        if out is not None:
            sh_box["out"] = mk(out)
            desc.out_bf16 = sh_box["out"].data_ptr()
        if out_res is not None:
            sh_box["out_res"] = mk(out_res)
            desc.out_res_bf16 = sh_box["out_res"].data_ptr()
        if agg_out is not None:
            sh_box["agg"] = mk(agg_out)
            desc.agg.out_bf16 = sh_box["agg"].data_ptr()
    end = None
    if _timer["t"] is not None:
        extra = W.d_out * ((1 if out is not None else 0) + (1 if want_res else 0))
        kind = f"rowmlp_fwd_{precision}" + ("_agg" if aligned is not None else "")
        tag, nbytes, flops = _rowmlp_cost(kind, srcs, W, batch, rows, extra)
        if aligned is not None:
            nbytes += 4 * batch * aligned["n_seg"] * W.d_out + 4 * aligned["n_seg"]
        if _timer["t"].want(tag):
            end = _timer["t"].start(tag, nbytes, flops)
    L.check(lib.nlam_rowmlp_fwd(ctypes.byref(desc), _stream()), "nlam_rowmlp_fwd")
    if end is not None:
        end.record()
    if aligned is not None:
        return None, out_res, agg_out
    return (out, out_res) if want_res else out


def rowmlp_bwd_raw(srcs, W, batch, rows, residual, tiles, precision, g0, need_src,
                   g1=None, g1_idx=None, g1_scale=None, aligned=None, g0_idx=None,
                   d_src_idx=None, reduce_src=-1, reduce_into=None, sink_params=None,
                   g0_sum=False, algo_dsrc_bytes=None, sp=None, batch_sum0=0, info=None):
    """Returns (list of per-row source grads or None, d_params (n_chunks, P)).
    aligned / g0_idx / d_src_idx / reduce_src: fused-aggregation path; the
    gradient rows of source `reduce_src` are segment-summed and ADDED into the
    existing tensor `reduce_into` ([batch, n_seg, width]).
    g0_sum: g0 is (n, rows, d_out) for a batch-1 descriptor and dOut is its sum over n (the
    backward of an expand()); summed inside the fused kernel's loads where that exists."""
    lib = L.load()
    dev = srcs[0][0].device
    bd = L.RowMlpBwd()
    _fill_desc(bd.fwd, srcs, W, batch, rows, residual, tiles, None, precision, aligned)
    bd.reduce_src = -1
    if g0_idx is not None:
        bd.g0_idx = g0_idx.data_ptr()
    keep = []
    g0_copies = 0
    if g0 is not None:
        if g0_sum and g0.dim() == 3 and g0.shape[0] > 1 and g0.stride(0) == 0:
            # a stride-0 batch view (the producer already summed / averaged over the batch,
            # see batch_sum0): read the one slice shape[0] times instead of materialising it
            g0_copies, g0 = g0.shape[0], g0[0:1]
        g0 = g0.contiguous()
        _check_cuda(g0, "grad_output")
        bd.g0 = g0.data_ptr()
    if g1 is not None:
        g1 = _rows3d(g1, "g1")
        bd.g1 = g1.data_ptr()
        bd.g1_idx = g1_idx.data_ptr()
        bd.g1_scale = g1_scale.data_ptr() if g1_scale is not None else None
        bd.g1_batch_stride = g1.stride(0) if g1.shape[0] > 1 else 0
        assert g1.stride(1) == W.d_out or g1.shape[1] == 1
    # sender pre-reduction / batch-summed gradient of a batch-shared source 0: fused kernel only
    lib_fused = None
    if sp is not None or batch_sum0:
        for i, (it, need) in enumerate(zip(srcs, need_src)):  # probe with plain outputs
            bd.d_src[i] = 1 if need else None
        lib_fused = lib.nlam_rowmlp_bwd_stages(ctypes.byref(bd)) == 2
        for i in range(len(srcs)):
            bd.d_src[i] = None
        if not lib_fused:
            sp, batch_sum0 = None, 0
    if info is not None:
        info["sp"], info["batch_sum0"] = sp is not None, batch_sum0
    d_srcs = []
    for i, (it, need) in enumerate(zip(srcs, need_src)):
        t = it[0]
        if need and sp is not None and i == 1:
            g = torch.empty((batch, sp["n_sp"], t.shape[2]), device=dev, dtype=torch.float32)
            bd.d_src[i] = g.data_ptr()
            bd.sp_src, bd.n_sp = 1, sp["n_sp"]
            bd.sp_tile_ptr, bd.sp_row_ptr = sp["tile_ptr"].data_ptr(), sp["row_ptr"].data_ptr()
            bd.sp_rows = sp["rows"].data_ptr()
            d_srcs.append(g)
            continue
        if need and batch_sum0 and i == 0:
            g = torch.empty((1, rows, t.shape[2]), device=dev, dtype=torch.float32)
            bd.d_src[i] = g.data_ptr()
            bd.src0_batch_sum = batch_sum0
            if d_src_idx is not None and d_src_idx[i] is not None:
                bd.d_src_idx[i] = d_src_idx[i].data_ptr()
            d_srcs.append(g)
            continue
        if need and i == reduce_src:
            assert reduce_into is not None and reduce_into.is_contiguous()
            bd.d_src[i] = reduce_into.data_ptr()
            bd.reduce_src = i
            bd.reduce_accumulate = 1
            d_srcs.append(reduce_into)
        elif need:
            g = torch.empty((batch, rows, t.shape[2]), device=dev, dtype=torch.float32)
            bd.d_src[i] = g.data_ptr()
            if d_src_idx is not None and d_src_idx[i] is not None:
                bd.d_src_idx[i] = d_src_idx[i].data_ptr()
            d_srcs.append(g)
        else:
            d_srcs.append(None)
    if g0_sum and g0 is not None and (g0.shape[0] > 1 or g0_copies > 1):
        assert batch == 1 and g1 is None and not residual
        if g0_copies > 1:
            bd.g0_sum_count, bd.g0_sum_stride = g0_copies, 0
        else:
            bd.g0_sum_count, bd.g0_sum_stride = g0.shape[0], g0.stride(0)
        if lib.nlam_rowmlp_bwd_stages(ctypes.byref(bd)) != 2 or W.d_out != 64:
            bd.g0_sum_count, bd.g0_sum_stride = 0, 0  # no fused kernel here: sum first
            g0 = g0 * float(g0_copies) if g0_copies > 1 else g0.sum(0, keepdim=True)
            bd.g0 = g0.data_ptr()
    sink = _grad_sink(sink_params) if (W.n_chunks == 1 and sink_params is not None) else None
    if sink is not None:
        d_params = None  # gradients are accumulated in place; autograd gets None
        bd.d_params = sink.data_ptr()
        bd.params_accumulate = 1
    else:
        d_params = torch.empty((W.n_chunks, W.param_floats()), device=dev, dtype=torch.float32)
        bd.d_params = d_params.data_ptr()
    nws = lib.nlam_rowmlp_bwd_workspace(ctypes.byref(bd.fwd))
    ws = torch.empty((max(int(nws), 4),), device=dev, dtype=torch.float32)
    bd.workspace = ws.data_ptr()
    bd.workspace_floats = ws.numel()
    keep.extend([g0, g1, ws])
    timer = _timer["t"]
    if timer is None:
        # deferred reduction only when the gradients go to the sink: autograd must not
        # hand out a d_params tensor whose values do not exist yet
        # 16: the inputs were saved by the forward pass, the kernel that ran just before this
        # one (upstream backward) did not write them
        bd.stage_mask = 16
        if _state.get("defer_reduce", False) and sink is not None:
            bd.stage_mask = 8 | 16
            _deferred_keep.append((ws, sink))
        L.check(lib.nlam_rowmlp_bwd_run(ctypes.byref(bd), _stream()), "nlam_rowmlp_bwd_run")
    else:
        # time the three launches separately (stage_mask) with CUDA events.
        # Algorithmic bytes: dgrad = distinct input rows + dOut rows + per-source
        # gradient rows + the bf16 a/dY/dH tile images; wgrad = input rows + images
        k = sum(it[0].shape[2] for it in srcs)
        src_bytes = sum(_src_bytes(it[0], it[1], batch) for it in srcs)
        w_bytes = 4 * W.n_chunks * W.param_floats()
        rows_all = batch * rows
        # output-gradient rows read: dense g0 rows + the DISTINCT gathered g1 rows
        dout_bytes = (4 * g0.numel() if g0 is not None else 0) \
            + (4 * g1.numel() if g1 is not None else 0)
        # source-gradient rows written: one row per (batch, row) and source, except the
        # segment-reduced source (read-modify-write of its [batch, n_seg, width] target)
        dsrc_bytes = 0
        for i, (it, need) in enumerate(zip(srcs, need_src)):
            if need and i == reduce_src:
                dsrc_bytes += 2 * 4 * reduce_into.numel()
            elif need:  # rows this launch writes (per edge, per (tile, sender), or batch-summed)
                dsrc_bytes += 4 * d_srcs[i].numel()
        dsrc = sum(it[0].shape[2] for it, n in zip(srcs, need_src) if n)
        img = 2 * (2 * W.d_hidden + W.d_out) if precision == "bf16" else 4 * (2 * W.d_hidden + W.d_out)
        fl = 2 * rows_all * (k * W.d_hidden + W.d_hidden * W.d_out)
        shape = f"rows={rows}|K={k}|dh={W.d_hidden}|dout={W.d_out}|B={batch}"
        agg = "_agg" if aligned is not None else ""
        if lib.nlam_rowmlp_bwd_stages(ctypes.byref(bd)) == 2:
            # one kernel: input rows + dOut rows in, per-source gradient rows out; no images.
            # algorithmic bytes: gradients at the size the ALGORITHM needs
            # (algo_dsrc_bytes: batch-shared inputs once, node gradients per node); the
            # implementation's own traffic (per-edge / per-batch rows) is kept beside it
            impl = src_bytes + w_bytes + dout_bytes + dsrc_bytes
            algo = impl if algo_dsrc_bytes is None else \
                src_bytes + w_bytes + dout_bytes + algo_dsrc_bytes
            stages = (
                (1, f"rowmlp_bwd_fused_{precision}{agg}|{shape}", algo,
                 (3 * fl if dsrc else 2 * fl + fl // 2), impl),
                (4, f"reduce_params_{precision}|{shape}", 2 * w_bytes, 0),
            )
        else:
            # two kernels: the a / dY / dH tile images that travel between them through HBM
            # (`img` bytes per row: bf16, or hi + lo bf16 in the fp32 split mode) are the
            # implementation's traffic, not the algorithm's
            algo_d = src_bytes + w_bytes + dout_bytes + (
                dsrc_bytes if algo_dsrc_bytes is None else algo_dsrc_bytes)
            stages = (
                (1, f"rowmlp_dgrad_{precision}{agg}|{shape}", algo_d,
                 2 * fl if dsrc else fl + fl // 2,
                 src_bytes + w_bytes + dout_bytes + dsrc_bytes + rows_all * img),
                (2, f"rowmlp_wgrad_{precision}{agg}|{shape}", src_bytes + w_bytes, fl,
                 src_bytes + rows_all * img),
                (4, f"reduce_params_{precision}|{shape}", 2 * w_bytes, 0),
            )
        for mask, tag, nbytes, flops, *impl_b in stages:
            bd.stage_mask = mask
            end = timer.start(tag, nbytes, flops, *impl_b) if timer.want(tag) else None
            L.check(lib.nlam_rowmlp_bwd_run(ctypes.byref(bd), _stream()), "nlam_rowmlp_bwd_run")
            if end is not None:
                end.record()
    return d_srcs, d_params


def segsum_raw(src, ptr, idx, n_out, scale=None, out=None, accumulate=False):
    """src (B, n, w) dense rows -> (B, n_out, w)."""
    lib = L.load()
    src = src.contiguous()
    B, _, w = src.shape
    if out is None:
        assert not accumulate
        out = torch.empty((B, n_out, w), device=src.device, dtype=torch.float32)
    d = L.SegSum()
    d.src = src.data_ptr()
    d.src_batch_stride = src.stride(0)
    d.ptr, d.idx = ptr.data_ptr(), idx.data_ptr()
    d.scale = scale.data_ptr() if scale is not None else None
    d.out = out.data_ptr()
    d.batch, d.n_out, d.width = B, n_out, w
    d.accumulate = 1 if accumulate else 0
    end = None
    if _timer["t"] is not None:
        tag = f"segsum|n_out={n_out}|m={idx.numel()}|w={w}|B={B}"
        if _timer["t"].want(tag):
            nbytes = 4 * B * w * (idx.numel() + n_out * (2 if accumulate else 1)) \
                + 4 * (idx.numel() + n_out)
            end = _timer["t"].start(tag, nbytes, B * w * idx.numel())
    L.check(lib.nlam_segsum_run(ctypes.byref(d), _stream()), "nlam_segsum_run")
    if end is not None:
        end.record()
    return out


def csr_build(key32, n_keys, want_inv_deg=False):
    """Stable sort of edge ids by key -> (ptr[n_keys+1], perm[M], inv_deg|None)."""
    lib = L.load()
    dev = key32.device
    m = key32.numel()
    ptr = torch.empty((n_keys + 1,), device=dev, dtype=torch.int32)
    perm = torch.empty((max(m, 1),), device=dev, dtype=torch.int32)
    inv = torch.empty((max(n_keys, 1),), device=dev, dtype=torch.float32) if want_inv_deg else None
    ws = torch.empty((n_keys + 1,), device=dev, dtype=torch.int32)
    L.check(lib.nlam_csr_build(key32.data_ptr(), m, n_keys, ptr.data_ptr(), perm.data_ptr(),
                               inv.data_ptr() if inv is not None else None, ws.data_ptr(),
                               _stream()), "nlam_csr_build")
    return ptr, perm[:m], inv


# ------------------------------------------------------------------ autograd
class _RowMLPFn(torch.autograd.Function):
    """out = [x +] MLP(x) on direct rows (no gather)."""

    @staticmethod
    def forward(ctx, meta, w1, b1, w2, b2, ln_g, ln_b, x):
        W = Weights(w1, b1, w2, b2, ln_g, ln_b, meta["n_chunks"])
        x3 = _rows3d(x, "mlp input")
        B, rows, _ = x3.shape
        ctx.sh_x = shadow_of(x3) if meta["precision"] == "bf16" else None
        box = meta.setdefault("_sh", {})
        out = rowmlp_fwd_raw([(x3, None, ctx.sh_x)], W, B, rows, meta["residual"], meta["tiles"],
                             meta["precision"], sh_box=box)
        ctx.meta = meta
        ctx.save_for_backward(w1, b1, w2, b2, ln_g, ln_b, x3)
        return out

    @staticmethod
    def backward(ctx, gout):
        w1, b1, w2, b2, ln_g, ln_b, x3 = ctx.saved_tensors
        meta = ctx.meta
        W = Weights(w1, b1, w2, b2, ln_g, ln_b, meta["n_chunks"])
        B, rows, _ = x3.shape
        d_srcs, d_params = rowmlp_bwd_raw(
            [(x3, None, ctx.sh_x)], W, B, rows, meta["residual"], meta["tiles"], meta["precision"],
            gout, [ctx.needs_input_grad[7]], sink_params=meta.get("params"))
        gw = W.split_grads(d_params)
        return (None, *gw, d_srcs[0])


class _RowMLPCatFn(torch.autograd.Function):
    """out = MLP([x0 | x1 | x2]) on direct rows: the concatenation along the feature
    axis happens in the kernel's gather (no cat kernel, no concatenated copy).  Sources
    with a batch dim of 1 are shared by the batch."""

    @staticmethod
    def forward(ctx, meta, w1, b1, w2, b2, ln_g, ln_b, *xs):
        W = Weights(w1, b1, w2, b2, ln_g, ln_b, 1)
        xs3 = [_rows3d(x, "mlp input") for x in xs]
        B = max(x.shape[0] for x in xs3)
        rows = xs3[0].shape[1]
        if any(x.shape[1] != rows or x.shape[0] not in (1, B) for x in xs3):
            raise ValueError(f"mlp_forward_cat: incompatible inputs {[tuple(x.shape) for x in xs3]}")
        out = rowmlp_fwd_raw([(x, None) for x in xs3], W, B, rows, False, None, meta["precision"])
        ctx.meta = meta
        ctx.n_x = len(xs3)
        ctx.save_for_backward(w1, b1, w2, b2, ln_g, ln_b, *xs3)
        return out

    @staticmethod
    def backward(ctx, gout):
        saved = ctx.saved_tensors
        w1, b1, w2, b2, ln_g, ln_b = saved[:6]
        xs3 = list(saved[6:])
        meta = ctx.meta
        W = Weights(w1, b1, w2, b2, ln_g, ln_b, 1)
        B = max(x.shape[0] for x in xs3)
        need = [bool(n) for n in ctx.needs_input_grad[7:7 + ctx.n_x]]
        d_srcs, d_params = rowmlp_bwd_raw([(x, None) for x in xs3], W, B, xs3[0].shape[1], False,
                                          None, meta["precision"], gout, need,
                                          sink_params=meta.get("params"))
        d_srcs = [g.sum(0, keepdim=True) if (g is not None and x.shape[0] == 1 and B > 1) else g
                  for g, x in zip(d_srcs, xs3)]
        return (None, *W.split_grads(d_params), *d_srcs)


def mlp_forward_cat(module, xs):
    """`module(torch.cat(xs, dim=-1))` for 2 or 3 tensors (B or 1, N, w_i) / (N, w_i) without
    materialising the concatenation.  Measured on the grid features of
    base_graph_model.py:118-131 (17 | 17 | 19 columns, 255 k rows): the three narrow
    row gathers cost more than the cat kernel saves (3.54 vs 3.46 ms per step), so the
    models keep the cat there; it pays for wide sources."""
    W = weights_of(module)
    if not 1 <= len(xs) <= L.MAX_SRC or W.n_chunks != 1:
        raise ValueError("mlp_forward_cat: 1..3 inputs, one weight set")
    xs3 = [x.unsqueeze(0) if x.dim() == 2 else x for x in xs]
    if sum(x.shape[-1] for x in xs3) != W.k:
        raise ValueError(f"mlp_forward_cat: widths {[x.shape[-1] for x in xs3]} != {W.k}")
    meta = {"precision": get_precision(), "params": W.t}
    return _RowMLPCatFn.apply(meta, *W.t, *xs3)


class _RowMLPExpandFn(torch.autograd.Function):
    """MLP(x).expand(B, -1, -1) for batch-less x (static graph features): the backward takes
    the (B, rows, d_out) gradient as it is and sums its batch slices inside the kernel's
    loads, instead of autograd's expand-backward writing and re-reading a summed copy."""

    @staticmethod
    def forward(ctx, meta, w1, b1, w2, b2, ln_g, ln_b, x):
        W = Weights(w1, b1, w2, b2, ln_g, ln_b, 1)
        x3 = _rows3d(x, "mlp input")
        box = meta.setdefault("_sh", {})
        out = rowmlp_fwd_raw([(x3, None)], W, 1, x3.shape[1], False, None, meta["precision"],
                             sh_box=box)
        ctx.meta = meta
        ctx.save_for_backward(w1, b1, w2, b2, ln_g, ln_b, x3)
        return out.expand(meta["batch"], -1, -1)

    @staticmethod
    def backward(ctx, gout):
        w1, b1, w2, b2, ln_g, ln_b, x3 = ctx.saved_tensors
        meta = ctx.meta
        W = Weights(w1, b1, w2, b2, ln_g, ln_b, 1)
        d_srcs, d_params = rowmlp_bwd_raw(
            [(x3, None)], W, 1, x3.shape[1], False, None, meta["precision"], gout,
            [ctx.needs_input_grad[7]], sink_params=meta.get("params"), g0_sum=True)
        return (None, *W.split_grads(d_params), d_srcs[0])


def mlp_forward_expand(module, x, batch):
    """`module(x).unsqueeze(0).expand(batch, -1, -1)` for x (rows, K) -- the reference's
    `expand_to_batch(embedder(static_features), B)` (base_graph_model.py:125-152)."""
    W = weights_of(module)
    if x.dim() != 2 or W.n_chunks != 1:
        raise ValueError("mlp_forward_expand: (rows, K) input, one weight set")
    meta = {"precision": get_precision(), "params": W.t, "batch": int(batch)}
    out = _RowMLPExpandFn.apply(meta, *W.t, x.unsqueeze(0))
    sh = meta.get("_sh", {}).get("out")
    if sh is not None:
        attach_shadow(out, sh.expand(int(batch), -1, -1))
    return out


def mlp_forward(module, x, residual=False):
    """Fused forward of a make_mlp module on x (..., K); `residual=True`
    returns x + MLP(x) (base_graph_model.py:143-145)."""
    blocks = blocks_of(module)
    lead = x.shape[:-1]
    x3 = x.reshape(1, -1, x.shape[-1]) if x.dim() != 3 else x
    out = _run_blocks(blocks, x3, residual, None)
    W = blocks[-1]
    res = out.reshape(*lead, W.d_out)
    sh = shadow_of(out)
    if sh is not None:
        attach_shadow(res, sh.reshape(*lead, W.d_out))
    return res


def _run_blocks(blocks, x3, residual, tiles):
    """Chain of fused-kernel launches (one per block of blocks_of) on direct rows."""
    y = x3
    for i, W in enumerate(blocks):
        meta = {"n_chunks": W.n_chunks, "residual": residual and len(blocks) == 1, "tiles": tiles,
                "precision": get_precision(),
                "params": W.t if (W.n_chunks == 1 and i == 0) else None}
        out = _RowMLPFn.apply(meta, *W.t, y)
        attach_shadow(out, meta.get("_sh", {}).get("out"))
        y = out
    if residual and len(blocks) > 1:
        y = x3 + y
    return y


def split_mlp_forward(module, x):
    """SplitMLPs.forward (interaction_net.py:151-163) as one kernel launch."""
    blocks = blocks_of(module)
    lead = x.shape[:-1]
    x3 = x.reshape(1, *x.shape[-2:]) if x.dim() == 2 else x.reshape(-1, *x.shape[-2:])
    tiles = module._tile_table(x.device)
    out = _run_blocks(blocks, x3, False, tiles)
    return out.reshape(*lead, blocks[-1].d_out)


def kernel_family(widths, d_hidden, d_out, precision):
    """nlam_rowmlp_path for an MLP over sources of these widths: 0 = fp32 FFMA kernels, 1 =
    bf16 tcgen05, 2 = fp32 on tcgen05 (split bf16 operands).  Only widths matter."""
    lib = L.load()
    desc = L.RowMlp()
    desc.n_src = len(widths)
    for i, w in enumerate(widths):
        desc.src[i].ptr, desc.src[i].width, desc.src[i].ld = 256, w, w  # (never dereferenced)
    desc.batch, desc.rows, desc.n_chunks, desc.residual_src = 1, 128, 1, -1
    desc.d_hidden, desc.d_out, desc.precision = d_hidden, d_out, _PREC[precision]
    return lib.nlam_rowmlp_path(ctypes.byref(desc))


def _use_aligned(plan, We, Wa, precision):
    """Fused-aggregation path: tensor-core kernels (bf16, or the fp32 split mode where its
    tiles fit), square d in {64, 128}, one weight set, and a graph whose in-degrees fit a
    128-row tile."""
    if not (plan.alignable and We.n_chunks == 1 and Wa.n_chunks == 1
            and We.d_hidden == We.d_out and We.d_out in (64, 128) and We.k == 3 * We.d_out
            and not _state.get("disable_aligned", False)):
        return False
    if precision == "bf16":
        return True
    d = We.d_out
    return (kernel_family((d, d, d), d, d, precision) == 2
            and kernel_family((d, d), Wa.d_hidden, Wa.d_out, precision) == 2)


class _InteractionNetFn(torch.autograd.Function):
    """Whole InteractionNet layer (interaction_net.py:86-131):
    gather -> edge MLP (+edge residual) -> segment sum/mean -> node MLP + residual."""

    @staticmethod
    def forward(ctx, meta, *args):
        ew, aw = args[0:6], args[6:12]
        send, rec, edge = args[12:15]
        plan = meta["plan"]
        We = Weights(*ew, meta["edge_chunks"])
        Wa = Weights(*aw, meta["aggr_chunks"])
        send3, rec3, edge3 = (_rows3d(send, "send_rep"), _rows3d(rec, "rec_rep"),
                              _rows3d(edge, "edge_rep"))
        B = max(send3.shape[0], rec3.shape[0], edge3.shape[0])
        prec = meta["precision"]
        M, n_rec = plan.n_edges, plan.num_rec
        if rec3.shape[1] != n_rec or edge3.shape[1] != M or send3.shape[1] < plan.n_send_idx:
            raise RuntimeError(
                f"InteractionNet: got send/rec/edge rows {send3.shape[1]}/{rec3.shape[1]}/"
                f"{edge3.shape[1]}, edge_index needs >={plan.n_send_idx}/{n_rec}/{M}")
        al = plan.aligned_tables(meta["aggr"]) if _use_aligned(plan, We, Wa, prec) else None
        bf = prec == "bf16"
        sh_send, sh_rec, sh_edge = ((shadow_of(send3), shadow_of(rec3), shadow_of(edge3))
                                    if bf else (None, None, None))
        ebox, nbox = {}, {}
        if al is not None:
            # receiver-sorted, receiver-aligned tiles: gather -> edge MLP -> (E' scatter)
            # -> per-receiver sum inside ONE kernel; the messages never reach HBM
            _, new_edge, aggr = rowmlp_fwd_raw(
                [(edge3, plan.perm, sh_edge), (send3, plan.send_sorted, sh_send),
                 (rec3, plan.recv_sorted, sh_rec)],
                We, B, M, False, None, prec, want_res=meta["update_edges"], aligned=al,
                sh_box=ebox)
        else:
            # message (+ E' = E + m as second output), then CSR segment sum / mean
            edge_out = rowmlp_fwd_raw(
                [(edge3, None, sh_edge), (send3, plan.send32, sh_send),
                 (rec3, plan.recv32, sh_rec)], We, B, M,
                False, plan.edge_tiles, prec, want_res=meta["update_edges"],
                sh_box=ebox if We.n_chunks == 1 else None)
            ebox.pop("out", None)  # (shadow of the raw messages: not used)
            new_edge = None
            if meta["update_edges"]:
                edge_out, new_edge = edge_out  # messages m_k and E' = E + m (:112)
            aggr = segsum_raw(edge_out, plan.rowptr, plan.perm, n_rec,
                              scale=plan.inv_deg if meta["aggr"] == "mean" else None)
        sh_aggr = ebox.get("agg")
        rec_out = rowmlp_fwd_raw([(rec3, None, sh_rec), (aggr, None, sh_aggr)], Wa, B, n_rec, True,
                                 plan.aggr_tiles, prec, sh_box=nbox if Wa.n_chunks == 1 else None)
        same = (send3.data_ptr() == rec3.data_ptr() and send3.shape == rec3.shape
                and send3.stride() == rec3.stride())
        meta["_sh"] = {"rec_out": nbox.get("out"), "new_edge": ebox.get("out_res")}
        meta = dict(meta, aligned=al is not None, same_send_rec=same)
        ctx.shadows = (sh_send, sh_rec, sh_edge, sh_aggr)
        ctx.meta = meta
        ctx.set_materialize_grads(False)  # unused outputs arrive as None, not zeros
        ctx.save_for_backward(*ew, *aw, send3, rec3, edge3, aggr)
        if meta["update_edges"]:
            return rec_out, new_edge
        return rec_out

    @staticmethod
    def backward(ctx, *grads):
        saved = ctx.saved_tensors
        ew, aw = saved[0:6], saved[6:12]
        send3, rec3, edge3, aggr = saved[12:16]
        meta = ctx.meta
        plan = meta["plan"]
        We = Weights(*ew, meta["edge_chunks"])
        Wa = Weights(*aw, meta["aggr_chunks"])
        prec = meta["precision"]
        B = max(send3.shape[0], rec3.shape[0], edge3.shape[0])
        M, n_rec = plan.n_edges, plan.num_rec
        d_rec_out = grads[0]
        d_edge_out = grads[1] if meta["update_edges"] else None
        dev = rec3.device
        sh_send, sh_rec, sh_edge, sh_aggr = ctx.shadows
        if d_rec_out is None:
            d_rec_out = torch.zeros((B, n_rec, Wa.d_out), device=dev)
        # node stage: R' = R + aggr_mlp([R | A])
        (dR, dA), dPa = rowmlp_bwd_raw(
            [(rec3, None, sh_rec), (aggr, None, sh_aggr)], Wa, B, n_rec, True, plan.aggr_tiles, prec,
            d_rec_out, [True, True], sink_params=meta.get("aggr_params"))
        # edge stage: dm_k = dE'_k + dA[r(k)] (/deg)
        need_send, need_rec, need_edge = ctx.needs_input_grad[13:16]
        scale = plan.inv_deg if meta["aggr"] == "mean" else None
        if meta["aligned"]:
            al = plan.aligned_tables(meta["aggr"])
            # batch-shared static edge embedding (g2m / m2g / first m2m layer): its gradient is
            # accumulated over the batch inside the kernel (1 = sum for a batch-1 tensor,
            # 2 = mean for a stride-0 expanded view, whose autograd gradient is summed again)
            shared_e = B > 1 and (edge3.shape[0] == 1 or edge3.stride(0) == 0)
            bs0 = 0
            if shared_e and need_edge and d_edge_out is None and not _state.get("no_batch_sum"):
                bs0 = 1 if edge3.shape[0] == 1 else 2
            binfo = {}
            (d_edge, dzS, d_rec), dPe = rowmlp_bwd_raw(
                [(edge3, plan.perm, sh_edge), (send3, plan.send_sorted, sh_send),
                 (rec3, plan.recv_sorted, sh_rec)],
                We, B, M, d_edge_out is not None, None, prec, d_edge_out,
                [need_edge, need_send, need_rec], g1=dA, g1_idx=plan.recv_sorted,
                g1_scale=scale, aligned=al, g0_idx=plan.perm,
                d_src_idx=[plan.perm, None, None], reduce_src=2 if need_rec else -1,
                reduce_into=dR, sink_params=meta.get("edge_params"),
                sp=plan.sp if (need_send and not _state.get("no_sender_partials")) else None,
                batch_sum0=bs0, info=binfo,
                algo_dsrc_bytes=4 * We.d_out * (
                    (M * (edge3.shape[0] if edge3.stride(0) != 0 else 1) if need_edge else 0)
                    + (B * plan.n_send_idx if need_send else 0)
                    + (2 * B * n_rec if need_rec else 0)))
            d_send = None
            if binfo.get("batch_sum0") == 2:
                d_edge = d_edge.expand(B, -1, -1)  # the batch MEAN: autograd's expand-backward
                #                                    sums the B slices back to the batch sum
            if need_send:
                # per-sender sum: over the per-edge rows, or over the (tile, sender) partials
                s_ptr, s_perm = ((plan.sp["csr_rowptr"], plan.sp["csr_perm"]) if binfo.get("sp")
                                 else (plan.ts_rowptr, plan.ts_perm))
                if (meta.get("same_send_rec") and need_rec and plan.n_send_idx == n_rec
                        and d_rec is dR and dR.shape[0] == B):
                    # sender and receiver rows are the SAME tensor (m2m layers,
                    # graph_lam.py:51-57): its gradient is d_send + d_rec -- accumulate the
                    # sender part into the receiver gradient instead of returning two
                    # tensors for autograd to add
                    segsum_raw(dzS, s_ptr, s_perm, n_rec, out=dR, accumulate=True)
                else:
                    d_send = segsum_raw(dzS, s_ptr, s_perm, plan.n_send_idx)
        else:
            (dzE, dzS, dzR), dPe = rowmlp_bwd_raw(
                [(edge3, None, sh_edge), (send3, plan.send32, sh_send),
                 (rec3, plan.recv32, sh_rec)], We, B, M,
                d_edge_out is not None,  # E' = E + m: the kernel adds dE' to the edge gradient
                plan.edge_tiles, prec, d_edge_out, [need_edge, need_send, need_rec],
                g1=dA, g1_idx=plan.recv32, g1_scale=scale, sink_params=meta.get("edge_params"))
            d_edge = dzE if need_edge else None
            d_send = None
            if need_send:
                d_send = segsum_raw(dzS, plan.t_rowptr, plan.t_perm, plan.n_send_idx)
            d_rec = None
            if need_rec:
                d_rec = segsum_raw(dzR, plan.rowptr, plan.perm, n_rec, out=dR, accumulate=True)
        if d_send is not None and send3.shape[1] > plan.n_send_idx:
            pad = torch.zeros((B, send3.shape[1] - plan.n_send_idx, d_send.shape[2]), device=dev)
            d_send = torch.cat((d_send, pad), dim=1)  # trailing senders without any edge

        def fit(g, t):  # batch-1 inputs broadcast against a larger batch
            if g is not None and t.shape[0] == 1 and g.shape[0] > 1:
                g = g.sum(0, keepdim=True)
            return g

        if meta["aligned"] and binfo.get("batch_sum0") == 1:
            pass  # d_edge is already the (1, M, d) batch sum

        return (None, *We.split_grads(dPe), *Wa.split_grads(dPa), fit(d_send, send3),
                fit(d_rec, rec3), fit(d_edge, edge3))


class _GatherMLPFn(torch.autograd.Function):
    """out = [LN](W2 . SiLU(W1 . [x0[idx0] | x1[idx1] | ...] + b1) + b2): the first block of a
    deeper edge / node MLP (`hidden_layers` > 1), sources gathered by row index (None =
    direct rows).  Backward: per-row source gradients, summed per node through the CSR of
    that index (deterministic)."""

    @staticmethod
    def forward(ctx, meta, w1, b1, w2, b2, ln_g, ln_b, *xs):
        W = Weights(w1, b1, w2, b2, ln_g, ln_b, meta["n_chunks"])
        xs3 = [_rows3d(x, "gather-mlp input") for x in xs]
        B = max(x.shape[0] for x in xs3)
        srcs = [(x, idx) for x, idx in zip(xs3, meta["idxs"])]
        out = rowmlp_fwd_raw(srcs, W, B, meta["rows"], False, meta["tiles"], meta["precision"])
        ctx.meta, ctx.n_x = meta, len(xs3)
        ctx.save_for_backward(w1, b1, w2, b2, ln_g, ln_b, *xs3)
        return out

    @staticmethod
    def backward(ctx, gout):
        saved = ctx.saved_tensors
        w1, b1, w2, b2, ln_g, ln_b = saved[:6]
        xs3 = list(saved[6:])
        meta = ctx.meta
        W = Weights(w1, b1, w2, b2, ln_g, ln_b, meta["n_chunks"])
        B = max(x.shape[0] for x in xs3)
        need = [bool(n) for n in ctx.needs_input_grad[7:7 + ctx.n_x]]
        srcs = [(x, idx) for x, idx in zip(xs3, meta["idxs"])]
        d_rows, d_params = rowmlp_bwd_raw(srcs, W, B, meta["rows"], False, meta["tiles"],
                                          meta["precision"], gout, need)
        grads = []
        for g, x, csr in zip(d_rows, xs3, meta["csrs"]):
            if g is not None and csr is not None:  # gathered source: sum the rows per node
                rowptr, perm, n_idx = csr
                g = segsum_raw(g, rowptr, perm, n_idx)
                if x.shape[1] > n_idx:  # trailing nodes no edge refers to
                    g = torch.cat((g, g.new_zeros((g.shape[0], x.shape[1] - n_idx, g.shape[2]))), 1)
            if g is not None and x.shape[0] == 1 and g.shape[0] > 1:
                g = g.sum(0, keepdim=True)
            grads.append(g)
        return (None, *W.split_grads(d_params), *grads)


class _SegSumFn(torch.autograd.Function):
    """A[b, i] = scale[i] * sum_{k: recv(k) = i} m[b, k]  (interaction_net.py:124-131) as its
    own autograd node (deep path); backward = the gather dm[b, k] = scale * dA[b, recv(k)],
    run through the same segment-sum kernel with one-element segments."""

    @staticmethod
    def forward(ctx, plan, mean, m):
        m3 = _rows3d(m, "messages")
        ctx.plan, ctx.mean = plan, mean
        return segsum_raw(m3, plan.rowptr, plan.perm, plan.num_rec,
                          scale=plan.inv_deg if mean else None)

    @staticmethod
    def backward(ctx, dA):
        plan = ctx.plan
        if not hasattr(plan, "_unit_ptr"):
            plan._unit_ptr = torch.arange(plan.n_edges + 1, device=dA.device, dtype=torch.int32)
            plan._edge_scale = plan.inv_deg[plan.recv32.long()].contiguous()
        dm = segsum_raw(dA.contiguous(), plan._unit_ptr, plan.recv32, plan.n_edges,
                        scale=plan._edge_scale if ctx.mean else None)
        return None, None, dm


def interaction_net_deep(plan, edge_mlp, aggr_mlp, send, rec, edge, aggr, update_edges):
    """InteractionNet.forward (interaction_net.py:86-131) for `hidden_layers` > 1: the first
    block of the edge MLP gathers its three inputs, the remaining blocks run on the M edge
    rows, then segment sum, then the node MLP the same way.  (One hidden layer -- every
    BASELINE config -- takes the fully fused _InteractionNetFn instead.)"""
    e_blocks, a_blocks = blocks_of(edge_mlp), blocks_of(aggr_mlp)
    prec = get_precision()
    send3, rec3, edge3 = (_rows3d(send, "send_rep"), _rows3d(rec, "rec_rep"),
                          _rows3d(edge, "edge_rep"))
    M, n_rec = plan.n_edges, plan.num_rec
    if rec3.shape[1] != n_rec or edge3.shape[1] != M or send3.shape[1] < plan.n_send_idx:
        raise RuntimeError("InteractionNet: input rows do not match edge_index")
    W0 = e_blocks[0]
    meta = {"n_chunks": W0.n_chunks, "tiles": plan.edge_tiles, "precision": prec, "rows": M,
            "idxs": [None, plan.send32, plan.recv32],
            "csrs": [None, (plan.t_rowptr, plan.t_perm, plan.n_send_idx),
                     (plan.rowptr, plan.perm, n_rec)]}
    m = _GatherMLPFn.apply(meta, *W0.t, edge3, send3, rec3)
    m = _run_blocks(e_blocks[1:], m, False, plan.edge_tiles)
    A = _SegSumFn.apply(plan, aggr == "mean", m)
    W0 = a_blocks[0]
    meta = {"n_chunks": W0.n_chunks, "tiles": plan.aggr_tiles, "precision": prec, "rows": n_rec,
            "idxs": [None, None], "csrs": [None, None]}
    u = _GatherMLPFn.apply(meta, *W0.t, rec3, A)
    u = _run_blocks(a_blocks[1:], u, False, plan.aggr_tiles)
    rec_out = rec3 + u
    if update_edges:
        return rec_out, edge3 + m
    return rec_out


class _StateStepFn(torch.autograd.Function):
    """(new_state, loss_sum) = fused AR state update + masked squared-error term
    (nlam_state_step_fwd / nlam_state_step_bwd_run)."""

    @staticmethod
    def forward(ctx, net_out, prev, truth, diff_std, diff_mean, inv_std, interior):
        lib = L.load()
        net_out = _rows3d(net_out, "state").contiguous()
        B, N, F = net_out.shape

        def dense_rows(t):  # batch-strided slices (init_states[:, 1]) are taken as they are
            t = _rows3d(t, "state")
            if t.stride(2) != 1 or t.stride(1) != F or (B > 1 and t.stride(0) < N * F):
                t = t.contiguous()
            return t

        prev, truth = dense_rows(prev), dense_rows(truth)
        d = L.StateStep()
        new_state = torch.empty_like(net_out)
        n_part = int(lib.nlam_state_step_partials(B * N))
        partial = torch.empty((n_part,), device=net_out.device, dtype=torch.float32)
        loss_sum = torch.empty((1,), device=net_out.device, dtype=torch.float32)
        d.net_out, d.prev, d.truth = net_out.data_ptr(), prev.data_ptr(), truth.data_ptr()
        d.diff_std, d.diff_mean = diff_std.data_ptr(), diff_mean.data_ptr()
        d.inv_std = inv_std.data_ptr() if inv_std is not None else None
        d.interior = interior.data_ptr()
        d.new_state, d.loss_partial, d.loss_sum = (new_state.data_ptr(), partial.data_ptr(),
                                                   loss_sum.data_ptr())
        d.rows, d.nodes, d.features = B * N, N, F
        d.prev_batch_stride = prev.stride(0) if B > 1 else 0
        d.truth_batch_stride = truth.stride(0) if B > 1 else 0
        L.check(lib.nlam_state_step_fwd(ctypes.byref(d), _stream()), "nlam_state_step_fwd")
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(new_state, truth, diff_std, inv_std, interior)
        ctx.mark_non_differentiable()
        return new_state, loss_sum.reshape(())

    @staticmethod
    def backward(ctx, d_new, d_loss):
        lib = L.load()
        new_state, truth, diff_std, inv_std, interior = ctx.saved_tensors
        B, N, F = new_state.shape
        bd = L.StateStepBwd()
        bd.fwd.new_state, bd.fwd.truth = new_state.data_ptr(), truth.data_ptr()
        bd.fwd.diff_std, bd.fwd.interior = diff_std.data_ptr(), interior.data_ptr()
        bd.fwd.inv_std = inv_std.data_ptr() if inv_std is not None else None
        bd.fwd.rows, bd.fwd.nodes, bd.fwd.features = B * N, N, F
        bd.fwd.truth_batch_stride = truth.stride(0) if B > 1 else 0
        if d_new is not None:
            d_new = d_new.contiguous()
            bd.d_new = d_new.data_ptr()
        if d_loss is not None:
            d_loss = d_loss.contiguous()
            bd.d_loss = d_loss.data_ptr()
        need_net, need_prev = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        d_net = torch.empty_like(new_state) if need_net else None
        d_prev = torch.empty_like(new_state) if need_prev else None
        if d_net is None and d_prev is None:
            return (None,) * 7
        bd.d_net_out = d_net.data_ptr() if d_net is not None else None
        bd.d_prev = d_prev.data_ptr() if d_prev is not None else None
        L.check(lib.nlam_state_step_bwd_run(ctypes.byref(bd), _stream()), "nlam_state_step_bwd_run")
        return d_net, d_prev, None, None, None, None, None


def state_step(net_out, prev, truth, diff_std, diff_mean, inv_std, interior):
    """new_state, loss_sum for one AR step (see _StateStepFn)."""
    return _StateStepFn.apply(net_out, prev, truth, diff_std, diff_mean, inv_std, interior)
