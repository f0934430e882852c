"""Data-parallel training step: the part of Lightning's
`Trainer(strategy="ddp")` (/root/reference/neural_lam/train_model.py:276-296)
that is on the hot path -- per-rank batch, gradient MEAN all-reduce over
NCCL/NVLink overlapped with the rest of backward, identical AdamW on every
rank (ar_model.py:191-195).  One process per GPU; the graph and the weights
are replicated (they are small), samples are sharded.

Gradients live in ONE flat fp32 buffer (parameters' .grad are views into it),
split into two buckets in backward order: the decoder/processor bucket is
all-reduced on a side stream as soon as its last gradient is written, while
the encoder part of backward is still running; the encoder bucket follows at
the end.  With 0.9-5 MB of gradients both transfers are latency-bound.
"""
import os

import torch
import torch.distributed as dist

from . import ops


def init_distributed():
    """Read RANK/LOCAL_RANK/WORLD_SIZE/MASTER_* (torchrun) and set the device.
    Returns (rank, world_size, device)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
        device = torch.device("cuda", local)
        backend = "nccl"
    else:
        device = torch.device("cpu")
        backend = "gloo"
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {"device_id": device} if backend == "nccl" else {}
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, device


class FlatGradBuckets:
    """Flat gradient storage + bucketed mean all-reduce with backward overlap.

    `early_names`: parameter-name prefixes whose gradients are complete before
    backward reaches the encoder (everything downstream of the g2m encoder).
    """

    def __init__(self, named_params, world_size, early_prefixes=(), overlap=True):
        self.world = world_size
        params = [(n, p) for n, p in named_params if p.requires_grad]
        early = [(n, p) for n, p in params if n.startswith(tuple(early_prefixes))] \
            if early_prefixes else []
        late = [(n, p) for n, p in params if not (early_prefixes and
                                                  n.startswith(tuple(early_prefixes)))]
        self.order = early + late
        total = sum(p.numel() for _, p in self.order)
        dev = self.order[0][1].device
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        off = 0
        for _, p in self.order:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.n_early = sum(p.numel() for _, p in early)
        self.overlap = overlap and world_size > 1 and self.n_early > 0 and dev.type == "cuda"
        self._pending = 0
        self._early_started = False
        self._enc_pending = 0  # encoder outputs (one per AR step) whose gradient is still to come
        self._early_params = [p for _, p in early]
        self._early_work = None
        self.side = torch.cuda.Stream(device=dev) if self.overlap else None

    def zero(self):
        self.flat.zero_()
        self._pending = len(self._early_params)
        self._early_work = None
        self._early_started = False
        self._enc_pending = 0

    def expect_encoder_output(self):
        """Forward pass: one more encoder output (AR step) downstream of which the early
        bucket's parameters are used."""
        self._enc_pending += 1

    def early_ready(self, grad=None):
        """Backward hook of ONE encoder output.  An unrolled rollout has one per AR step
        and backward visits them last step first, each followed by more decoder /
        processor gradient contributions of the earlier steps: the early bucket is only
        final -- and may only be handed to NCCL -- when the FIRST step's encoder-output
        gradient has arrived, i.e. when every registered hook has fired."""
        self._enc_pending -= 1
        if self._enc_pending > 0:
            return grad
        if (self.flat.is_cuda and torch.cuda.is_current_stream_capturing()
                and not self.capture_collectives):
            return grad
        if self.overlap and self._early_work is None and not self._early_started:
            self._early_started = True
            ops.flush_param_grads()  # the early bucket's gradients must be final
            self._launch_early()
        return grad

    # True while the trainer captures a CUDA graph that contains the collectives: the early
    # all-reduce then becomes a forked branch of the graph (side stream joined into the
    # capture) that runs next to the encoder part of backward at every replay
    capture_collectives = False

    def _launch_early(self):
        # all early gradients are written on the compute stream: reduce them
        # on the side stream while backward continues
        self.side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            if torch.cuda.is_current_stream_capturing():
                dist.all_reduce(self.flat[:self.n_early], op=dist.ReduceOp.AVG)
                self._early_work = "captured"
            else:
                self._early_work = dist.all_reduce(
                    self.flat[:self.n_early], op=dist.ReduceOp.AVG, async_op=True)

    def reduce(self):
        """Finish the gradient mean over ranks (call after backward)."""
        if self.world == 1:
            return
        if (self.flat.is_cuda and torch.cuda.is_current_stream_capturing()
                and not self.capture_collectives):
            raise RuntimeError("the gradient all-reduce must stay outside this CUDA-graph capture")
        if self.flat.is_cuda:
            if self._early_work is not None:
                dist.all_reduce(self.flat[self.n_early:], op=dist.ReduceOp.AVG)
                if self._early_work != "captured":
                    self._early_work.wait()
                torch.cuda.current_stream().wait_stream(self.side)
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        else:  # gloo has no AVG
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(self.world)


# gradients of these sub-modules are final before backward enters the encoder
# (the static-feature embedders are created first in predict_step, so autograd runs
# their backward last: they belong to the late bucket)
EARLY_PREFIXES = ("output_map", "m2g_gnn", "processor", "mesh_down_gnns", "mesh_down_same_gnns",
                  "mesh_up_gnns", "mesh_up_same_gnns", "mesh_init_gnns", "mesh_read_gnns")


class DataParallelTrainer:
    """`step(batch)`: forward (AR rollout + loss), backward, gradient mean over
    ranks, AdamW.  `step_from_host(batch)`: same, starting from pinned host
    tensors and ending with the loss on the host (the end-to-end call)."""

    def __init__(self, model, rank=0, world_size=1, overlap=True, use_cuda_graph=False):
        """use_cuda_graph: capture forward + backward + gradient mean + AdamW of one
        step into a CUDA graph at the first call and replay it afterwards (static
        shapes; removes the per-launch host cost of the ~470 kernels of a step --
        GraphLAM -- or of HiLAM's 64 dependent layers)."""
        self.model = model
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self._graph_has_collectives = False
        # measured on 8 x B200 (bench.py, GraphLAM configs[1]): collectives inside the graph
        # 3.410 ms/step, outside 3.378 ms/step (2 GPUs: 3.377 vs 3.364) -- NCCL's kernels as
        # graph nodes plus the fork/join cost more than the ~40 us of all-reduce they hide, so
        # the default keeps them outside; NLAM_GRAPH_NCCL=1 captures them
        self.graph_collectives = os.environ.get("NLAM_GRAPH_NCCL", "0") == "1"
        self._static_batch = None
        self._static_loss = None
        self.rank, self.world = rank, world_size
        self.device = next(model.parameters()).device
        if world_size > 1:
            # same initial weights everywhere (DDP broadcasts rank 0's)
            for p in model.parameters():
                dist.broadcast(p.data, src=0)
        self.buckets = FlatGradBuckets(list(model.named_parameters()), world_size,
                                       EARLY_PREFIXES, overlap)
        self.optimizer = model.configure_optimizers()
        # During THIS trainer's forward/backward (and only then, see _forward_backward)
        # weight gradients go straight into the flat buffer (no per-parameter
        # accumulation kernels; autograd never "sees" them, so the early all-reduce is
        # triggered by the backward of the encoder output instead) and their per-MLP
        # partial reductions are queued and run by one launch (ops.flush_param_grads)
        # after backward / before the early all-reduce.
        if self.buckets.overlap and hasattr(model, "g2m_gnn"):
            model.g2m_gnn.register_forward_hook(self._encoder_output_hook)

    def _encoder_output_hook(self, _module, _inputs, output):
        # fires in backward once everything downstream of the g2m encoder (decoder,
        # processor) has been differentiated, i.e. its kernels are enqueued
        if torch.is_tensor(output) and output.requires_grad and self._in_step:
            self.buckets.expect_encoder_output()
            output.register_hook(self.buckets.early_ready)

    _in_step = False

    def _forward_backward(self, batch):
        self.buckets.zero()
        # the gradient sink / deferred reduction are process-wide switches of ops: scope
        # them to this call so that plain `loss.backward(); optimizer.step()` elsewhere in
        # the process (another model, the eval helpers) keeps its complete .grad
        prev = ops.scoped_trainer_flags(True, True)
        self._in_step = True
        try:
            loss = self.model.training_step(batch)
            loss.backward()
            ops.flush_param_grads()  # queued parameter-gradient reductions: one launch
        except BaseException:
            ops.discard_param_grads()  # never leave pointers to dead workspaces queued
            raise
        finally:
            self._in_step = False
            ops.scoped_trainer_flags(*prev)
        return loss.detach()

    def _eager_step(self, batch):
        loss = self._forward_backward(batch)
        self.buckets.reduce()
        self.optimizer.step()
        return loss

    def _capture(self, batch):
        self._static_batch = tuple(torch.empty_like(t) for t in batch)
        for dst, src in zip(self._static_batch, batch):
            dst.copy_(src)
        # warm-up on a side stream (allocator pools, CSR plans, kernel attributes,
        # lazily created optimizer state); weights and optimizer state are put
        # back IN PLACE afterwards (the graph refers to these very tensors)
        params = [p for p in self.model.parameters()]
        saved_params = [p.detach().clone() for p in params]
        had_state = len(self.optimizer.state) > 0
        saved_state = {id(p): {k: (v.clone() if torch.is_tensor(v) else v)
                               for k, v in st.items()}
                       for p, st in self.optimizer.state.items()} if had_state else None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._eager_step(self._static_batch)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad():
            for p, q in zip(params, saved_params):
                p.copy_(q)
            for p, st in self.optimizer.state.items():
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if had_state:
                            v.copy_(saved_state[id(p)][k])
                        else:
                            v.zero_()
        # One process: the whole step is one graph.  Several ranks, default: forward +
        # backward are the graph, one all-reduce and AdamW are launched after each replay.
        # graph_collectives (NLAM_GRAPH_NCCL=1): the whole step is the graph, NCCL
        # all-reduces included -- the decoder / processor bucket's all-reduce is a forked
        # branch next to the encoder part of backward, the encoder bucket's follows, AdamW
        # closes the graph (works with capture_error_mode="thread_local"; not faster, see
        # __init__).
        self._graph = torch.cuda.CUDAGraph()
        self._graph_has_collectives = self.world > 1 and self.graph_collectives
        if self._graph_has_collectives:
            self.buckets.capture_collectives = True
            try:
                # thread_local: the NCCL watchdog thread's event queries must not abort
                # the capture
                with torch.cuda.graph(self._graph, capture_error_mode="thread_local"):
                    self._static_loss = self._eager_step(self._static_batch)
            finally:
                self.buckets.capture_collectives = False
            return
        with torch.cuda.graph(self._graph):
            if self.world == 1:
                self._static_loss = self._eager_step(self._static_batch)
            else:
                self._static_loss = self._forward_backward(self._static_batch)

    def _replay(self):
        self._graph.replay()
        if self.world > 1 and not self._graph_has_collectives:
            self.buckets._early_work = None  # no hook fired: whole buffer in one all-reduce
            self.buckets.reduce()
            self.optimizer.step()

    def step(self, batch):
        if not self.use_cuda_graph:
            return self._eager_step(batch)
        if self._graph is None:
            self._capture(batch)
        for dst, src in zip(self._static_batch, batch):
            dst.copy_(src, non_blocking=True)
        self._replay()
        return self._static_loss

    def fit_from_host(self, host_batches):
        """Train on a sequence of pinned host batches (what a DataLoader with
        pin_memory=True yields): the host->device copy of batch i+1 runs on a copy
        stream while step i computes; each step's loss is read back to pinned host
        memory asynchronously and returned as a list of floats at the end.
        Every copy happens inside this call (weather_dataset.py:603-696 hands
        batches to Lightning the same way)."""
        host_batches = list(host_batches)
        if not host_batches:
            return []
        dev = self.device
        copy_stream = torch.cuda.Stream(device=dev)
        compute = torch.cuda.current_stream()
        staging = [tuple(torch.empty_like(t, device=dev) for t in host_batches[0])
                   for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]   # H2D into staging[k] finished
        consumed = [torch.cuda.Event() for _ in range(2)]  # step reading staging[k] finished
        losses = torch.empty(len(host_batches), dtype=torch.float32).pin_memory()

        def upload(i):
            k = i % 2
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(consumed[k])
                for dst, src in zip(staging[k], host_batches[i]):
                    dst.copy_(src, non_blocking=True)
                ready[k].record(copy_stream)

        upload(0)
        for i in range(len(host_batches)):
            if i + 1 < len(host_batches):
                upload(i + 1)
            k = i % 2
            compute.wait_event(ready[k])
            loss = self.step(staging[k])  # graph mode: D2D into the static batch + replay
            consumed[k].record(compute)
            losses[i:i + 1].copy_(loss.reshape(1), non_blocking=True)
        compute.synchronize()
        return losses.tolist()

    def fit_from_feed(self, feed, index_batches, new_slices=None):
        """Train from a device-resident series (device_feed.DeviceWeatherFeed): per step the
        host sends the sample indices only -- and, when streaming (`new_slices`: one tuple
        (state (k, N, d), forcing (k, N, f) | None, times (k,) | None) of pinned host tensors
        per step), the NEWEST time slice(s), whose upload + standardisation run on a copy
        stream under the previous step's compute.  The batch is assembled on the device by
        one kernel; each step's loss is read back asynchronously.  Replaces the per-batch
        70 MB host feed of `fit_from_host` (the reference's DataLoader path,
        weather_dataset.py:603-696) by 6 MB per new time step."""
        index_batches = [list(ib) for ib in index_batches]
        if not index_batches:
            return []
        compute = torch.cuda.current_stream()
        copy_stream = torch.cuda.Stream(device=self.device) if new_slices is not None else None
        uploaded = torch.cuda.Event()
        assembled = torch.cuda.Event()
        losses = torch.empty(len(index_batches), dtype=torch.float32).pin_memory()
        static = None
        for i, idx in enumerate(index_batches):
            if new_slices is not None:
                # the ring slots being overwritten may still be read by the previous step's
                # batch assembly: order the upload after it
                copy_stream.wait_event(assembled) if i > 0 else copy_stream.wait_stream(compute)
                feed.append(*new_slices[i], stream=copy_stream)
                uploaded.record(copy_stream)
                compute.wait_event(uploaded)
            if self.use_cuda_graph and self._graph is not None and static is None:
                static = self._static_batch
            if static is not None and len(idx) == static[0].shape[0]:
                batch = feed.batch(idx, out=static)  # straight into the graph's input tensors
                assembled.record(compute)
                self._replay()
                loss = self._static_loss
            else:
                batch = feed.batch(idx)
                assembled.record(compute)
                loss = self.step(batch)
            losses[i:i + 1].copy_(loss.reshape(1), non_blocking=True)
        compute.synchronize()
        return losses.tolist()

    def step_from_host(self, host_batch):
        if self.use_cuda_graph and self._graph is not None:
            for dst, src in zip(self._static_batch, host_batch):
                dst.copy_(src, non_blocking=True)  # pinned host -> static device batch
            self._replay()
            return float(self._static_loss.item())
        batch = tuple(t.to(self.device, non_blocking=True) for t in host_batch)
        loss = self.step(batch)
        return float(loss.item())  # device -> host read of the step's result
