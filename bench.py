"""Benchmark of the InteractionNet / GraphLAM training hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--precision fp32|bf16] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one GraphLAM training step (AR rollout forward, loss, backward,
gradient mean over ranks, AdamW) on one synthetic MEPS-shaped batch per rank
(BASELINE.json configs[1]: 1-level mesh, hidden_dim 64, 4 processor layers,
268 x 238 grid, 17 state variables, random-init weights).  Rank 0 prints ONE
JSON line (contract: see the task statement / DESIGN.md "Measurement").

`--impl reference` times the reference's CPU path (oracle/port.py, the
restatement pinned to the unmodified reference by oracle/make_golden.py -- the
reference itself and PyG cannot travel to the GPU box) on the host cores, one
sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GraphLAM train samples/s"
UNIT = "samples/s"
WORKLOAD = ("GraphLAM 1-level mesh, hidden_dim=64, 4 processor layers, synthetic MEPS "
            "268x238 grid, 17 state vars, ar_steps=1 (BASELINE.json configs[1])")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("NLAM_PRECISION", "bf16"),
                    choices=["fp32", "bf16"])
    ap.add_argument("--batch", type=int, default=4, help="samples per GPU per step")
    ap.add_argument("--hidden-dim", type=int, default=64)
    ap.add_argument("--processor-layers", type=int, default=4)
    ap.add_argument("--graph", default="1level", choices=["1level", "multiscale", "hierarchical"],
                    help="mesh graph (BASELINE configs[1] = 1level; configs[2] = multiscale; "
                         "configs[3], [4] = hierarchical)")
    ap.add_argument("--model", default="graph_lam",
                    choices=["graph_lam", "hi_lam", "hi_lam_parallel"],
                    help="model family (hi_lam / hi_lam_parallel need --graph hierarchical)")
    ap.add_argument("--ar-steps", type=int, default=1, help="rollout steps per sample")
    ap.add_argument("--scale", type=int, default=1,
                    help="domain scale of the synthetic MEPS grid (2 = 536x476, configs[4])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp32-line", action="store_true")
    ap.add_argument("--cuda-graph", type=int, default=1,
                    help="replay the train step as one CUDA graph (single-GPU runs)")
    return ap.parse_args()


def make_case(root, a):
    from neural_lam_b200 import create_graph, synthetic

    ds = synthetic.meps_datastore(root, scale=a.scale, seed=0)
    args = synthetic.ModelArgs(hidden_dim=a.hidden_dim, processor_layers=a.processor_layers,
                               graph=a.graph, loss="wmse")
    create_graph.create_graph(os.path.join(root, "graph", a.graph),
                              ds.get_xy("state", stacked=False),
                              n_max_levels=1 if a.graph == "1level" else None,
                              hierarchical=a.graph == "hierarchical")
    return ds, args


def is_default_case(a):
    return (a.hidden_dim, a.processor_layers, a.graph, a.model, a.ar_steps, a.scale) == \
        (64, 4, "1level", "graph_lam", 1, 1)


def config_dict(a, extra=None):
    workload = WORKLOAD
    if not is_default_case(a):
        workload = (f"{a.model} on the {a.graph} mesh, hidden_dim={a.hidden_dim}, "
                    f"{a.processor_layers} processor layers, synthetic MEPS "
                    f"{268 * a.scale}x{238 * a.scale} grid, 17 state vars, ar_steps={a.ar_steps}")
    cfg = {"workload": workload, "global_batch": a.batch * a.gpus, "batch_per_gpu": a.batch,
           "ar_steps": a.ar_steps, "hidden_dim": a.hidden_dim,
           "processor_layers": a.processor_layers,
           "parallelism": f"dp{a.gpus}" if a.gpus > 1 else "single"}
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------ CPU arm
def _cpu_model(a):
    """(model, kind, describe): the reference's own GraphLAM / HiLAM / HiLAMParallel on the
    CPU -- the UNMODIFIED reference files (imported from /root/reference in the build
    container, from the archive oracle/stage_ref.py packed of them on the GPU box) behind
    oracle/ref_stubs.py; oracle/port.py (the restatement pinned to them) only when neither
    is present."""
    import torch
    import torch._dynamo  # noqa: F401  (torch.optim imports it lazily and it probes
    #                       importlib specs: must happen before the stub modules exist)

    from oracle import ref_stubs

    with tempfile.TemporaryDirectory() as root:
        ds, args = make_case(root, a)
        torch.manual_seed(42)
        if ref_stubs.reference_location() is not None:
            ref_stubs.import_reference()
            from neural_lam import config as ref_config
            from neural_lam import models as ref_models

            cls = {"graph_lam": ref_models.GraphLAM, "hi_lam": ref_models.HiLAM,
                   "hi_lam_parallel": ref_models.HiLAMParallel}[a.model]
            cfg = ref_config.NeuralLAMConfig(
                datastore=ref_config.DatastoreSelection(kind="mdp", config_path=""))
            model = cls(args, cfg, ds)
            kind = "reference"
            what = ("the unmodified reference files (neural_lam.models, PyG MessagePassing "
                    "restated by oracle/ref_stubs.py)")
        else:
            from oracle import port

            model = port.MODELS[a.model](args, None, ds)
            kind, what = "port", "oracle/port.py (restatement pinned to the reference)"
    return model, ds, kind, what


def time_cpu_reference(a, steps, warmup, batch):
    """The reference's CPU path on the host cores (all of them): samples/s of
    training_step + backward + AdamW at `batch` samples per step, fp32."""
    import torch

    from neural_lam_b200 import synthetic

    import contextlib

    cores = os.cpu_count()
    torch.set_num_threads(cores)
    with contextlib.redirect_stdout(sys.stderr):  # the reference prints while it builds
        model, ds, kind, what = _cpu_model(a)
    opt = model.configure_optimizers()
    batch_t = synthetic.synthetic_batch(ds, batch, a.ar_steps, seed=1)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        loss = model.training_step(batch_t)
        loss.backward()
        opt.step()
        times.append(time.perf_counter() - t0)
    timed = times[warmup:]
    total = sum(timed)
    return {"value": batch * len(timed) / total, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{len(timed)} train steps (fwd+bwd+AdamW) of batch {batch} after {warmup} "
                      f"warm-up, {what}, fp32, torch CPU with {cores} threads",
            "ms_per_step": 1e3 * total / len(timed), "batch": batch,
            "loss": float(loss.item())}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = time_cpu_reference(a, a.steps, min(a.warmup, 2), a.batch)
    line = {
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": min(a.warmup, 2), "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": config_dict(a, {"reference_sample_batch": res["batch"],
                                  "global_batch": res["batch"], "parallelism": "cpu"}),
        "impl": "reference",
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "loss": res["loss"],
    }
    print(json.dumps(line), flush=True)


def layer_edges_per_s(model, batch_size, device, peaks, iters=5):
    """BASELINE metric (i): InteractionNet fwd+bwd edges/s per layer type, on the
    model's own layers (MEPS g2m / m2m / m2g edge sets), CUDA-event timed, with the
    SURVEY.md 8(d) roofline fractions of the layer call:
      flops_fwd = B (M 8 d^2 + N_r 6 d^2), fwd+bwd = 3 x (recompute not counted)
      bytes_fwd = s d B (M(1+u) + N_s + 2 N_r) + 12 M
      bytes_bwd = s d B (M(2+u) + 2 N_s + 3 N_r) + 12 M
    (s = 4 bytes per stored element, u = 1 with update_edges, N_s counted once when it aliases
    N_r; the input edge rows of a batch-shared static edge embedding count M d s once)."""
    import torch

    from neural_lam_b200 import ops

    d = model.args.hidden_dim
    out = {}
    layers = {"g2m": (model.g2m_gnn, model.num_grid_nodes, model.num_mesh_nodes),
              "m2m": (model.processor[0], model.num_mesh_nodes, model.num_mesh_nodes),
              "m2g": (model.m2g_gnn, model.num_mesh_nodes, model.num_grid_nodes)}
    for name, (net, n_send, n_rec) in layers.items():
        M = net.edge_index.shape[1]
        mk = lambda *shape: ops.make_shadow(torch.randn(*shape, device=device,
                                                        requires_grad=True))
        # inputs as the model hands them over: fp32 master + bf16 shadow (bf16 mode)
        send = mk(batch_size, n_send, d)
        rec = send if name == "m2m" else mk(batch_size, n_rec, d)
        if net.update_edges:
            edge = mk(batch_size, M, d)
        else:  # static edge embedding shared by the batch (stride-0 expand)
            edge = ops.expand_with_shadow(mk(M, d), batch_size)

        def run():
            o = net(send, rec, edge)
            o = o if isinstance(o, tuple) else (o,)
            sum(x.sum() for x in o).backward()

        for _ in range(2):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        B, u, s = batch_size, (1 if net.update_edges else 0), 4
        ns = 0 if name == "m2m" else n_send  # aliases N_r
        e_in = M if net.update_edges else M / B  # static edge rows: once, not B x
        by_fwd = s * d * B * (e_in + M * u + ns + 2 * n_rec) + 12 * M
        by_bwd = s * d * B * (e_in + M * (1 + u) + 2 * ns + 3 * n_rec) + 12 * M
        fl = 3 * B * (M * 8 * d * d + n_rec * 6 * d * d)
        gbs, tfs = (by_fwd + by_bwd) / (ms / 1e3) / 1e9, fl / (ms / 1e3) / 1e12
        out[name] = {"edges": M, "ms_fwd_bwd": ms, "edges_per_s": batch_size * M / (ms / 1e3),
                     "algorithmic_bytes": by_fwd + by_bwd, "algorithmic_flops": fl,
                     "achieved_gbs": gbs, "achieved_tflops": tfs,
                     "hbm_frac": gbs / peaks["hbm_gbs"],
                     "tensor_frac": tfs / peaks["bf16_tflops_sustained"],
                     "frac": max(gbs / peaks["hbm_gbs"], tfs / peaks["bf16_tflops_sustained"])}
    model.zero_grad(set_to_none=False)
    return out


# ------------------------------------------------------------------ GPU arm
def run_ours(a):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry
    entry.build()
    from neural_lam_b200 import config as nl_config
    from neural_lam_b200 import lib, models, ops, synthetic, train

    # NCCL prints its version banner on stdout while the communicator is created:
    # keep stdout to the one JSON line by pointing fd 1 at stderr until then
    multi = int(os.environ.get("WORLD_SIZE", "1")) > 1
    if multi:
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
    rank, world, device = train.init_distributed()
    if multi:
        dist.barrier()  # creates the communicator
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    assert device.type == "cuda", "bench.py needs a GPU (no CPU fallback for the product path)"
    assert world == a.gpus, f"--gpus {a.gpus} but WORLD_SIZE={world}"
    ops.set_precision(a.precision)
    with tempfile.TemporaryDirectory() as root:
        ds, args = make_case(root, a)
        torch.manual_seed(42)
        model = models.MODELS[a.model](args, nl_config.default_config(), ds)
    model = model.to(device)
    use_graph = bool(a.cuda_graph)
    trainer = train.DataParallelTrainer(model, rank, world, use_cuda_graph=False)

    # rotating set of distinct batches, > 2x L2 in total, so that no step finds its
    # inputs in L2 from the previous one (activations per step are GBs anyway)
    n_rot = 4
    host = [synthetic.synthetic_batch(ds, a.batch, a.ar_steps, seed=1000 * rank + i, pin_memory=True)
            for i in range(n_rot)]
    dev_batches = [tuple(t.to(device) for t in hb) for hb in host]
    in_bytes = sum(t.numel() * t.element_size() for t in host[0])

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    L = lib.load()
    # ---- warm-up (builds CSR plans, allocator pools); find the dominant kernel
    timer = ops.KernelTimer()
    n_warm = max(a.warmup, 3)
    launches_per_step = 0
    for i in range(n_warm):
        if i == n_warm - 2:  # launches of one ordinary (untimed-kernel) eager step
            l0 = L.nlam_launch_count()
        if i == n_warm - 1:
            launches_per_step = L.nlam_launch_count() - l0
            torch.cuda.synchronize()
            ops.set_timer(timer)  # per-kernel CUDA events (reductions launched one by one)
        trainer.step(dev_batches[i % n_rot])
    torch.cuda.synchronize()
    ops.set_timer(None)
    summ = timer.summary()
    dominant = max(summ, key=lambda t: summ[t][1])
    step_kernel_ms = sum(v[1] for v in summ.values())

    # ---- roofline of the dominant kernel: CUDA events around each of its launches
    # over K eager steps (events cannot be recorded inside a replayed CUDA graph;
    # the kernels and their inputs are the same as in the timed region below)
    timer = ops.KernelTimer(only=dominant)
    ops.set_timer(timer)
    for i in range(a.steps):
        trainer.step(dev_batches[i % n_rot])
    torch.cuda.synchronize()
    ops.set_timer(None)
    n_l, dom_ms, dom_bytes, dom_flops = timer.summary()[dominant]

    if use_graph:  # capture once, then every step is one graph replay
        trainer.use_cuda_graph = True
        for i in range(2):
            trainer.step(dev_batches[i % n_rot])

    # ---- timed region: device-resident inputs
    clocks = ClockSampler(device.index or 0)
    sync_all()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        loss = trainer.step(dev_batches[i % n_rot])
    e1.record()
    sync_all()
    launches = launches_per_step * a.steps  # kernels of this library executed in the region
    clk = clocks.stop()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = a.steps * a.batch * world / (ms / 1e3)

    peaks = {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            mp = json.load(f)
        peaks = {"hbm_gbs": mp["hbm_gbs"], "bf16_tflops_sustained": mp["bf16_tflops_sustained"],
                 "src": "measured"}
    except (OSError, KeyError, ValueError):
        pass
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            traffic = json.load(f).get(dominant)
    except (OSError, ValueError):
        pass
    avg_s = dom_ms / 1e3 / n_l
    gbs = dom_bytes / avg_s / 1e9
    roofline = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": gbs / peaks["hbm_gbs"], "traffic": traffic, "kernel": dominant,
                "launches_timed": n_l, "avg_us": avg_s * 1e6,
                "algorithmic_bytes_per_launch": dom_bytes,
                "implementation_bytes_per_launch": timer.impl_bytes(dominant),
                "bytes_convention": "SURVEY 8(d): distinct fp32 rows read + gradients at the size "
                                    "the algorithm needs (batch-shared static rows once, node "
                                    "gradients per node); implementation_bytes = rows this "
                                    "kernel actually moves",
                "tflops": dom_flops / avg_s / 1e12, "peak_src": peaks["src"],
                "share_of_step_kernel_time": summ[dominant][1] / step_kernel_ms,
                "timed": f"CUDA events around each launch over {a.steps} eager steps"}

    # ---- end to end, (a) through the device feed (device_feed.DeviceWeatherFeed = the
    # reference's WeatherDataset on the GPU): the standardised series is resident in a
    # ring in HBM; every step uploads ONE new raw time step from pinned host memory
    # (copy + standardisation on a side stream), assembles its batch on the device from
    # host-chosen sample indices (the newest sample + batch-1 shuffled older ones), trains,
    # and reads the loss back
    from neural_lam_b200.device_feed import DeviceWeatherFeed
    ring = 48
    n_pre = ring - 8
    series = synthetic.synthetic_series(ds, n_pre + a.steps + 2, seed=77 + rank, pin_memory=True)
    st_h, fo_h, ti_h, (s_m, s_s, f_m, f_s) = series
    feed = DeviceWeatherFeed(st_h.shape[1], st_h.shape[2], fo_h.shape[2], s_m, s_s, f_m, f_s,
                             ar_steps=a.ar_steps, capacity=ring, ring=True, device=device)
    feed.append(st_h[:n_pre], fo_h[:n_pre], ti_h[:n_pre])
    need = 2 + a.ar_steps + 1  # time steps one sample spans (past = future = 1)
    import random
    rnd = random.Random(5)

    def feed_step_args(k):  # k-th streamed step: new slice n_pre + k, then a batch
        t_new = n_pre + k
        newest = t_new + 1 - need
        lo = max(0, t_new + 1 - ring)
        idx = [newest] + [rnd.randint(lo, newest) for _ in range(a.batch - 1)]
        return idx, (st_h[t_new:t_new + 1], fo_h[t_new:t_new + 1], ti_h[t_new:t_new + 1])

    warm = [feed_step_args(k) for k in range(2)]
    trainer.fit_from_feed(feed, [w[0] for w in warm], [w[1] for w in warm])
    sync_all()
    steps_f = [feed_step_args(k) for k in range(2, 2 + a.steps)]
    t0 = time.perf_counter()
    e0.record()
    feed_losses = trainer.fit_from_feed(feed, [s[0] for s in steps_f], [s[1] for s in steps_f])
    e1.record()
    sync_all()
    wall_feed = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([max(e0.elapsed_time(e1), 0.0)], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_feed = float(t.item())
    e2e_feed = {"value": a.steps * a.batch * world / (ms_feed / 1e3), "unit": UNIT,
                "h2d_bytes_per_step": feed.bytes_per_time_step(), "d2h_bytes_per_step": 4,
                "ms_per_step": ms_feed / a.steps, "wall_ms_per_step": wall_feed / a.steps,
                "last_loss": feed_losses[-1],
                "api": "DataParallelTrainer.fit_from_feed(DeviceWeatherFeed ring of "
                       f"{ring} time steps; per step: 1 new raw time step from pinned host "
                       "memory, standardised + windowed + batched on the GPU)"}
    del feed

    # ---- (b) the reference's own feed shape: every step's WHOLE batch starts in pinned
    # host memory (H2D inside the region, overlapped with the previous step's compute)
    trainer.fit_from_host([host[i % n_rot] for i in range(2)])
    sync_all()
    t0 = time.perf_counter()
    e0.record()
    host_losses = trainer.fit_from_host([host[i % n_rot] for i in range(a.steps)])
    e1.record()
    sync_all()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ms_e2e = max(e0.elapsed_time(e1), 0.0)
    t = torch.tensor([ms_e2e], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())
    e2e_host = {"value": a.steps * a.batch * world / (ms_e2e / 1e3), "unit": UNIT,
                "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / a.steps, "wall_ms_per_step": wall_ms / a.steps,
                "last_loss": host_losses[-1],
                "api": "DataParallelTrainer.fit_from_host(pinned batches)"}
    e2e = e2e_feed

    layers = None
    fp32_mode = None
    if rank == 0 and world == 1:
        if a.model == "graph_lam":  # per-layer figure on the model's own g2m / m2m / m2g layers
            layers = layer_edges_per_s(model, a.batch, device, peaks)
        if a.precision == "bf16" and is_default_case(a) and not a.no_fp32_line:
            # the fp32 half of configs[1] ("bf16/fp32"): same model, weights and batches in
            # the fp32 parity mode (rtol 1e-4 outputs / 1e-3 gradients vs the reference): fp32
            # operands as split bf16 tiles on the tensor cores, three UMMAs per product
            # (NLAM_FP32_SPLIT=0: the FFMA kernels); same CUDA-graph trainer as the headline
            ops.set_precision("fp32")
            tr32 = train.DataParallelTrainer(model, rank, world, use_cuda_graph=bool(a.cuda_graph))
            for i in range(4):
                tr32.step(dev_batches[i % n_rot])
            torch.cuda.synchronize()
            l0 = L.nlam_launch_count()
            e0.record()
            n32 = 10
            for i in range(n32):
                l32 = tr32.step(dev_batches[i % n_rot])
            e1.record()
            torch.cuda.synchronize()
            ms32 = e0.elapsed_time(e1) / n32
            fam = ops.kernel_family((a.hidden_dim,) * 3, a.hidden_dim, a.hidden_dim, "fp32")
            fp32_mode = {"value": a.batch / (ms32 / 1e3), "unit": UNIT, "ms_per_step": ms32,
                         "steps": n32, "dtype": "f32", "loss": float(l32.item()),
                         "kernels": "tcgen05, split bf16 operands (3 UMMAs per product)"
                                    if fam == 2 else "fp32 FFMA",
                         "note": "fp32 parity mode, device-resident batches"}
            ops.set_precision(a.precision)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        # same config as the GPU arm where that stays within ~30 s of CPU work
        cpu_batch = a.batch if is_default_case(a) else 1
        cpu = time_cpu_reference(a, 3 if is_default_case(a) else 1, 1, cpu_batch)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        act_mb = 4 * a.batch * (255136 + 79236 + 4 * 51520) * a.hidden_dim / 1e6
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if a.precision == "fp32" else "bf16",
            "data": "synthetic",
            "config": config_dict(a, {
                "precision": a.precision, "cuda_graph": use_graph,
                "l2": f"{n_rot} rotating input batches ({n_rot * in_bytes / 1e6:.0f} MB) and "
                      f"~{act_mb:.0f} MB of edge activations per step, both > 126 MB L2; "
                      "no explicit flush"}),
            "clocks": clk, "e2e": e2e, "e2e_host_batches": e2e_host,
            "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "impl": "ours", "loss": float(loss.item()),
            "interaction_net_fwd_bwd": layers, "roofline_layers": layers,
            "fp32_mode": fp32_mode,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # every collective of this run has completed on every rank; skip the NCCL /
        # CUDA-graph teardown (it can block at interpreter exit) and leave at once
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
