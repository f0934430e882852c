"""Eval / rollout path: validation_step = forward-only AR rollout +
per-step loss + entry MSEs, GraphLAM on the synthetic MEPS grid.  Prints one JSON line:
GPU samples/s (CUDA events, device-resident batches, optional CUDA graph of the whole
rollout) and the oracle port's CPU figure on one sample.

    python tools/bench_eval.py --ar-steps 10 --batch 4
"""
import argparse, json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry

ap = argparse.ArgumentParser()
ap.add_argument("--ar-steps", type=int, default=10)
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--cuda-graph", type=int, default=1)
ap.add_argument("--no-cpu", action="store_true")
a = ap.parse_args()
entry.build()
from neural_lam_b200 import config as nl_config, create_graph, models, ops, synthetic

dev = torch.device("cuda:0")
with tempfile.TemporaryDirectory() as root:
    ds = synthetic.meps_datastore(root, seed=0)
    args = synthetic.ModelArgs(hidden_dim=64, processor_layers=4, graph="1level", loss="wmse")
    args.val_steps_to_log = [1, 2, 3]
    create_graph.create_graph(os.path.join(root, "graph", "1level"),
                              ds.get_xy("state", stacked=False), n_max_levels=1, hierarchical=False)
    torch.manual_seed(42)
    model = models.GraphLAM(args, nl_config.default_config(), ds).to(dev)
    cpu = None
    if not a.no_cpu:
        from oracle import port
        torch.manual_seed(42)
        ref = port.GraphLAM(args, None, ds)
        b1 = synthetic.synthetic_batch(ds, 1, a.ar_steps, seed=1)
        torch.set_num_threads(os.cpu_count())
        with torch.no_grad():
            ref.validation_step(b1)
            t0 = time.perf_counter()
            ref.validation_step(b1)
            cpu = 1.0 / (time.perf_counter() - t0)
ops.set_precision(a.precision)
batches = [tuple(t.to(dev) for t in synthetic.synthetic_batch(ds, a.batch, a.ar_steps, seed=10 + i))
           for i in range(3)]
static = tuple(torch.empty_like(t) for t in batches[0])


def run(batch):
    for s, t in zip(static, batch):
        s.copy_(t)
    return model.validation_step(static)


for i in range(3):
    log, _ = run(batches[i % 3])
torch.cuda.synchronize()
graph = None
if a.cuda_graph:
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = model.validation_step(static)


def step(batch):
    if graph is None:
        return run(batch)
    for s, t in zip(static, batch):
        s.copy_(t)
    graph.replay()
    return out


for i in range(2):
    step(batches[i % 3])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.steps):
    log, entry_mse = step(batches[i % 3])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(json.dumps({
    "metric": "GraphLAM eval rollout samples/s", "value": a.batch / (ms / 1e3), "unit": "samples/s",
    "ms_per_step": ms, "ms_per_ar_step": ms / a.ar_steps, "ar_steps": a.ar_steps, "batch": a.batch,
    "dtype": a.precision, "cuda_graph": bool(a.cuda_graph),
    "val_mean_loss": float(log["val_mean_loss"]),
    "cpu_baseline": {"value": cpu, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                     "sample": "one validation_step of batch 1, same ar_steps, oracle/port.py fp32"},
    "workload": "GraphLAM 1-level mesh, hidden_dim=64, 4 processor layers, synthetic MEPS 268x238 grid"}))
