"""Where do the per-step fill (zero) kernels of a train step come from?  torch.profiler with
Python stacks over one eager GraphLAM step."""
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

entry.build()
from neural_lam_b200 import config as nl_config  # noqa: E402
from neural_lam_b200 import create_graph, models, ops, synthetic, train  # noqa: E402

dev = torch.device("cuda:0")
ops.set_precision("bf16")
model_name = sys.argv[1] if len(sys.argv) > 1 else "graph_lam"
hier = model_name != "graph_lam"
with tempfile.TemporaryDirectory() as root:
    ds = synthetic.meps_datastore(root, seed=0)
    args = synthetic.ModelArgs(hidden_dim=64, processor_layers=4,
                               graph="hierarchical" if hier else "1level", loss="wmse")
    create_graph.create_graph(os.path.join(root, "graph", args.graph), ds.get_xy("state", stacked=False),
                              n_max_levels=None if hier else 1, hierarchical=hier)
    model = models.MODELS[model_name](args, nl_config.default_config(), ds).to(dev)
trainer = train.DataParallelTrainer(model)
batch = synthetic.synthetic_batch(ds, 4, 1, seed=1, device=dev)
for _ in range(3):
    trainer.step(batch)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    trainer.step(batch)
    torch.cuda.synchronize()
import collections  # noqa: E402

cnt = collections.Counter()
for ev in prof.events():
    if ev.name in ("aten::fill_", "aten::zero_", "aten::zeros", "aten::zeros_like", "aten::ones_like",
                   "aten::cat", "aten::sum", "aten::add", "aten::add_", "aten::copy_", "aten::mul",
                   "aten::index", "aten::div"):
        stack = [s for s in (ev.stack or []) if "neural-lam-dev_b200" in s or "torch/autograd" in s
                 or "optim" in s][:3]
        cnt[(ev.name, tuple(stack))] += 1
for (name, stack), n in cnt.most_common(40):
    print(n, name, " <- ".join(s.split("/")[-1] for s in stack))
