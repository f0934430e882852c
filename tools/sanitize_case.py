"""Small InteractionNet + MLP fwd/bwd in bf16 mode (fused backward kernel, multi-context
forward, narrow-input and narrow-output variants) for compute-sanitizer runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
from neural_lam_b200 import ops, utils
from neural_lam_b200.interaction_net import InteractionNet

dev = torch.device("cuda:0")
ops.set_precision("bf16")
g = torch.Generator().manual_seed(0)
d, M, n_send, n_rec, B = 64, 40000, 3000, 2500, 2
s = torch.randint(0, n_send, (M,), generator=g) + n_rec
r = torch.randint(0, n_rec, (M,), generator=g)
s[0], r[0], s[1], r[1] = n_rec, 0, n_rec + n_send - 1, n_rec - 1
for upd in (True, False):
    net = InteractionNet(torch.stack((s, r)), d, update_edges=upd).to(dev)
    xs = [torch.randn(B, n, d, generator=g).to(dev).requires_grad_() for n in (n_send, n_rec, M)]
    out = net(*xs)
    outs = out if isinstance(out, tuple) else (out,)
    sum(o.square().sum() for o in outs).backward()
for bp, ln, rows in (([3, 64, 64], True, 5000), ([64, 64, 17], False, 5000), ([64, 64, 64], True, 700)):
    mlp = utils.make_mlp(bp, layer_norm=ln).to(dev)
    x = torch.randn(B, rows, bp[0], device=dev, requires_grad=bp[0] == 64)
    mlp(x).square().sum().backward()
torch.cuda.synchronize()
print("sanitize case done")
