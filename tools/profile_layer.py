"""One InteractionNet layer (MEPS m2g shape by default) fwd+bwd, for ncu captures
and per-layer edges/s timing."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
from neural_lam_b200 import ops
from neural_lam_b200.interaction_net import InteractionNet

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="m2g", choices=["m2g", "m2m", "g2m"])
ap.add_argument("--d", type=int, default=64)
ap.add_argument("--B", type=int, default=4)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
if a.kind == "m2g":
    n_send, n_rec = 6561, 63784
    r = torch.arange(n_rec).repeat_interleave(4)
    s = torch.randint(0, n_send, (4 * n_rec,), generator=g)
    order = torch.argsort(s, stable=True)  # sender-major like create_graph
    s, r = s[order], r[order]
    upd = False
elif a.kind == "g2m":
    n_send, n_rec, M = 63784, 6561, 79236
    s = torch.sort(torch.randint(0, n_send, (M,), generator=g)).values
    r = torch.randint(0, n_rec, (M,), generator=g)
    upd = False
else:
    n_send = n_rec = 6561
    M = 51520
    s = torch.sort(torch.randint(0, n_send, (M,), generator=g)).values
    r = torch.randint(0, n_rec, (M,), generator=g)
    upd = True
s[0], s[-1] = 0, n_send - 1
r[0], r[1] = 0, n_rec - 1
ei = torch.stack((s + n_rec, r))
M = ei.shape[1]
torch.manual_seed(0)
net = InteractionNet(ei, a.d, update_edges=upd).to(dev)
ops.set_precision(a.precision)
send = torch.randn(a.B, n_send, a.d, device=dev, requires_grad=True)
rec = torch.randn(a.B, n_rec, a.d, device=dev, requires_grad=True)
edge_base = torch.randn(M, a.d, device=dev, requires_grad=True)
edge = edge_base.unsqueeze(0).expand(a.B, -1, -1) if not upd else \
    torch.randn(a.B, M, a.d, device=dev, requires_grad=True)

def step():
    out = net(send, rec, edge)
    outs = out if isinstance(out, tuple) else (out,)
    sum(o.sum() for o in outs).backward()

for _ in range(2):
    step()
torch.cuda.synchronize()
timer = ops.KernelTimer()
ops.set_timer(timer)
t0 = time.perf_counter()
for _ in range(a.iters):
    step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / a.iters
ops.set_timer(None)
print(f"{a.kind} d={a.d} B={a.B} M={M}: {dt*1e3:.3f} ms/iter fwd+bwd -> {a.B*M/dt/1e6:.1f} M edges/s")
for tag, (n, ms, nb, fl) in sorted(timer.summary().items(), key=lambda kv: -kv[1][1]):
    print(f"  {ms/n*1e3:9.1f} us  {nb/(ms/n*1e-3)/1e9:8.1f} GB/s  {fl/(ms/n*1e-3)/1e12:6.1f} TF/s  {tag}")
