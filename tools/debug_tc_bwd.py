"""Scratch: gradient error statistics of the bf16 tensor-core backward vs the fp32 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
from neural_lam_b200 import ops, utils
from neural_lam_b200.interaction_net import InteractionNet
from oracle import port

dev = torch.device("cuda:0")

def stats(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    err = (a - b).abs()
    return f"rel_to_max {err.max() / (b.abs().max() + 1e-30):.2e} (max|ref| {b.abs().max():.2e}) nan {int(torch.isnan(a).sum())}"

def mlp_case(bp, ln, rows, B, residual=False, xgrad=True):
    torch.manual_seed(0)
    ref = port.make_mlp(bp, layer_norm=ln)
    if ln:
        with torch.no_grad():
            ref[3].weight.uniform_(0.5, 1.5); ref[3].bias.uniform_(-0.5, 0.5)
    mlp = utils.make_mlp(bp, layer_norm=ln)
    mlp.load_state_dict(ref.state_dict()); mlp = mlp.to(dev)
    x = torch.randn(B, rows, bp[0])
    w = torch.randn(B, rows, bp[-1])
    xr = x.clone().requires_grad_(xgrad)
    yr = ref(xr) + (xr if residual else 0)
    (yr * w).sum().backward()
    ops.set_precision("bf16")
    xg = x.clone().to(dev).requires_grad_(xgrad)
    yg = ops.mlp_forward(mlp, xg, residual=residual)
    (yg * w.to(dev)).sum().backward()
    torch.cuda.synchronize()
    ops.set_precision("fp32")
    print(f"mlp {bp} ln={ln} rows={rows} B={B} res={residual}:", flush=True)
    if xgrad:
        print("   dx  ", stats(xg.grad, xr.grad))
    for (n, p), (_, q) in zip(ref.named_parameters(), mlp.named_parameters()):
        print(f"   d{n:9s}", stats(q.grad, p.grad), flush=True)

def inet_case(d, M, n_send, n_rec, B, update, aggr):
    g = torch.Generator().manual_seed(d + M)
    s = torch.randint(0, n_send, (M,), generator=g) + n_rec
    r = torch.randint(0, n_rec, (M,), generator=g)
    s[0], r[0], s[1], r[1] = n_rec, 0, n_rec + n_send - 1, n_rec - 1
    ei = torch.stack((s, r))
    torch.manual_seed(3)
    ref = port.InteractionNet(ei.clone(), d, update_edges=update, aggr=aggr)
    net = InteractionNet(ei.clone(), d, update_edges=update, aggr=aggr)
    net.load_state_dict(ref.state_dict()); net = net.to(dev)
    xs = [torch.randn(B, n, d, generator=g) for n in (n_send, n_rec, M)]
    a = [x.clone().requires_grad_() for x in xs]
    b = [x.clone().to(dev).requires_grad_() for x in xs]
    o_ref = ref(*a); o_ref = o_ref if isinstance(o_ref, tuple) else (o_ref,)
    gw = torch.Generator().manual_seed(9)
    ws = [torch.randn(o.shape, generator=gw) for o in o_ref]
    sum((o * w).sum() for o, w in zip(o_ref, ws)).backward()
    ops.set_precision("bf16")
    o = net(*b); o = o if isinstance(o, tuple) else (o,)
    sum((oo * w.to(dev)).sum() for oo, w in zip(o, ws)).backward()
    torch.cuda.synchronize()
    ops.set_precision("fp32")
    print(f"inet d={d} M={M} B={B} upd={update} {aggr}:", flush=True)
    for x, y, n in zip(b, a, ("send", "rec", "edge")):
        print(f"   d{n:5s}", stats(x.grad, y.grad))
    for (n, p), (_, q) in zip(ref.named_parameters(), net.named_parameters()):
        print(f"   d{n:18s}", stats(q.grad, p.grad), flush=True)

mlp_case([64, 64, 64], True, 128, 1)
mlp_case([64, 64, 64], True, 1000, 2)
mlp_case([64, 64, 64], False, 300, 1)
mlp_case([64, 64, 64], True, 777, 2, residual=True)
mlp_case([3, 64, 64], True, 1000, 1, xgrad=False)
mlp_case([56, 64, 64], True, 700, 2)
mlp_case([64, 64, 17], False, 513, 2)
mlp_case([128, 128, 128], True, 300, 2)
mlp_case([32, 32, 32], True, 200, 1)
mlp_case([16, 16, 16], True, 200, 1)
inet_case(64, 20000, 3000, 2500, 2, True, "sum")
inet_case(64, 9000, 4000, 700, 1, False, "mean")
inet_case(128, 6000, 900, 900, 2, True, "sum")
inet_case(32, 1500, 200, 300, 3, True, "sum")
print("done")
