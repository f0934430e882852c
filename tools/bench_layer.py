"""Time one InteractionNet layer type of the MEPS GraphLAM (fwd and fwd+bwd) on synthetic
inputs with bf16 shadows, for A/B runs of kernel options:  python tools/bench_layer.py m2g tma=0"""
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

entry.build()
from neural_lam_b200 import create_graph, lib, ops, synthetic, utils  # noqa: E402
from neural_lam_b200.interaction_net import InteractionNet  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "m2g"
    red = {"sp": 1, "bs": 0}
    prec = "bf16"
    for kv in sys.argv[2:]:
        k, v = kv.split("=")
        if k == "prec":  # prec=fp32 [fp32_split=0]
            prec = v
        elif k in red:
            red[k] = int(v)
        else:
            lib.load().nlam_set_option(k.encode(), int(v))
    ops.set_backward_reductions(bool(red["sp"]), bool(red["bs"]))
    dev = torch.device("cuda:0")
    B, d = 4, 64
    with tempfile.TemporaryDirectory() as root:
        ds = synthetic.meps_datastore(root, seed=3)
        gdir = os.path.join(root, "graph", "1level")
        create_graph.create_graph(gdir, ds.get_xy("state", stacked=False), n_max_levels=1)
        _, graph = utils.load_graph(gdir)
    ei = {"m2g": graph["m2g_edge_index"], "g2m": graph["g2m_edge_index"],
          "m2m": graph["m2m_edge_index"]}[which]
    if isinstance(ei, (list, tuple)):
        ei = ei[0]
    ops.set_precision(prec)
    torch.manual_seed(0)
    net = InteractionNet(ei.clone(), d, update_edges=(which == "m2m")).to(dev)
    M = ei.shape[1]
    n_rec = int(net.num_rec)
    n_send = int(ei[0].max() - ei[0].min()) + 1
    mk = lambda *s: ((ops.make_shadow if prec == "bf16" else (lambda x: x))(
        torch.randn(*s, device=dev).requires_grad_()))
    rec = mk(B, n_rec, d)
    send = rec if which == "m2m" else mk(B, n_send, d)
    edge = mk(B, M, d) if which == "m2m" else ops.expand_with_shadow(mk(M, d), B)
    print(f"{which} {prec}: M={M} n_send={n_send} n_rec={n_rec} B={B}")
    for mode in ("fwd", "fwd+bwd"):
        ts = []
        for it in range(13):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if mode == "fwd":
                with torch.no_grad():
                    net(send, rec, edge)
            else:
                out = net(send, rec, edge)
                outs = out if isinstance(out, tuple) else (out,)
                sum(o.sum() for o in outs).backward()
            e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(e0.elapsed_time(e1))
        ts.sort()
        print(f"  {mode}: median {ts[len(ts)//2]*1e3:.0f} us  min {ts[0]*1e3:.0f} us  "
              f"({B*M/ts[len(ts)//2]/1e3:.0f} M edges/s)")


if __name__ == "__main__":
    main()
