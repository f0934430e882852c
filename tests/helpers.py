"""Shared test helpers (CPU side)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def build_model_case(case, root):
    """Datastore, args and batch of a tests/golden/models.pt entry
    (same construction as oracle/make_golden.py:build_case)."""
    from neural_lam_b200 import create_graph, synthetic

    if case["store"] == "dummy":
        ds = synthetic.dummy_datastore(root, n_1d=case["n_1d"], seed=3)
    else:
        ds = synthetic.meps_datastore(root, seed=3)
    args = synthetic.ModelArgs(**case["args"])
    gdir = os.path.join(root, "graph", args.graph)
    create_graph.create_graph(gdir, ds.get_xy("state", stacked=False), **case["graph"])
    batch = synthetic.synthetic_batch(ds, case["B"], case["ar_steps"], seed=17)
    return ds, args, batch


def make_inet_inputs(case, device="cpu"):
    """Inputs of a tests/golden/interaction_net.pt entry (same RNG stream as
    oracle/make_golden.py:run_inet).  Returns leaves and the (possibly
    batch-expanded) views handed to forward."""
    g = torch.Generator().manual_seed(5)
    B, d = case["B"], case["d"]
    M = case["edge_index"].shape[1]

    def leaf(n, expand):
        if expand:
            base = torch.randn(n, d, generator=g).to(device).requires_grad_()
            return base, base.unsqueeze(0).expand(B, -1, -1)
        base = torch.randn(B, n, d, generator=g).to(device).requires_grad_()
        return base, base

    rec_leaf, rec = leaf(case["n_rec"], case.get("expand_rec", False))
    if case["same"]:
        send_leaf, send = rec_leaf, rec
    else:
        send_leaf, send = leaf(case["n_send"], False)
    edge_leaf, edge = leaf(M, case.get("expand_edges", False))
    return (send_leaf, rec_leaf, edge_leaf), (send, rec, edge)


def inet_loss(outs):
    gw = torch.Generator().manual_seed(9)
    return sum((o * torch.randn(o.shape, generator=gw).to(o.device)).sum() for o in outs)
