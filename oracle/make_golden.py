"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.pt.

Runs the UNMODIFIED reference (/root/reference, behind oracle/ref_stubs.py)
on seeded inputs, asserts that oracle/port.py reproduces it, and stores the
reference's outputs/gradients as golden vectors.  Can only run in the build
container (the GPU box has no /root/reference); the vectors it writes are
committed so the tests can run anywhere.

    python oracle/make_golden.py            # all fixtures
    python oracle/make_golden.py --skip-meps  # only the small ones
"""
import argparse
import hashlib
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import port, ref_stubs  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def tensor_sha(t):
    return hashlib.sha256(t.contiguous().cpu().numpy().tobytes()).hexdigest()


# ---------------------------------------------------------------- layer cases
def inet_cases():
    """(name, kwargs) for single-InteractionNet fixtures.  Edge lists are
    random with GLOBAL node ids (senders and receivers from disjoint or shared
    ranges, with gaps at the low end so that the per-row min subtraction of
    interaction_net.py:56 matters)."""
    g = torch.Generator().manual_seed(1234)

    def rand_edges(m, n_send, n_rec, send_off, rec_off, force_max=True):
        s = torch.randint(0, n_send, (m,), generator=g) + send_off
        r = torch.randint(0, n_rec, (m,), generator=g) + rec_off
        if force_max:
            r[-1] = rec_off + n_rec - 1
            s[0] = send_off + n_send - 1
        return torch.stack((s, r))

    cases = []
    # m2m-like: same node set, update_edges, sum
    cases.append(("m2m_sum_d16", dict(edge_index=rand_edges(90, 20, 20, 0, 0), d=16, n_send=20,
                                      n_rec=20, update_edges=True, aggr="sum", same=True, B=2)))
    cases.append(("m2m_mean_d8", dict(edge_index=rand_edges(70, 15, 15, 0, 0), d=8, n_send=15,
                                      n_rec=15, update_edges=True, aggr="mean", same=True, B=3)))
    # g2m-like: senders numbered after receivers, lowest sender unused (quirk E1)
    ei = rand_edges(120, 50, 12, 13, 0)
    ei[0][ei[0] == 13] = 14
    cases.append(("g2m_shift_d16", dict(edge_index=ei, d=16, n_send=49, n_rec=12,
                                        update_edges=False, aggr="sum", same=False, B=2,
                                        expand_edges=True, expand_rec=True)))
    # m2g-like: receivers numbered after senders, in-degree 4, static edge emb
    r = torch.arange(30).repeat_interleave(4) + 9
    s = torch.randint(0, 9, (120,), generator=g)
    s[0] = 8
    s[1] = 0
    cases.append(("m2g_d32", dict(edge_index=torch.stack((s, r)), d=32, n_send=9, n_rec=30,
                                  update_edges=False, aggr="sum", same=False, B=2,
                                  expand_edges=True)))
    # HiLAMParallel-like: chunked edge and node MLPs
    cases.append(("split_d16", dict(edge_index=rand_edges(100, 25, 25, 0, 0), d=16, n_send=25,
                                    n_rec=25, update_edges=True, aggr="sum", same=True, B=2,
                                    edge_chunk_sizes=[40, 35, 25], aggr_chunk_sizes=[15, 10])))
    # in-degree > 128 on one receiver (forces the non-tile-aligned path)
    ei = rand_edges(400, 30, 10, 10, 0)
    ei[1][:300] = 3
    cases.append(("bigdeg_d16", dict(edge_index=ei, d=16, n_send=30, n_rec=10,
                                     update_edges=True, aggr="mean", same=False, B=1)))
    return cases


def run_inet(cls, case, state_dict=None):
    torch.manual_seed(7)
    kw = {}
    for k in ("edge_chunk_sizes", "aggr_chunk_sizes"):
        if k in case:
            kw[k] = case[k]
    net = cls(case["edge_index"].clone(), case["d"], update_edges=case["update_edges"],
              aggr=case["aggr"], **kw)
    if state_dict is not None:
        net.load_state_dict(state_dict)
    else:
        # non-trivial LayerNorm affine so its gradients are exercised
        g = torch.Generator().manual_seed(11)
        with torch.no_grad():
            for n, p in net.named_parameters():
                if n.endswith("3.weight"):
                    p.copy_(1.0 + 0.3 * torch.randn(p.shape, generator=g))
                if n.endswith("3.bias"):
                    p.copy_(0.3 * torch.randn(p.shape, generator=g))
    g = torch.Generator().manual_seed(5)
    B, d = case["B"], case["d"]
    M = case["edge_index"].shape[1]

    def leaf(n, expand):
        if expand:
            base = torch.randn(n, d, generator=g).requires_grad_()
            return base, base.unsqueeze(0).expand(B, -1, -1)
        base = torch.randn(B, n, d, generator=g).requires_grad_()
        return base, base

    rec_leaf, rec = leaf(case["n_rec"], case.get("expand_rec", False))
    if case["same"]:
        send_leaf, send = rec_leaf, rec
    else:
        send_leaf, send = leaf(case["n_send"], False)
    edge_leaf, edge = leaf(M, case.get("expand_edges", False))
    out = net(send, rec, edge)
    outs = out if isinstance(out, tuple) else (out,)
    gw = torch.Generator().manual_seed(9)
    loss = sum((o * torch.randn(o.shape, generator=gw)).sum() for o in outs)
    loss.backward()
    res = {
        "outputs": [o.detach().clone() for o in outs],
        "grad_rec": rec_leaf.grad.clone(),
        "grad_edge": edge_leaf.grad.clone(),
        "param_grads": {n: p.grad.clone() for n, p in net.named_parameters()},
        "state_dict": {k: v.clone() for k, v in net.state_dict().items()},
        "local_edge_index": net.edge_index.clone(),
        "num_rec": int(net.num_rec),
    }
    if not case["same"]:
        res["grad_send"] = send_leaf.grad.clone()
    return res


def check_close(a, b, what, rtol=1e-5, atol=1e-6):
    if isinstance(a, dict):
        assert a.keys() == b.keys(), what
        for k in a:
            check_close(a[k], b[k], f"{what}.{k}", rtol, atol)
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), what
        for i, (x, y) in enumerate(zip(a, b)):
            check_close(x, y, f"{what}[{i}]", rtol, atol)
    elif torch.is_tensor(a):
        if a.dtype in (torch.int64, torch.int32, torch.bool):
            assert torch.equal(a, b), what
        else:
            torch.testing.assert_close(a, b, rtol=rtol, atol=atol, msg=lambda m: f"{what}: {m}")
    else:
        assert a == b, (what, a, b)


def make_inet_golden(ref_inet_cls):
    out = {}
    for name, case in inet_cases():
        ref = run_inet(ref_inet_cls, case)
        mine = run_inet(port.InteractionNet, case, state_dict=ref["state_dict"])
        check_close(ref, mine, f"inet/{name}")
        out[name] = {"case": case, "ref": ref}
        print(f"  inet/{name}: port == reference "
              f"(M={case['edge_index'].shape[1]}, d={case['d']})")
    torch.save(out, os.path.join(GOLDEN, "interaction_net.pt"))


# ---------------------------------------------------------------- model cases
def model_cases(skip_meps):
    from neural_lam_b200 import synthetic

    cases = [
        # BASELINE config 1: reference's CPU-runnable case (tests/test_training.py:71-87
        # with hidden_dim 8): dummy datastore 100x100, 1-level graph, B=2, ar_steps=3
        dict(name="graphlam_dummy_d8", model="graph_lam", store="dummy", n_1d=100,
             graph=dict(n_max_levels=1, hierarchical=False),
             args=dict(hidden_dim=8, processor_layers=2, loss="mse", graph="1level"),
             B=2, ar_steps=3),
        dict(name="graphlam_multiscale_mean_d16", model="graph_lam", store="dummy", n_1d=90,
             graph=dict(n_max_levels=None, hierarchical=False),
             args=dict(hidden_dim=16, processor_layers=2, loss="wmse", graph="multiscale",
                       mesh_aggr="mean"),
             B=2, ar_steps=2),
        dict(name="hilam_d16", model="hi_lam", store="dummy", n_1d=90,
             graph=dict(n_max_levels=None, hierarchical=True),
             args=dict(hidden_dim=16, processor_layers=2, loss="wmse", graph="hierarchical"),
             B=2, ar_steps=2),
        dict(name="hilam_parallel_d16", model="hi_lam_parallel", store="dummy", n_1d=90,
             graph=dict(n_max_levels=None, hierarchical=True),
             args=dict(hidden_dim=16, processor_layers=2, loss="wmse", graph="hierarchical"),
             B=2, ar_steps=2),
    ]
    # NOTE: output_std=True cannot be pinned: the unmodified reference fails in
    # grid_embedder (grid_dim is computed from 2*grid_output_dim = 4*d_state at
    # ar_model.py:111-116, but predict_step concatenates 2*d_state state features).
    if not skip_meps:
        cases.append(
            # BASELINE config 2 at full size, one sample
            dict(name="graphlam_meps_d64", model="graph_lam", store="meps",
                 graph=dict(n_max_levels=1, hierarchical=False),
                 args=dict(hidden_dim=64, processor_layers=4, loss="wmse", graph="1level"),
                 B=1, ar_steps=1, summary_only=True))
    return cases, synthetic


def build_case(case, synthetic, root):
    from neural_lam_b200 import create_graph

    if case["store"] == "dummy":
        ds = synthetic.dummy_datastore(root, n_1d=case["n_1d"], seed=3)
    else:
        ds = synthetic.meps_datastore(root, seed=3)
    args = synthetic.ModelArgs(**case["args"])
    gdir = os.path.join(root, "graph", args.graph)
    create_graph.create_graph(gdir, ds.get_xy("state", stacked=False), **case["graph"])
    batch = synthetic.synthetic_batch(ds, case["B"], case["ar_steps"], seed=17)
    return ds, args, batch


def run_model(model_cls, config, ds, args, batch, state_dict=None):
    torch.manual_seed(42)
    model = model_cls(args, config, ds)
    if state_dict is not None:
        model.load_state_dict(state_dict)
    loss = model.training_step(batch)
    loss.backward()
    with torch.no_grad():
        pred, _ = model.predict_step(batch[0][:, 1], batch[0][:, 0], batch[2][:, 0])
    return model, {
        "loss": loss.detach().clone(),
        "pred_step": pred.clone(),
        "param_grads": {n: p.grad.clone() for n, p in model.named_parameters()},
        "state_dict": {k: v.clone() for k, v in model.state_dict().items()},
    }


def make_model_golden(skip_meps):
    from neural_lam import config as ref_config
    from neural_lam import models as ref_models

    ref_cls = {"graph_lam": ref_models.GraphLAM, "hi_lam": ref_models.HiLAM,
               "hi_lam_parallel": ref_models.HiLAMParallel}
    cfg = ref_config.NeuralLAMConfig(
        datastore=ref_config.DatastoreSelection(kind="mdp", config_path=""))
    cases, synthetic = model_cases(skip_meps)
    out = {}
    for case in cases:
        with tempfile.TemporaryDirectory() as root:
            ds, args, batch = build_case(case, synthetic, root)
            _, ref = run_model(ref_cls[case["model"]], cfg, ds, args, batch)
            _, mine = run_model(port.MODELS[case["model"]], cfg, ds, args, batch,
                                state_dict=ref["state_dict"])
        check_close(ref, mine, f"model/{case['name']}", rtol=1e-4, atol=1e-6)
        n_params = sum(v.numel() for v in ref["state_dict"].values())
        print(f"  model/{case['name']}: port == reference, loss {ref['loss'].item():.6f}, "
              f"{n_params} params")
        entry = {"case": case, "loss": ref["loss"], "state_dict": ref["state_dict"],
                 "batch_sha": [tensor_sha(t) for t in batch]}
        if case.get("summary_only"):
            # full-size: keep scalars / norms / a strided slice only
            entry["pred_slice"] = ref["pred_step"][:, ::997].clone()
            entry["grad_norms"] = {k: v.norm() for k, v in ref["param_grads"].items()}
            entry["grad_slices"] = {k: v.reshape(-1)[::53].clone()
                                    for k, v in ref["param_grads"].items()}
        else:
            entry["pred_step"] = ref["pred_step"]
            entry["param_grads"] = ref["param_grads"]
        out[case["name"]] = entry
    torch.save(out, os.path.join(GOLDEN, "models.pt"))


# ---------------------------------------------------------------- graph pins
def make_graph_golden(skip_meps):
    """sha256 of the reference create_graph outputs (edge_index exact; features
    to 1e-6 via stored tensors for small grids)."""
    import numpy as np
    from neural_lam import create_graph as ref_cg

    def make_xy(nx, ny, dx=2500.0):
        xy = np.zeros((nx, ny, 2))
        xy[:, :, 0] = (dx * np.arange(nx))[:, None]
        xy[:, :, 1] = (dx * np.arange(ny))[None, :]
        return xy

    specs = [("g30x28_1level", 30, 28, 1, False), ("g30x28_multiscale", 30, 28, None, False),
             ("g90x95_hier", 90, 95, None, True)]
    if not skip_meps:
        specs += [("meps_1level", 238, 268, 1, False), ("meps_multiscale", 238, 268, None, False),
                  ("meps_hier", 238, 268, None, True)]
    out = {}
    for name, nx, ny, levels, hier in specs:
        with tempfile.TemporaryDirectory() as d:
            ref_cg.create_graph(d, make_xy(nx, ny), levels, hier, False)
            entry = {"nx": nx, "ny": ny, "n_max_levels": levels, "hierarchical": hier,
                     "sha": {}, "shapes": {}, "feat_sum": {}}
            for fn in sorted(os.listdir(d)):
                t = torch.load(os.path.join(d, fn), weights_only=True)
                ts = t if isinstance(t, list) else [t]
                key = fn[:-3]
                if "edge_index" in key:
                    entry["sha"][key] = [tensor_sha(x) for x in ts]
                else:
                    entry["feat_sum"][key] = [x.double().abs().sum().item() for x in ts]
                entry["shapes"][key] = [tuple(x.shape) for x in ts]
            out[name] = entry
            print(f"  graph/{name}: ", {k: v for k, v in entry['shapes'].items() if 'index' in k})
    torch.save(out, os.path.join(GOLDEN, "graphs.pt"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-meps", action="store_true")
    a = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    ref_stubs.import_reference()
    from neural_lam.interaction_net import InteractionNet as RefInet

    print("InteractionNet fixtures")
    make_inet_golden(RefInet)
    print("model fixtures")
    make_model_golden(a.skip_meps)
    print("graph fixtures")
    make_graph_golden(a.skip_meps)
    print("golden vectors written to", GOLDEN)


if __name__ == "__main__":
    main()
