"""CPU: the oracle (oracle/port.py) against the golden vectors produced by the
UNMODIFIED reference (oracle/make_golden.py).  This is what pins the oracle."""
import tempfile

import pytest
import torch

from helpers import build_model_case, inet_loss, load_golden, make_inet_inputs
from oracle import port

INET = load_golden("interaction_net.pt")
MODELS = load_golden("models.pt")


@pytest.mark.parametrize("name", sorted(INET))
def test_port_interaction_net_matches_reference(name):
    case, ref = INET[name]["case"], INET[name]["ref"]
    kw = {k: case[k] for k in ("edge_chunk_sizes", "aggr_chunk_sizes") if k in case}
    net = port.InteractionNet(case["edge_index"].clone(), case["d"],
                              update_edges=case["update_edges"], aggr=case["aggr"], **kw)
    # index arithmetic is bit-exact (interaction_net.py:55-61)
    assert torch.equal(net.edge_index, ref["local_edge_index"])
    assert net.num_rec == ref["num_rec"]
    assert list(net.state_dict()) == list(ref["state_dict"])
    net.load_state_dict(ref["state_dict"])
    (send_leaf, rec_leaf, edge_leaf), (send, rec, edge) = make_inet_inputs(case)
    out = net(send, rec, edge)
    outs = out if isinstance(out, tuple) else (out,)
    for o, r in zip(outs, ref["outputs"]):
        torch.testing.assert_close(o, r, rtol=1e-5, atol=1e-6)
    inet_loss(outs).backward()
    torch.testing.assert_close(rec_leaf.grad, ref["grad_rec"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(edge_leaf.grad, ref["grad_edge"], rtol=1e-5, atol=1e-6)
    if not case["same"]:
        torch.testing.assert_close(send_leaf.grad, ref["grad_send"], rtol=1e-5, atol=1e-6)
    for n, p in net.named_parameters():
        torch.testing.assert_close(p.grad, ref["param_grads"][n], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("name", [n for n in sorted(MODELS) if "meps" not in n])
def test_port_model_matches_reference(name):
    entry = MODELS[name]
    case = entry["case"]
    with tempfile.TemporaryDirectory() as root:
        ds, args, batch = build_model_case(case, root)
        model = port.MODELS[case["model"]](args, None, ds)
    assert list(model.state_dict()) == list(entry["state_dict"])
    model.load_state_dict(entry["state_dict"])
    loss = model.training_step(batch)
    torch.testing.assert_close(loss, entry["loss"], rtol=1e-5, atol=1e-6)
    loss.backward()
    for n, p in model.named_parameters():
        torch.testing.assert_close(p.grad, entry["param_grads"][n], rtol=1e-4, atol=1e-6)
    with torch.no_grad():
        pred, _ = model.predict_step(batch[0][:, 1], batch[0][:, 0], batch[2][:, 0])
    torch.testing.assert_close(pred, entry["pred_step"], rtol=1e-5, atol=1e-5)


def test_stub_scatter_equals_dense_adjacency():
    """The only restated third-party arithmetic (PyG gather + scatter) against
    an independent dense one-hot formulation."""
    g = torch.Generator().manual_seed(0)
    M, n_s, n_r, d, B = 50, 9, 7, 4, 2
    s = torch.randint(0, n_s, (M,), generator=g)
    r = torch.randint(0, n_r, (M,), generator=g)
    r[0], s[0] = n_r - 1, n_s - 1
    s[1], r[1] = 0, 0
    for aggr in ("sum", "mean"):
        net = port.InteractionNet(torch.stack((s, r)), d, aggr=aggr)
        send, rec, edge = (torch.randn(B, n, d, generator=g) for n in (n_s, n_r, M))
        new_rec, new_edge = net(send, rec, edge)
        S = torch.nn.functional.one_hot(s, n_s).float()
        R = torch.nn.functional.one_hot(r, n_r).float()
        z = torch.cat((edge, torch.einsum("ms,bsd->bmd", S, send),
                       torch.einsum("mr,brd->bmd", R, rec)), -1)
        msg = net.edge_mlp(z)
        agg = torch.einsum("mr,bmd->brd", R, msg)
        if aggr == "mean":
            agg = agg / R.sum(0).clamp(min=1)[None, :, None]
        exp_rec = rec + net.aggr_mlp(torch.cat((rec, agg), -1))
        torch.testing.assert_close(new_rec, exp_rec, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(new_edge, edge + msg, rtol=1e-6, atol=1e-6)


def test_port_eval_path_matches_reference_golden():
    """validation_step / test_step of the oracle port against the unmodified reference's
    (tests/golden/eval.pt, written by oracle/make_golden_eval.py)."""
    ev = load_golden("eval.pt")
    models_pt = load_golden("models.pt")
    assert len(ev) >= 4
    for name, entry in ev.items():
        case = entry["case"]
        with tempfile.TemporaryDirectory() as root:
            ds, args, batch = build_model_case(case, root)
            args.val_steps_to_log = entry["val_steps_to_log"]
            model = port.MODELS[case["model"]](args, None, ds)
        model.load_state_dict(models_pt[name]["state_dict"])
        ref = entry["ref"]
        with torch.no_grad():
            vlog, vmse = model.validation_step(batch)
            tlog, tentry, spatial = model.test_step(batch)
        for k, v in ref["val_log"].items():
            torch.testing.assert_close(vlog[k], v, rtol=1e-4, atol=1e-6, msg=f"{name} {k}")
        for k, v in ref["test_log"].items():
            torch.testing.assert_close(tlog[k], v, rtol=1e-4, atol=1e-6, msg=f"{name} {k}")
        torch.testing.assert_close(vmse, ref["val_mse"], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(tentry["mse"], ref["test_mse"], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(tentry["mae"], ref["test_mae"], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(spatial, ref["spatial"], rtol=1e-4, atol=1e-6)
