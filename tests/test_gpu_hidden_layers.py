"""GPU: `--hidden_layers` > 1 (train_model.py:94, utils.py:191-214) -- deeper MLPs run as a
chain of fused-kernel blocks (ops.blocks_of); parity against the oracle port in both modes."""
import tempfile

import pytest
import torch

from helpers import inet_loss

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    import __graft_entry__ as entry
    entry.build()
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _close(a, b, what, tol):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert not torch.isnan(a).any(), f"{what}: NaN"
    scale = b.abs().max().item() + 1e-30
    err = (a - b).abs().max().item()
    assert err <= tol * scale, f"{what}: max err {err:.3e} > {tol} * {scale:.3e}"


@pytest.fixture(params=["fp32", "bf16"])
def mode(request):
    from neural_lam_b200 import ops
    ops.set_precision(request.param)
    yield (1e-4, 1e-3) if request.param == "fp32" else (2e-2, 2e-2)
    ops.set_precision("fp32")


@pytest.mark.parametrize("blueprint,ln", [([56, 64, 64, 64], True), ([64, 64, 64, 64, 17], False),
                                          ([3, 16, 16, 16, 16], True)])
def test_deep_mlp(dev, mode, blueprint, ln):
    from neural_lam_b200 import utils
    from oracle import port
    torch.manual_seed(0)
    ref = port.make_mlp(blueprint, layer_norm=ln)
    mlp = utils.make_mlp(blueprint, layer_norm=ln)
    assert list(mlp.state_dict()) == list(ref.state_dict())  # same keys as the reference
    mlp.load_state_dict(ref.state_dict())
    mlp = mlp.to(dev)
    x = torch.randn(2, 700, blueprint[0])
    w = torch.randn(2, 700, blueprint[-1])
    xr, xg = x.clone().requires_grad_(), x.clone().to(dev).requires_grad_()
    yr, yg = ref(xr), mlp(xg)
    _close(yg, yr, "out", mode[0])
    (yr * w).sum().backward()
    (yg * w.to(dev)).sum().backward()
    _close(xg.grad, xr.grad, "dx", mode[1])
    for (n, p), (_, q) in zip(ref.named_parameters(), mlp.named_parameters()):
        _close(q.grad, p.grad, f"d{n}", mode[1])


@pytest.mark.parametrize("h,update,aggr,chunks", [(2, True, "sum", False), (3, False, "mean", False),
                                                  (2, True, "sum", True)])
def test_deep_interaction_net(dev, mode, h, update, aggr, chunks):
    from neural_lam_b200.interaction_net import InteractionNet
    from oracle import port
    g = torch.Generator().manual_seed(h)
    M, n_send, n_rec, d, B = 3000, 400, 300, 64, 2
    s = torch.randint(0, n_send, (M,), generator=g) + n_rec
    r = torch.randint(0, n_rec, (M,), generator=g)
    s[0], r[0], s[1], r[1] = n_rec, 0, n_rec + n_send - 1, n_rec - 1
    ei = torch.stack((s, r))
    kw = dict(edge_chunk_sizes=[1000, 1500, 500], aggr_chunk_sizes=[100, 200]) if chunks else {}
    torch.manual_seed(3)
    ref = port.InteractionNet(ei.clone(), d, update_edges=update, aggr=aggr, hidden_layers=h, **kw)
    net = InteractionNet(ei.clone(), d, update_edges=update, aggr=aggr, hidden_layers=h, **kw)
    assert list(net.state_dict()) == list(ref.state_dict())
    net.load_state_dict(ref.state_dict())
    net = net.to(dev)
    xs = [torch.randn(B, n, d, generator=g) for n in (n_send, n_rec, M)]
    a = [x.clone().requires_grad_() for x in xs]
    b = [x.clone().to(dev).requires_grad_() for x in xs]
    o_ref, o = ref(*a), net(*b)
    o_ref = o_ref if isinstance(o_ref, tuple) else (o_ref,)
    o = o if isinstance(o, tuple) else (o,)
    for x, y in zip(o, o_ref):
        _close(x, y, "output", mode[0])
    inet_loss(o_ref).backward()
    inet_loss(o).backward()
    for x, y, n in zip(b, a, ("send", "rec", "edge")):
        _close(x.grad, y.grad, f"grad {n}", mode[1])
    for (n, p), (_, q) in zip(ref.named_parameters(), net.named_parameters()):
        _close(q.grad, p.grad, f"grad {n}", mode[1])


@pytest.mark.parametrize("model_name,graph,hier", [("graph_lam", "multiscale", False),
                                                   ("hi_lam", "hierarchical", True)])
def test_deep_model_train_step(dev, model_name, graph, hier):
    """GraphLAM / HiLAM with hidden_layers = 2 (fp32 mode): loss and every gradient against
    the oracle port."""
    from helpers import build_model_case
    from neural_lam_b200 import config as nl_config
    from neural_lam_b200 import models
    from oracle import port
    case = dict(store="dummy", n_1d=60, graph=dict(n_max_levels=None, hierarchical=hier),
                args=dict(hidden_dim=16, hidden_layers=2, processor_layers=2, loss="wmse",
                          graph=graph), B=2, ar_steps=2)
    with tempfile.TemporaryDirectory() as root:
        ds, args, batch = build_model_case(case, root)
        torch.manual_seed(42)
        ref = port.MODELS[model_name](args, None, ds)
        model = models.MODELS[model_name](args, nl_config.default_config(), ds)
    model.load_state_dict(ref.state_dict())
    model = model.to(dev)
    loss_ref = ref.training_step(batch)
    loss_ref.backward()
    loss = model.training_step(tuple(t.to(dev) for t in batch))
    loss.backward()
    _close(loss, loss_ref, "loss", 1e-4)
    for (n, p), (_, q) in zip(ref.named_parameters(), model.named_parameters()):
        _close(q.grad, p.grad, f"grad {n}", 2e-3)
