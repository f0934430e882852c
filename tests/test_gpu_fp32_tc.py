"""GPU: the fp32 mode on the tensor cores (csrc/rowmlp_tc.cuh `put8`: fp32 operands as split
bf16 hi + lo tiles, three UMMAs per product) against the oracle port at the fp32 tolerances
of BASELINE.json's north_star (1e-4 outputs, 1e-3 gradients), and against the FFMA kernels
(option fp32_split = 0) that it replaces where the tiles fit shared memory."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    import __graft_entry__ as entry
    entry.build()
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert not torch.isnan(a).any()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def _set_split(v):
    from neural_lam_b200 import lib
    assert lib.load().nlam_set_option(b"fp32_split", v) == 0


def _run_mlp(mlp, x, w, dev):
    xg = x.clone().to(dev).requires_grad_()
    y = mlp(xg)
    mlp.zero_grad()
    (y * w.to(dev)).sum().backward()
    return [y, xg.grad] + [p.grad.clone() for p in mlp.parameters()]


# (blueprint, layer_norm): square 64 (embedder / deep block), two K blocks, narrow input,
# output map without LayerNorm, padded widths; d = 128: the split tiles of the backward do not
# fit shared memory (hi + lo of W1, W2, the tile and the staging region = 256 KB) -> FFMA both
# times
@pytest.mark.parametrize("blueprint,ln", [([64, 64, 64], True), ([128, 64, 64], True),
                                          ([17, 64, 64], True), ([64, 64, 17], False),
                                          ([3, 16, 16], True), ([40, 24, 8], True),
                                          ([128, 128, 128], True)])
def test_mlp_split_vs_oracle_and_ffma(dev, blueprint, ln):
    from neural_lam_b200 import ops, utils
    from oracle import port
    ops.set_precision("fp32")
    torch.manual_seed(1)
    ref = port.make_mlp(blueprint, layer_norm=ln)
    mlp = utils.make_mlp(blueprint, layer_norm=ln)
    mlp.load_state_dict(ref.state_dict())
    mlp = mlp.to(dev)
    x = torch.randn(3, 1111, blueprint[0]) * 2.0
    w = torch.randn(3, 1111, blueprint[-1])
    xr = x.clone().requires_grad_()
    yr = ref(xr)
    (yr * w).sum().backward()
    want = [yr, xr.grad] + [p.grad for p in ref.parameters()]
    try:
        _set_split(1)
        got = _run_mlp(mlp, x, w, dev)
        _set_split(0)
        ffma = _run_mlp(mlp, x, w, dev)
    finally:
        _set_split(1)
    names = ["out", "dx"] + [n for n, _ in ref.named_parameters()]
    for n, a, f, b in zip(names, got, ffma, want):
        tol = 1e-4 if n == "out" else 1e-3
        assert _rel(a, b) <= tol, f"split {n}: {_rel(a, b):.2e}"
        assert _rel(f, b) <= tol, f"ffma {n}: {_rel(f, b):.2e}"
    same = all(torch.equal(a, f) for a, f in zip(got, ffma))
    # the two paths round differently: bit-equal results mean the option did not switch kernels
    assert same == (blueprint[1] == 128), "fp32_split did not select the expected kernels"


@pytest.mark.parametrize("update,aggr,chunks", [(True, "sum", False), (False, "mean", False),
                                                (True, "sum", True)])
def test_interaction_net_split(dev, update, aggr, chunks):
    from neural_lam_b200 import ops
    from neural_lam_b200.interaction_net import InteractionNet
    from oracle import port
    ops.set_precision("fp32")
    g = torch.Generator().manual_seed(5)
    M, n_send, n_rec, d, B = 5000, 500, 400, 64, 2
    s = torch.randint(0, n_send, (M,), generator=g) + n_rec
    r = torch.randint(0, n_rec, (M,), generator=g)
    ei = torch.stack((s, r))
    kw = dict(edge_chunk_sizes=[2000, 2500, 500], aggr_chunk_sizes=[150, 250]) if chunks else {}
    torch.manual_seed(3)
    ref = port.InteractionNet(ei.clone(), d, update_edges=update, aggr=aggr, **kw)
    net = InteractionNet(ei.clone(), d, update_edges=update, aggr=aggr, **kw)
    net.load_state_dict(ref.state_dict())
    net = net.to(dev)
    xs = [torch.randn(B, n, d, generator=g) for n in (n_send, n_rec, M)]
    a = [x.clone().requires_grad_() for x in xs]
    b = [x.clone().to(dev).requires_grad_() for x in xs]
    outs_r, outs_g = ref(*a), net(*b)
    outs_r = outs_r if isinstance(outs_r, tuple) else (outs_r,)
    outs_g = outs_g if isinstance(outs_g, tuple) else (outs_g,)
    ws = [torch.randn(o.shape, generator=g) for o in outs_r]
    sum((o * w).sum() for o, w in zip(outs_r, ws)).backward()
    sum((o * w.to(dev)).sum() for o, w in zip(outs_g, ws)).backward()
    for o_g, o_r in zip(outs_g, outs_r):
        assert _rel(o_g, o_r) <= 1e-4
    for x_g, x_r in zip(b, a):
        assert _rel(x_g.grad, x_r.grad) <= 1e-3
    for (n, p), (_, q) in zip(ref.named_parameters(), net.named_parameters()):
        assert _rel(q.grad, p.grad) <= 1e-3, n


def test_split_is_much_closer_than_bf16(dev):
    """The split mode is an fp32 mode: ~100x closer to the oracle than the bf16 mode."""
    from neural_lam_b200 import ops, utils
    from oracle import port
    torch.manual_seed(2)
    ref = port.make_mlp([128, 64, 64], layer_norm=True)
    mlp = utils.make_mlp([128, 64, 64], layer_norm=True)
    mlp.load_state_dict(ref.state_dict())
    mlp = mlp.to(dev)
    x = torch.randn(1, 4096, 128)
    want = ref(x)
    err = {}
    for mode in ("fp32", "bf16"):
        ops.set_precision(mode)
        with torch.no_grad():
            err[mode] = _rel(mlp(x.to(dev)), want)
    ops.set_precision("fp32")
    assert err["fp32"] < 3e-5 and err["bf16"] > 20 * err["fp32"], err
