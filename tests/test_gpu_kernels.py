"""GPU parity tests: CUDA path (through the C ABI) vs the CPU oracle.
Tolerances (BASELINE.json north_star): fp32 outputs rtol 1e-4, gradients rtol
1e-3; index work bit-exact."""
import numpy as np
import pytest
import torch

from helpers import inet_loss, load_golden, make_inet_inputs

pytestmark = pytest.mark.gpu

RTOL_OUT, RTOL_GRAD = 1e-4, 1e-3


@pytest.fixture(scope="module")
def dev():
    import __graft_entry__ as entry
    entry.build()
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _close(a, b, rtol, what):
    a, b = a.detach().cpu(), b.detach().cpu()
    scale = b.abs().max().item() + 1e-30
    torch.testing.assert_close(a, b, rtol=rtol, atol=rtol * scale * 0.1,
                               msg=lambda m: f"{what}: {m}")


# ------------------------------------------------------------------ integer work
@pytest.mark.parametrize("m,n_keys,seed", [(0, 5, 0), (1, 1, 1), (1000, 37, 2), (50000, 6561, 3),
                                           (4096, 3, 4), (300000, 70000, 5)])
def test_csr_build_bit_exact(dev, m, n_keys, seed):
    from neural_lam_b200 import ops
    rng = np.random.default_rng(seed)
    key = rng.integers(0, n_keys, size=m).astype(np.int32)
    if m > 10:
        key[: m // 3] = key[0]  # one heavy key
    ptr, perm, inv = ops.csr_build(torch.from_numpy(key).to(dev), n_keys, True)
    torch.cuda.synchronize()
    want_perm = np.argsort(key, kind="stable").astype(np.int32)
    cnt = np.bincount(key, minlength=n_keys)
    want_ptr = np.concatenate(([0], np.cumsum(cnt))).astype(np.int32)
    assert np.array_equal(ptr.cpu().numpy(), want_ptr)
    assert np.array_equal(perm.cpu().numpy(), want_perm)
    assert np.array_equal(inv.cpu().numpy()[:n_keys],
                          (1.0 / np.maximum(cnt, 1)).astype(np.float32))


@pytest.mark.parametrize("width", [3, 16, 64, 128])
def test_segsum_matches_index_add(dev, width):
    from neural_lam_b200 import ops
    g = torch.Generator().manual_seed(width)
    m, n_out, B = 5000, 321, 3
    key = torch.randint(0, n_out, (m,), generator=g)
    src = torch.randn(B, m, width, generator=g)
    ptr, perm, inv = ops.csr_build(key.to(torch.int32).to(dev), n_out, True)
    out = ops.segsum_raw(src.to(dev), ptr, perm, n_out)
    want = torch.zeros(B, n_out, width).index_add_(1, key, src)
    # same summation order as the CPU loop -> tight tolerance
    torch.testing.assert_close(out.cpu(), want, rtol=1e-6, atol=1e-6)
    out2 = ops.segsum_raw(src.to(dev), ptr, perm, n_out, scale=inv)
    cnt = torch.bincount(key, minlength=n_out).clamp(min=1).float()
    torch.testing.assert_close(out2.cpu(), want / cnt[None, :, None], rtol=1e-6, atol=1e-6)
    base = torch.randn(B, n_out, width, generator=g)
    out3 = ops.segsum_raw(src.to(dev), ptr, perm, n_out, out=base.to(dev), accumulate=True)
    torch.testing.assert_close(out3.cpu(), base + want, rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------ plain MLPs
@pytest.mark.parametrize("blueprint,ln,rows,B", [
    ([3, 64, 64], True, 1000, 1),      # edge-feature embedder
    ([2, 16, 16], True, 77, 1),        # mesh embedder, tiny
    ([56, 64, 64], True, 700, 2),      # grid embedder (MEPS grid_dim)
    ([17, 8, 8], True, 130, 2),        # grid embedder of the dummy datastore
    ([64, 64, 17], False, 513, 2),     # output map, no LayerNorm
    ([128, 128, 128], True, 300, 2),   # d=128
    ([32, 32, 32], True, 64, 1),
])
def test_fused_mlp_matches_torch(dev, blueprint, ln, rows, B):
    from neural_lam_b200 import utils
    from oracle import port
    torch.manual_seed(0)
    ref = port.make_mlp(blueprint, layer_norm=ln)
    if ln:
        with torch.no_grad():
            ref[3].weight.uniform_(0.5, 1.5)
            ref[3].bias.uniform_(-0.5, 0.5)
    mlp = utils.make_mlp(blueprint, layer_norm=ln)
    mlp.load_state_dict(ref.state_dict())
    mlp = mlp.to(dev)
    x = torch.randn(B, rows, blueprint[0])
    xr = x.clone().requires_grad_()
    xg = x.clone().to(dev).requires_grad_()
    w = torch.randn(B, rows, blueprint[-1])
    yr = ref(xr)
    yg = mlp(xg)
    _close(yg, yr, RTOL_OUT, "mlp out")
    (yr * w).sum().backward()
    (yg * w.to(dev)).sum().backward()
    _close(xg.grad, xr.grad, RTOL_GRAD, "mlp dx")
    for (n, p), (_, q) in zip(ref.named_parameters(), mlp.named_parameters()):
        _close(q.grad, p.grad, RTOL_GRAD, f"mlp d{n}")


def test_fused_mlp_residual(dev):
    from neural_lam_b200 import ops, utils
    torch.manual_seed(1)
    mlp = utils.make_mlp([64, 64, 64]).to(dev)
    x = torch.randn(2, 333, 64, device=dev, requires_grad=True)
    x2 = x.detach().clone().requires_grad_()
    y = ops.mlp_forward(mlp, x, residual=True)
    import torch.nn as nn
    y2 = x2 + nn.Sequential.forward(mlp, x2)
    _close(y, y2, RTOL_OUT, "residual out")
    w = torch.randn_like(y)
    g1 = torch.autograd.grad((y * w).sum(), [x] + list(mlp.parameters()))
    g2 = torch.autograd.grad((y2 * w).sum(), [x2] + list(mlp.parameters()))
    for a, b in zip(g1, g2):
        _close(a, b, RTOL_GRAD, "residual grads")


# ------------------------------------------------------------------ InteractionNet
INET = load_golden("interaction_net.pt")


@pytest.mark.parametrize("name", sorted(INET))
def test_interaction_net_vs_reference_golden(dev, name):
    """CUDA InteractionNet against outputs/gradients of the UNMODIFIED
    reference (tests/golden/interaction_net.pt)."""
    from neural_lam_b200.interaction_net import InteractionNet
    case, ref = INET[name]["case"], INET[name]["ref"]
    kw = {k: case[k] for k in ("edge_chunk_sizes", "aggr_chunk_sizes") if k in case}
    net = InteractionNet(case["edge_index"].clone(), case["d"],
                         update_edges=case["update_edges"], aggr=case["aggr"], **kw)
    assert torch.equal(net.edge_index, ref["local_edge_index"])  # bit-exact index work
    assert int(net.num_rec) == ref["num_rec"]
    net.load_state_dict(ref["state_dict"])
    net = net.to(dev)
    (send_leaf, rec_leaf, edge_leaf), (send, rec, edge) = make_inet_inputs(case, device=dev)
    out = net(send, rec, edge)
    outs = out if isinstance(out, tuple) else (out,)
    for i, (o, r) in enumerate(zip(outs, ref["outputs"])):
        _close(o, r, RTOL_OUT, f"{name} output {i}")
    inet_loss(outs).backward()
    _close(rec_leaf.grad, ref["grad_rec"], RTOL_GRAD, "grad rec")
    _close(edge_leaf.grad, ref["grad_edge"], RTOL_GRAD, "grad edge")
    if not case["same"]:
        _close(send_leaf.grad, ref["grad_send"], RTOL_GRAD, "grad send")
    for n, p in net.named_parameters():
        _close(p.grad, ref["param_grads"][n], RTOL_GRAD, f"grad {n}")


@pytest.mark.parametrize("d,M,n_send,n_rec,B,update,aggr", [
    (64, 20000, 3000, 2500, 2, True, "sum"),
    (64, 9000, 4000, 700, 1, False, "sum"),
    (128, 6000, 900, 900, 2, True, "mean"),
    (32, 1500, 200, 300, 3, True, "sum"),
])
def test_interaction_net_vs_oracle(dev, d, M, n_send, n_rec, B, update, aggr):
    from neural_lam_b200.interaction_net import InteractionNet
    from oracle import port
    g = torch.Generator().manual_seed(d + M)
    s = torch.randint(0, n_send, (M,), generator=g) + n_rec
    r = torch.randint(0, n_rec, (M,), generator=g)
    s[0], r[0], s[1], r[1] = n_rec, 0, n_rec + n_send - 1, n_rec - 1
    ei = torch.stack((s, r))
    torch.manual_seed(3)
    ref = port.InteractionNet(ei.clone(), d, update_edges=update, aggr=aggr)
    net = InteractionNet(ei.clone(), d, update_edges=update, aggr=aggr)
    net.load_state_dict(ref.state_dict())
    net = net.to(dev)
    xs = [torch.randn(B, n, d, generator=g) for n in (n_send, n_rec, M)]
    a = [x.clone().requires_grad_() for x in xs]
    b = [x.clone().to(dev).requires_grad_() for x in xs]
    o_ref = ref(*a)
    o = net(*b)
    o_ref = o_ref if isinstance(o_ref, tuple) else (o_ref,)
    o = o if isinstance(o, tuple) else (o,)
    for x, y in zip(o, o_ref):
        _close(x, y, RTOL_OUT, "output")
    inet_loss(o_ref).backward()
    inet_loss(o).backward()
    for x, y, n in zip(b, a, ("send", "rec", "edge")):
        _close(x.grad, y.grad, RTOL_GRAD, f"grad {n}")
    for (n, p), (_, q) in zip(ref.named_parameters(), net.named_parameters()):
        _close(q.grad, p.grad, RTOL_GRAD, f"grad {n}")


def test_interaction_net_deterministic(dev):
    """No float atomics: two runs are bit-identical."""
    from neural_lam_b200.interaction_net import InteractionNet
    g = torch.Generator().manual_seed(0)
    M, n, d = 30000, 2000, 64
    ei = torch.stack((torch.randint(0, n, (M,), generator=g), torch.randint(0, n, (M,), generator=g)))
    torch.manual_seed(0)
    net = InteractionNet(ei, d).to(dev)
    x = torch.randn(2, n, d, device=dev)
    e = torch.randn(2, M, d, device=dev)
    outs = []
    for _ in range(2):
        xx, ee = x.clone().requires_grad_(), e.clone().requires_grad_()
        r, eo = net(xx, xx, ee)
        (r.sum() + eo.square().sum()).backward()
        outs.append((r.detach().clone(), xx.grad.clone(), ee.grad.clone(),
                     [p.grad.clone() for p in net.parameters()]))
        net.zero_grad()
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[0][2], outs[1][2])
    for p, q in zip(outs[0][3], outs[1][3]):
        assert torch.equal(p, q)
