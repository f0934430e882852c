"""GPU parity at BASELINE.json's full sizes (configs 3-5, config 4 with its ar_steps = 3): the CUDA
path (fp32 mode) against the CPU oracle run in the same process on the same seeded
inputs; bf16 mode checked on the loss and the whole gradient (2e-2).  One sample per case
keeps the oracle to seconds (config 5 with 2 of its processor layers)."""
import tempfile

import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = {
    # config 3: GraphCast-LAM multiscale mesh, hidden_dim 128, 8 processor layers
    "multiscale_d128": dict(model="graph_lam", scale=1, graph=dict(n_max_levels=None, hierarchical=False),
                            args=dict(hidden_dim=128, processor_layers=8, graph="multiscale"), ar=1),
    # config 4: HiLAM hierarchical 4-level mesh, hidden_dim 64, AR rollout
    "hilam_d64_ar3": dict(model="hi_lam", scale=1, graph=dict(n_max_levels=None, hierarchical=True),
                          args=dict(hidden_dim=64, processor_layers=4, graph="hierarchical"), ar=3),
    # config 5: HiLAMParallel on the 4x domain (536 x 476 grid), hidden_dim 128
    "hilam_parallel_4x_d128": dict(model="hi_lam_parallel", scale=2,
                                   graph=dict(n_max_levels=None, hierarchical=True),
                                   args=dict(hidden_dim=128, processor_layers=2, graph="hierarchical"),
                                   ar=1),
}


@pytest.fixture(scope="module")
def dev():
    import __graft_entry__ as entry
    entry.build()
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).norm().item() / (b.norm().item() + 1e-30)


@pytest.mark.parametrize("name", sorted(CASES))
def test_full_size_train_step_vs_oracle(dev, name):
    import os
    from neural_lam_b200 import config as nl_config
    from neural_lam_b200 import create_graph, models, ops, synthetic
    from oracle import port
    case = CASES[name]
    torch.set_num_threads(os.cpu_count())
    with tempfile.TemporaryDirectory() as root:
        ds = synthetic.meps_datastore(root, scale=case["scale"], seed=5)
        args = synthetic.ModelArgs(loss="wmse", **case["args"])
        create_graph.create_graph(os.path.join(root, "graph", args.graph),
                                  ds.get_xy("state", stacked=False), **case["graph"])
        torch.manual_seed(11)
        ref = port.MODELS[case["model"]](args, None, ds)
        model = models.MODELS[case["model"]](args, nl_config.default_config(), ds)
    model.load_state_dict(ref.state_dict())
    model = model.to(dev)
    batch = synthetic.synthetic_batch(ds, 1, case["ar"], seed=23)
    loss_ref = ref.training_step(batch)
    loss_ref.backward()
    gbatch = tuple(t.to(dev) for t in batch)

    ops.set_precision("fp32")
    loss = model.training_step(gbatch)
    loss.backward()
    torch.testing.assert_close(loss.cpu(), loss_ref.detach(), rtol=1e-4, atol=1e-6)
    worst = max((_rel_l2(q.grad, p.grad), n) for (n, p), (_, q) in
                zip(ref.named_parameters(), model.named_parameters()))
    assert worst[0] <= 1e-3, f"fp32 gradient {worst[1]}: relative L2 error {worst[0]:.3e}"

    ops.set_precision("bf16")
    try:
        model.zero_grad()
        loss16 = model.training_step(gbatch)
        loss16.backward()
        assert abs(loss16.item() - loss_ref.item()) <= 2e-2 * abs(loss_ref.item())
        got = torch.cat([q.grad.reshape(-1) for _, q in model.named_parameters()])
        want = torch.cat([p.grad.reshape(-1) for _, p in ref.named_parameters()])
        # whole gradient within the stated 2e-2 (measured 5e-3 .. 9e-3, profiles/parity_bf16_r2.txt)
        assert _rel_l2(got, want) <= 2e-2, f"bf16 whole-gradient relative L2 {_rel_l2(got, want):.3e}"
    finally:
        ops.set_precision("fp32")
