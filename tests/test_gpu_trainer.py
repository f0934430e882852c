"""GPU: the trainer's fast paths (gradient sink into the flat buffer, CUDA-graph
replay, overlapped host feed) give the same losses / weights as plain autograd."""
import tempfile

import pytest
import torch

from helpers import build_model_case, load_golden

pytestmark = pytest.mark.gpu
CASE = load_golden("models.pt")["graphlam_multiscale_mean_d16"]


@pytest.fixture(scope="module")
def dev():
    import __graft_entry__ as entry
    entry.build()
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _model(dev):
    from neural_lam_b200 import config as nl_config
    from neural_lam_b200 import models
    with tempfile.TemporaryDirectory() as root:
        ds, args, batch = build_model_case(CASE["case"], root)
        model = models.GraphLAM(args, nl_config.default_config(), ds)
    model.load_state_dict(CASE["state_dict"])
    return model.to(dev), tuple(t.to(dev) for t in batch)


def test_grad_sink_equals_autograd(dev):
    from neural_lam_b200 import ops, train
    ops.set_param_grad_sink(False)
    model, batch = _model(dev)
    loss = model.training_step(batch)
    loss.backward()
    want = {n: p.grad.clone() for n, p in model.named_parameters()}
    model2, _ = _model(dev)
    trainer = train.DataParallelTrainer(model2)  # flat gradient buffer
    trainer.buckets.zero()
    prev = ops.scoped_trainer_flags(True, True)  # what the trainer does around its backward
    assert prev == (False, False)  # constructing a trainer leaves the process-wide flags alone
    loss2 = model2.training_step(batch)
    loss2.backward()
    from neural_lam_b200 import lib
    assert lib.load().nlam_rowmlp_bwd_pending() > 0  # queued, not yet reduced
    ops.flush_param_grads()
    assert lib.load().nlam_rowmlp_bwd_pending() == 0
    ops.scoped_trainer_flags(*prev)
    assert torch.equal(loss, loss2)
    for n, p in model2.named_parameters():
        torch.testing.assert_close(p.grad, want[n], rtol=1e-6, atol=1e-7, msg=n)
    torch.testing.assert_close(loss.cpu(), CASE["loss"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("graph", [False, True])
def test_trainer_paths_agree(dev, graph):
    from neural_lam_b200 import ops, train
    model_a, batch = _model(dev)
    model_b, _ = _model(dev)
    ta = train.DataParallelTrainer(model_a, use_cuda_graph=False)
    tb = train.DataParallelTrainer(model_b, use_cuda_graph=graph)
    la = [ta.step(batch).item() for _ in range(3)]
    host = tuple(t.cpu().pin_memory() for t in batch)
    lb = tb.fit_from_host([host] * 3)
    assert la == pytest.approx(lb, rel=1e-5)
    assert la[2] < la[0]  # it trains
    for p, q in zip(model_a.parameters(), model_b.parameters()):
        torch.testing.assert_close(p, q, rtol=1e-5, atol=1e-7)


def test_trainer_leaves_plain_autograd_intact(dev):
    """A trainer step must not leave the gradient sink / deferred reductions switched on:
    plain `loss.backward()` afterwards yields complete gradients with nothing queued."""
    from neural_lam_b200 import lib, ops, train
    model, batch = _model(dev)
    trainer = train.DataParallelTrainer(model)
    trainer.step(batch)
    assert (ops._state.get("grad_sink", False), ops._state.get("defer_reduce", False)) == (False, False)
    other, _ = _model(dev)
    ref, _ = _model(dev)
    for m in (other, ref):
        m.training_step(batch).backward()
    assert lib.load().nlam_rowmlp_bwd_pending() == 0
    for (n, p), q in zip(other.named_parameters(), ref.parameters()):
        assert p.grad is not None and torch.equal(p.grad, q.grad), n


def test_failed_backward_leaves_no_queued_reductions(dev):
    from neural_lam_b200 import lib, ops, train
    model, batch = _model(dev)
    trainer = train.DataParallelTrainer(model)

    class Boom(RuntimeError):
        pass

    def raise_boom(_grad):
        raise Boom("backward interrupted")

    # fails once the decoder / processor backward has already queued its reductions
    handle = model.g2m_gnn.register_forward_hook(
        lambda _m, _i, out: out.register_hook(raise_boom) and None)
    with pytest.raises(Boom):
        trainer.step(batch)
    handle.remove()
    assert lib.load().nlam_rowmlp_bwd_pending() == 0
    assert (ops._state.get("grad_sink", False), ops._state.get("defer_reduce", False)) == (False, False)
    trainer.step(batch)  # and the trainer still works afterwards
