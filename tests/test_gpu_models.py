"""GPU parity: GraphLAM / HiLAM / HiLAMParallel train step on the CUDA path vs
golden vectors from the UNMODIFIED reference (tests/golden/models.pt).
fp32: loss and one-step prediction rtol 1e-4, parameter gradients rtol 1e-3."""
import tempfile

import pytest
import torch

from helpers import build_model_case, load_golden

pytestmark = pytest.mark.gpu
MODELS = load_golden("models.pt")


@pytest.fixture(scope="module")
def dev():
    import __graft_entry__ as entry
    entry.build()
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _close(a, b, rtol, what, atol_frac=0.1):
    a, b = a.detach().cpu(), b.detach().cpu()
    scale = b.abs().max().item() + 1e-30
    torch.testing.assert_close(a, b, rtol=rtol, atol=rtol * scale * atol_frac,
                               msg=lambda m: f"{what}: {m}")


@pytest.mark.parametrize("name", sorted(MODELS))
def test_train_step_matches_reference(dev, name):
    from neural_lam_b200 import config as nl_config
    from neural_lam_b200 import models
    entry = MODELS[name]
    case = entry["case"]
    with tempfile.TemporaryDirectory() as root:
        ds, args, batch = build_model_case(case, root)
        model = models.MODELS[case["model"]](args, nl_config.default_config(), ds)
    assert list(model.state_dict()) == list(entry["state_dict"])
    model.load_state_dict(entry["state_dict"])
    model = model.to(dev)
    batch = tuple(t.to(dev) for t in batch)
    loss = model.training_step(batch)
    _close(loss, entry["loss"], 1e-4, f"{name} loss")
    loss.backward()
    with torch.no_grad():
        pred, _ = model.predict_step(batch[0][:, 1], batch[0][:, 0], batch[2][:, 0])
    if case.get("summary_only"):
        _close(pred[:, ::997], entry["pred_slice"], 1e-4, f"{name} pred")
        for n, p in model.named_parameters():
            _close(p.grad.norm(), entry["grad_norms"][n], 1e-3, f"{name} |grad {n}|")
            _close(p.grad.reshape(-1)[::53], entry["grad_slices"][n], 1e-3, f"{name} grad {n}")
    else:
        _close(pred, entry["pred_step"], 1e-4, f"{name} pred")
        for n, p in model.named_parameters():
            _close(p.grad, entry["param_grads"][n], 1e-3, f"{name} grad {n}")


EVAL = load_golden("eval.pt")


@pytest.mark.parametrize("precision,rtol", [("fp32", 1e-4), ("bf16", 2e-2)])
@pytest.mark.parametrize("name", sorted(EVAL))
def test_eval_path_matches_reference(dev, name, precision, rtol):
    """validation_step / test_step (forward-only rollout through the CUDA kernels) against
    the unmodified reference's logged values and metric entries (tests/golden/eval.pt)."""
    from neural_lam_b200 import config as nl_config
    from neural_lam_b200 import models, ops
    entry = EVAL[name]
    case, ref = entry["case"], entry["ref"]
    with tempfile.TemporaryDirectory() as root:
        ds, args, batch = build_model_case(case, root)
        args.val_steps_to_log = entry["val_steps_to_log"]
        model = models.MODELS[case["model"]](args, nl_config.default_config(), ds)
    model.load_state_dict(MODELS[name]["state_dict"])
    model = model.to(dev)
    batch = tuple(t.to(dev) for t in batch)
    ops.set_precision(precision)
    try:
        vlog, vmse = model.validation_step(batch)
        tlog, tentry, spatial = model.test_step(batch)
    finally:
        ops.set_precision("fp32")
    for k, v in ref["val_log"].items():
        _close(vlog[k], v, rtol, f"{name} {k}")
    for k, v in ref["test_log"].items():
        _close(tlog[k], v, rtol, f"{name} {k}")
    _close(vmse, ref["val_mse"], rtol, "val entry mse")
    _close(tentry["mse"], ref["test_mse"], rtol, "test entry mse")
    _close(tentry["mae"], ref["test_mae"], rtol, "test entry mae")
    _close(spatial, ref["spatial"], rtol, "spatial loss maps")
    assert not any(p.grad is not None for p in model.parameters())  # forward only
