"""CPU: reference-layout checkpoints (incl. the legacy key rename of
ar_model.py:698-721) load into the product models."""
import os
import tempfile

import torch

from helpers import build_model_case, load_golden

ENTRY = load_golden("models.pt")["graphlam_dummy_d8"]


def _model():
    from neural_lam_b200 import config as nl_config
    from neural_lam_b200 import models
    with tempfile.TemporaryDirectory() as root:
        ds, args, _ = build_model_case(ENTRY["case"], root)
        return models.GraphLAM(args, nl_config.default_config(), ds)


def test_reference_state_dict_and_legacy_keys_load():
    from neural_lam_b200 import checkpoint
    model = _model()
    legacy = {}
    for k, v in ENTRY["state_dict"].items():  # pre-refactoring naming of the grid MLP
        legacy[k.replace("encoding_grid_mlp", "g2m_gnn.grid_mlp")] = v
    assert any(k.startswith("g2m_gnn.grid_mlp") for k in legacy)
    checkpoint.load_checkpoint(model, {"state_dict": legacy, "optimizer_states": []})
    for k, v in model.state_dict().items():
        assert torch.equal(v, ENTRY["state_dict"][k]), k


def test_round_trip_with_optimizer():
    from neural_lam_b200 import checkpoint
    model = _model()
    model.load_state_dict(ENTRY["state_dict"])
    opt = model.configure_optimizers()
    for p in model.parameters():
        p.grad = torch.ones_like(p)
    opt.step()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "last.ckpt")
        checkpoint.save_checkpoint(model, path, optimizer=opt, epoch=3)
        other = _model()
        other.restore_opt = True
        opt2 = other.configure_optimizers()
        ckpt = checkpoint.load_checkpoint(other, path, optimizer=opt2)
    assert ckpt["epoch"] == 3
    for a, b in zip(model.parameters(), other.parameters()):
        assert torch.equal(a, b)
    s1, s2 = opt.state_dict()["state"], opt2.state_dict()["state"]
    assert s1.keys() == s2.keys()
    for k in s1:
        assert torch.equal(s1[k]["exp_avg"], s2[k]["exp_avg"])
