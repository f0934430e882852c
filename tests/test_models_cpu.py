"""CPU: model construction, parameter naming and counts of the product classes
against the golden state_dicts of the reference (no compute without a GPU)."""
import tempfile

import pytest
import torch

from helpers import build_model_case, load_golden

MODELS = load_golden("models.pt")


@pytest.mark.parametrize("name", [n for n in sorted(MODELS) if "meps" not in n])
def test_state_dict_contract(name):
    from neural_lam_b200 import config as nl_config
    from neural_lam_b200 import models
    entry = MODELS[name]
    with tempfile.TemporaryDirectory() as root:
        ds, args, _ = build_model_case(entry["case"], root)
        model = models.MODELS[entry["case"]["model"]](args, nl_config.default_config(), ds)
    sd = model.state_dict()
    assert list(sd) == list(entry["state_dict"])
    for k, v in entry["state_dict"].items():
        assert sd[k].shape == v.shape, k
    model.load_state_dict(entry["state_dict"])  # reference weights load unchanged
    opt = model.configure_optimizers()
    assert isinstance(opt, torch.optim.AdamW)
    assert opt.defaults["betas"] == (0.9, 0.95) and opt.defaults["weight_decay"] == 0.01


def test_meps_param_count():
    """GraphLAM d=64, 4 layers on MEPS dims = 214 865 params."""
    sd = MODELS["graphlam_meps_d64"]["state_dict"]
    assert sum(v.numel() for v in sd.values()) == 214865
