"""Golden vectors of the eval path: the UNMODIFIED reference's
`validation_step` / `test_step` (ar_model.py:324-435) on the small model cases of
tests/golden/models.pt (same weights, same seeded batch), with Lightning's `log_dict`
captured.  The oracle port is checked against them here and in tests/.

    python oracle/make_golden_eval.py        # needs /root/reference; writes tests/golden/eval.pt

TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_stubs  # noqa: E402

ref_stubs.import_reference()
from oracle import make_golden, port  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def run_reference(model, batch):
    out = {}
    captured = {}
    model.log_dict = lambda d, **kw: captured.update({k: v.detach().clone() for k, v in d.items()})
    model.plot_examples = lambda *a, **k: None
    model.plotted_examples = 10**9  # no example plots
    with torch.no_grad():
        model.val_metrics["mse"].clear()
        model.validation_step(batch, 0)
        out["val_log"] = dict(captured)
        out["val_mse"] = model.val_metrics["mse"][-1].clone()
        captured.clear()
        model.test_step(batch, 0)
        out["test_log"] = dict(captured)
        out["test_mse"] = model.test_metrics["mse"][-1].clone()
        out["test_mae"] = model.test_metrics["mae"][-1].clone()
        out["spatial"] = model.spatial_loss_maps[-1].clone()
    return out


def run_port(model, batch):
    with torch.no_grad():
        vlog, vmse = model.validation_step(batch)
        tlog, entry, spatial = model.test_step(batch)
    return {"val_log": vlog, "val_mse": vmse, "test_log": tlog, "test_mse": entry["mse"],
            "test_mae": entry["mae"], "spatial": spatial}


def main():
    from neural_lam import config as ref_config
    from neural_lam import models as ref_models

    ref_cls = {"graph_lam": ref_models.GraphLAM, "hi_lam": ref_models.HiLAM,
               "hi_lam_parallel": ref_models.HiLAMParallel}
    cfg = ref_config.NeuralLAMConfig(
        datastore=ref_config.DatastoreSelection(kind="mdp", config_path=""))
    models_pt = torch.load(os.path.join(GOLDEN, "models.pt"), weights_only=False)
    cases, synthetic = make_golden.model_cases(skip_meps=True)
    out = {}
    for case in cases:
        entry = models_pt[case["name"]]
        with tempfile.TemporaryDirectory() as root:
            ds, args, batch = make_golden.build_case(case, synthetic, root)
            args.val_steps_to_log = [1, 2]  # every case unrolls >= 2 steps
            ref = ref_cls[case["model"]](args, cfg, ds)
            ref.load_state_dict(entry["state_dict"])
            mine = port.MODELS[case["model"]](args, cfg, ds)
            mine.load_state_dict(entry["state_dict"])
        r, m = run_reference(ref, batch), run_port(mine, batch)
        make_golden.check_close(r, m, f"eval/{case['name']}", rtol=1e-4, atol=1e-6)
        print(f"  eval/{case['name']}: port == reference, val_mean_loss "
              f"{r['val_log']['val_mean_loss'].item():.6f}")
        out[case["name"]] = {"case": case, "val_steps_to_log": [1, 2], "ref": r}
    torch.save(out, os.path.join(GOLDEN, "eval.pt"))
    print("wrote", os.path.join(GOLDEN, "eval.pt"))


if __name__ == "__main__":
    main()
