"""TEST INFRASTRUCTURE ONLY.  Writes tests/golden/bf16_yardstick.pt and
tests/golden/meps_grads.pt.

1. Yard-stick for the bf16 tolerance.  BASELINE.json's north_star states 2e-2 for the bf16
   mode.  How far does the REFERENCE ITSELF move when it runs in ITS bf16 mode?  The
   reference's `--precision bf16-mixed` is Lightning autocast; here
   the unmodified reference models run on the CPU once in fp32 and once inside
   `torch.autocast("cpu", dtype=torch.bfloat16)` with the same weights and batch, and the
   deviation of loss / prediction / every parameter gradient is recorded (max-norm and L2
   relative errors).  tests/test_gpu_bf16.py bounds this repo's bf16 error by
   max(2e-2, the reference's own deviation) per quantity.
2. Full parameter gradients of the MEPS-size fp32 reference run (models.pt keeps only norms
   and strided slices of them), for an element-wise check of the full-size case.

Runs only in the build container (needs /root/reference):
    python oracle/make_golden_autocast.py
"""
import os
import sys
import tempfile

import torch
import torch._dynamo  # noqa: F401  (before the stub modules exist, see bench.py)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import make_golden, ref_stubs  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def rel_max(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def run(model_cls, cfg, ds, args, batch, state_dict, autocast):
    torch.manual_seed(42)
    model = model_cls(args, cfg, ds)
    model.load_state_dict(state_dict)
    if autocast:
        with torch.autocast("cpu", dtype=torch.bfloat16):
            loss = model.training_step(batch)
    else:
        loss = model.training_step(batch)
    loss.backward()
    with torch.no_grad():
        if autocast:
            with torch.autocast("cpu", dtype=torch.bfloat16):
                pred, _ = model.predict_step(batch[0][:, 1], batch[0][:, 0], batch[2][:, 0])
        else:
            pred, _ = model.predict_step(batch[0][:, 1], batch[0][:, 0], batch[2][:, 0])
    return (loss.detach().float(), pred.float(),
            {n: p.grad.float().clone() for n, p in model.named_parameters()})


def main():
    torch.set_num_threads(os.cpu_count())
    ref_stubs.import_reference()
    from neural_lam import config as ref_config
    from neural_lam import models as ref_models

    ref_cls = {"graph_lam": ref_models.GraphLAM, "hi_lam": ref_models.HiLAM,
               "hi_lam_parallel": ref_models.HiLAMParallel}
    cfg = ref_config.NeuralLAMConfig(
        datastore=ref_config.DatastoreSelection(kind="mdp", config_path=""))
    golden = torch.load(os.path.join(GOLDEN, "models.pt"), weights_only=False)
    cases, synthetic = make_golden.model_cases(skip_meps=False)
    yard, meps = {}, {}
    for case in cases:
        name = case["name"]
        sd = golden[name]["state_dict"]
        with tempfile.TemporaryDirectory() as root:
            ds, args, batch = make_golden.build_case(case, synthetic, root)
            l32, p32, g32 = run(ref_cls[case["model"]], cfg, ds, args, batch, sd, False)
            l16, p16, g16 = run(ref_cls[case["model"]], cfg, ds, args, batch, sd, True)
        assert torch.allclose(l32, golden[name]["loss"], rtol=1e-5), name  # same run as models.pt
        all32 = torch.cat([g32[n].reshape(-1) for n in g32])
        all16 = torch.cat([g16[n].reshape(-1) for n in g32])
        yard[name] = {
            "loss": abs((l16 - l32).item()) / abs(l32.item()),
            "pred_max": rel_max(p16, p32),
            "grad_all_l2": rel_l2(all16, all32),
            "grad_max": {n: rel_max(g16[n], g32[n]) for n in g32},
            "grad_l2": {n: rel_l2(g16[n], g32[n]) for n in g32},
        }
        worst = max(yard[name]["grad_l2"].values())
        print(f"  {name}: reference bf16-autocast vs its fp32: loss {yard[name]['loss']:.2e}, "
              f"pred {yard[name]['pred_max']:.2e}, all gradients L2 {yard[name]['grad_all_l2']:.2e}, "
              f"worst parameter L2 {worst:.2e}, worst max-norm {max(yard[name]['grad_max'].values()):.2e}")
        if case.get("summary_only"):
            meps[name] = {"param_grads": g32}
    torch.save({"how": "unmodified reference, torch.autocast('cpu', bfloat16) vs fp32, same "
                       "weights and batch as tests/golden/models.pt", "cases": yard},
               os.path.join(GOLDEN, "bf16_yardstick.pt"))
    torch.save(meps, os.path.join(GOLDEN, "meps_grads.pt"))
    print("written")


if __name__ == "__main__":
    main()
