"""bf16-mode parity report (GPU): this repo's deviation from the reference's fp32 goldens next
to the reference's OWN bf16-autocast deviation (tests/golden/bf16_yardstick.pt), per quantity.
    python tools/parity_report.py > profiles/parity_bf16_r2.txt
With `fp32` as argument: the fp32 mode (split-operand tensor-core kernels where they apply, then
again with option fp32_split = 0 = FFMA kernels) against the same goldens.
    python tools/parity_report.py fp32 > profiles/parity_fp32_r2.txt"""
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402

entry.build()
from helpers import build_model_case, load_golden  # noqa: E402
from neural_lam_b200 import config as nl_config  # noqa: E402
from neural_lam_b200 import models, ops  # noqa: E402


def rel_max(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def fp32_report():
    from neural_lam_b200 import lib
    dev = torch.device("cuda:0")
    MODELS = load_golden("models.pt")
    meps = load_golden("meps_grads.pt")
    ops.set_precision("fp32")
    for split in (1, 0):
        lib.load().nlam_set_option(b"fp32_split", split)
        print(f"#### fp32 mode, option fp32_split = {split} "
              f"({'tcgen05, split bf16 operands where the tiles fit' if split else 'FFMA kernels'})")
        for name, entry_ in MODELS.items():
            case = entry_["case"]
            with tempfile.TemporaryDirectory() as root:
                ds, args, batch = build_model_case(case, root)
                model = models.MODELS[case["model"]](args, nl_config.default_config(), ds)
            model.load_state_dict(entry_["state_dict"])
            model = model.to(dev)
            batch = tuple(t.to(dev) for t in batch)
            loss = model.training_step(batch)
            loss.backward()
            want = entry_.get("param_grads") or meps[name]["param_grads"]
            got = {n: p.grad.detach().float().cpu() for n, p in model.named_parameters()}
            d = int(getattr(args, "hidden_dim", 0))
            fam = ops.kernel_family((d, d, d), d, d, "fp32")
            all_g = torch.cat([got[n].reshape(-1) for n in want])
            all_w = torch.cat([want[n].reshape(-1) for n in want])
            worst = max((rel_max(got[n], want[n]), n) for n in want)
            print(f"== {name} (d={d}, edge-MLP kernel family {fam}): loss "
                  f"{abs(loss.item() - entry_['loss'].item()) / abs(entry_['loss'].item()):.2e};  "
                  f"all gradients L2 {rel_l2(all_g, all_w):.2e};  worst parameter max-norm "
                  f"{worst[0]:.2e} ({worst[1]});  bound 1e-3")
    lib.load().nlam_set_option(b"fp32_split", 1)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "fp32":
        return fp32_report()
    dev = torch.device("cuda:0")
    MODELS = load_golden("models.pt")
    yard = load_golden("bf16_yardstick.pt")["cases"]
    meps = load_golden("meps_grads.pt")
    ops.set_precision("bf16")
    for name, entry_ in MODELS.items():
        case = entry_["case"]
        with tempfile.TemporaryDirectory() as root:
            ds, args, batch = build_model_case(case, root)
            model = models.MODELS[case["model"]](args, nl_config.default_config(), ds)
        model.load_state_dict(entry_["state_dict"])
        model = model.to(dev)
        batch = tuple(t.to(dev) for t in batch)
        loss = model.training_step(batch)
        loss.backward()
        want = entry_.get("param_grads") or meps[name]["param_grads"]
        got = {n: p.grad.detach().float().cpu() for n, p in model.named_parameters()}
        y = yard[name]
        all_g = torch.cat([got[n].reshape(-1) for n in want])
        all_w = torch.cat([want[n].reshape(-1) for n in want])
        print(f"== {name}: loss {abs(loss.item() - entry_['loss'].item()) / abs(entry_['loss'].item()):.2e} "
              f"(reference autocast {y['loss']:.2e});  all gradients L2 {rel_l2(all_g, all_w):.2e} "
              f"(reference autocast {y['grad_all_l2']:.2e})")
        worst = []
        for n in want:
            e2, em = rel_l2(got[n], want[n]), rel_max(got[n], want[n])
            worst.append((e2, em, n))
        worst.sort(reverse=True)
        n_over = sum(1 for e2, em, n in worst if em > 2e-2)
        n_over_y = sum(1 for e2, em, n in worst if em > max(2e-2, y["grad_max"][n]))
        print(f"   parameters with max-norm error > 2e-2: {n_over} of {len(worst)}; "
              f"> max(2e-2, reference autocast): {n_over_y}")
        for e2, em, n in worst[:8]:
            print(f"   {n:48s} L2 {e2:.2e} max {em:.2e} | reference autocast L2 "
                  f"{y['grad_l2'][n]:.2e} max {y['grad_max'][n]:.2e}")


if __name__ == "__main__":
    main()
