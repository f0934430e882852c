"""Checkpoint interop with the reference.

The reference saves Lightning checkpoints: a dict with `state_dict` (module
parameters only -- graph tensors are non-persistent buffers) and
`optimizer_states` (/root/reference/neural_lam/train_model.py:264-270,294-296).
`on_load_checkpoint` renames the legacy keys `g2m_gnn.grid_mlp.*` to
`encoding_grid_mlp.*` and drops the optimizer state unless `--restore_opt`
(/root/reference/neural_lam/models/ar_model.py:698-721).  Parameter names are
identical here, so the same files load into these models.
"""
import torch


def migrate_state_dict(state_dict):
    """Legacy key rename of ar_model.py:706-718 (returns a new dict)."""
    out = {}
    for key, value in state_dict.items():
        if key.startswith("g2m_gnn.grid_mlp"):
            key = key.replace("g2m_gnn.grid_mlp", "encoding_grid_mlp")
        out[key] = value
    return out


def load_checkpoint(model, path_or_dict, optimizer=None, map_location="cpu"):
    """Load a reference (Lightning) checkpoint, or a plain state_dict, into
    `model`; restores the optimizer only when one is given and the model was
    built with `restore_opt` (ar_model.py:719-721).  Returns the checkpoint dict."""
    ckpt = path_or_dict
    if not isinstance(ckpt, dict):
        ckpt = torch.load(path_or_dict, map_location=map_location, weights_only=False)
    state_dict = ckpt.get("state_dict", ckpt)
    model.load_state_dict(migrate_state_dict(state_dict))
    if optimizer is not None and getattr(model, "restore_opt", False):
        states = ckpt.get("optimizer_states")
        if states:
            optimizer.load_state_dict(states[0])
    return ckpt


def save_checkpoint(model, path, optimizer=None, epoch=0, global_step=0, **extra):
    """Write `state_dict` / `optimizer_states` under the key names of the reference's
    (Lightning) checkpoints, plus the bookkeeping keys Lightning's loader reads first
    (`pytorch-lightning_version`, `epoch`, `global_step`, `lr_schedulers`).

    Interop is one-directional by design: reference checkpoints load here
    (`load_checkpoint`), and the reference can take these weights with
    `model.load_state_dict(torch.load(path)["state_dict"])`.  A full Lightning resume
    (`trainer.fit(ckpt_path=...)`) additionally wants `loops`, `callbacks` and
    `hyper_parameters`, which only a Lightning trainer can produce; they are not written
    (the trainer and its callbacks are out of scope)."""
    ckpt = {"state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
            "pytorch-lightning_version": "2.0.0", "epoch": int(epoch),
            "global_step": int(global_step), "lr_schedulers": []}
    if optimizer is not None:
        ckpt["optimizer_states"] = [optimizer.state_dict()]
    ckpt.update(extra)
    torch.save(ckpt, path)
    return ckpt
