"""Auto-regressive rollout, loss and optimizer of the training hot path
(/root/reference/neural_lam/models/ar_model.py:21-131, 191-309), as a plain
`nn.Module`: Lightning (logging, checkpoints, eval plots) is out of scope, the
training loop is neural_lam_b200.train."""
import torch
from torch import nn

from .. import metrics
from ..loss_weighting import get_state_feature_weighting


class ARModel(nn.Module):
    def __init__(self, args, config, datastore):
        super().__init__()
        self.args = args
        self._datastore = datastore
        num_state_vars = datastore.get_num_data_vars(category="state")
        num_forcing_vars = datastore.get_num_data_vars(category="forcing")
        da_static = datastore.get_dataarray(category="static", split=None)
        da_stats = datastore.get_standardization_dataarray(category="state")

        def f32(a):
            return torch.tensor(a, dtype=torch.float32)

        # (grid_index, static_feature) order, ar_model.py:51-60
        self.register_buffer(
            "grid_static_features",
            f32(da_static.transpose("grid_index", "static_feature").values),
            persistent=False)
        for name, da in (("state_mean", da_stats.state_mean), ("state_std", da_stats.state_std),
                         ("diff_mean", da_stats.state_diff_mean),
                         ("diff_std", da_stats.state_diff_std)):
            self.register_buffer(name, f32(da.values), persistent=False)

        self.feature_weights = f32(get_state_feature_weighting(config=config, datastore=datastore))
        self.output_std = bool(args.output_std)
        if self.output_std:
            self.grid_output_dim = 2 * num_state_vars
        else:
            self.grid_output_dim = num_state_vars
            # inverse of the multiplicative wMSE weighting, ar_model.py:96-103
            self.register_buffer("per_var_std", self.diff_std / torch.sqrt(self.feature_weights),
                                 persistent=False)
        self.num_grid_nodes, grid_static_dim = self.grid_static_features.shape
        self.grid_dim = (2 * self.grid_output_dim + grid_static_dim + num_forcing_vars * (
            args.num_past_forcing_steps + args.num_future_forcing_steps + 1))
        self.loss = metrics.get_metric(args.loss)

        boundary_mask = f32(datastore.boundary_mask.values).unsqueeze(1)
        self.register_buffer("boundary_mask", boundary_mask, persistent=False)
        self.register_buffer("interior_mask", 1.0 - boundary_mask, persistent=False)
        self.restore_opt = getattr(args, "restore_opt", False)
        self._num_interior = int(self.interior_mask.sum().item())

    def configure_optimizers(self):
        """AdamW(lr, betas=(0.9, 0.95)), default weight decay (ar_model.py:191-195)."""
        on_gpu = next(self.parameters()).is_cuda
        return torch.optim.AdamW(self.parameters(), lr=self.args.lr, betas=(0.9, 0.95),
                                 fused=on_gpu, capturable=on_gpu)

    @property
    def interior_mask_bool(self):
        return self.interior_mask[:, 0].to(torch.bool)

    @property
    def interior_index(self):
        """Indices of the interior grid nodes (== interior_mask_bool.nonzero()), cached:
        the metrics accept them in place of the boolean mask and then run without the
        device->host sync of boolean indexing (CUDA-graph capturable eval rollout)."""
        idx = getattr(self, "_interior_index", None)
        if idx is None or idx.device != self.interior_mask.device:
            idx = self.interior_mask_bool.nonzero().squeeze(1)
            self._interior_index = idx
        return idx

    @staticmethod
    def expand_to_batch(x, batch_size):
        """Stride-0 batch view (ar_model.py:204-209)."""
        from .. import ops
        return ops.expand_with_shadow(x, batch_size)

    def predict_step(self, prev_state, prev_prev_state, forcing):
        raise NotImplementedError("No prediction step implemented")

    class _RolloutCache:
        """Scope in which embeddings of static graph features are computed once
        (BaseGraphModel.embed_static) instead of once per AR step."""

        def __init__(self, model):
            self.model = model

        def __enter__(self):
            self.outer = getattr(self.model, "_static_cache", None)
            if self.outer is None:
                self.model._static_cache = {}

        def __exit__(self, *exc):
            if self.outer is None:
                self.model._static_cache = None
            return False

    def rollout_cache(self):
        return ARModel._RolloutCache(self)

    def unroll_prediction(self, init_states, forcing_features, true_states):
        """ar_model.py:220-267: sequential rollout; the boundary is overwritten
        with the true state before feeding back."""
        with self.rollout_cache():
            return self._unroll_prediction(init_states, forcing_features, true_states)

    def _unroll_prediction(self, init_states, forcing_features, true_states):
        prev_prev_state, prev_state = init_states[:, 0], init_states[:, 1]
        predictions, pred_stds = [], []
        for i in range(forcing_features.shape[1]):
            pred_state, pred_std = self.predict_step(prev_state, prev_prev_state,
                                                     forcing_features[:, i])
            new_state = (self.boundary_mask * true_states[:, i]
                         + self.interior_mask * pred_state)
            predictions.append(new_state)
            if self.output_std:
                pred_stds.append(pred_std)
            prev_prev_state, prev_state = prev_state, new_state
        prediction = torch.stack(predictions, dim=1)
        pred_std = torch.stack(pred_stds, dim=1) if self.output_std else self.per_var_std
        return prediction, pred_std

    def common_step(self, batch):
        """ar_model.py:269-285."""
        init_states, target_states, forcing_features, batch_times = batch
        prediction, pred_std = self.unroll_prediction(init_states, forcing_features,
                                                      target_states)
        return prediction, target_states, pred_std, batch_times

    @torch.no_grad()
    def validation_step(self, batch, batch_idx=0):
        """ar_model.py:324-361 without the Lightning logging: forward-only rollout through
        the same kernels; returns ({"val_loss_unroll{k}", "val_mean_loss"}, entry MSEs
        (B, pred_steps, d_f)) -- what the reference logs / appends to `val_metrics`."""
        prediction, target, pred_std, _ = self.common_step(batch)
        mask = self.interior_index
        time_step_loss = torch.mean(self.loss(prediction, target, pred_std, mask=mask), dim=0)
        log = {f"val_loss_unroll{step}": time_step_loss[step - 1]
               for step in self.args.val_steps_to_log if step <= len(time_step_loss)}
        log["val_mean_loss"] = torch.mean(time_step_loss)
        return log, metrics.mse(prediction, target, pred_std, mask=mask, sum_vars=False)

    @torch.no_grad()
    def test_step(self, batch, batch_idx=0):
        """ar_model.py:375-435 without logging / plotting: ({"test_loss_unroll{k}",
        "test_mean_loss"}, {"mse", "mae" [, "output_std"]: (B, pred_steps, d_f)}, spatial
        loss maps (B, len(val_steps_to_log), num_grid_nodes))."""
        prediction, target, pred_std, _ = self.common_step(batch)
        mask = self.interior_index
        time_step_loss = torch.mean(self.loss(prediction, target, pred_std, mask=mask), dim=0)
        log = {f"test_loss_unroll{step}": time_step_loss[step - 1]
               for step in self.args.val_steps_to_log}
        log["test_mean_loss"] = torch.mean(time_step_loss)
        entry = {name: metrics.get_metric(name)(prediction, target, pred_std, mask=mask,
                                                sum_vars=False) for name in ("mse", "mae")}
        if self.output_std:
            entry["output_std"] = torch.mean(pred_std.index_select(-2, mask), dim=-2)
        spatial = self.loss(prediction, target, pred_std, average_grid=False)
        return log, entry, spatial[:, [step - 1 for step in self.args.val_steps_to_log]]

    def training_step(self, batch):
        """Mean over batch and unrolled steps of the masked loss (ar_model.py:287-309)."""
        if (self.loss in (metrics.wmse, metrics.mse) and not self.output_std
                and hasattr(self, "net_output") and batch[0].is_cuda):
            return self._fused_training_step(batch)
        prediction, target, pred_std, _ = self.common_step(batch)
        if self.loss in (metrics.wmse, metrics.mse) and not self.output_std:
            return self._masked_squared_loss(prediction, target, pred_std)
        return torch.mean(self.loss(prediction, target, pred_std, mask=self.interior_mask_bool))

    def _fused_training_step(self, batch):
        """Same value as the generic path, with the state update, boundary blend and
        loss term of every AR step done by one kernel (ops.state_step)."""
        from .. import ops
        init_states, target_states, forcing_features, _ = batch
        prev_prev_state, prev_state = init_states[:, 0], init_states[:, 1]
        if not hasattr(self, "_inv_std_buf") or self._inv_std_buf.device != prev_state.device:
            self._inv_std_buf = (1.0 / self.per_var_std).contiguous()
            self._interior_flat = self.interior_mask[:, 0].contiguous()
        inv_std = self._inv_std_buf if self.loss is metrics.wmse else None
        n_steps = forcing_features.shape[1]
        total = None
        with self.rollout_cache():
            for i in range(n_steps):
                net_output = self.net_output(prev_state, prev_prev_state, forcing_features[:, i])
                new_state, loss_sum = ops.state_step(
                    net_output, prev_state, target_states[:, i], self.diff_std, self.diff_mean,
                    inv_std, self._interior_flat)
                total = loss_sum if total is None else total + loss_sum
                prev_prev_state, prev_state = prev_state, new_state
        return total / float(init_states.shape[0] * n_steps * self._num_interior)

    def _masked_squared_loss(self, prediction, target, pred_std):
        """mean_{B,T}( mean_{interior nodes}( sum_vars ((pred-target)/std)^2 ) ) -- the
        same number as metrics.wmse/mse with mask=interior_mask_bool
        (metrics.py:21-84), written as a multiply by the 0/1 interior mask so
        that no boolean-index gather (a device->host sync) is needed."""
        inv_var = 1.0 / (pred_std**2) if self.loss is metrics.wmse else None
        sq = (prediction - target) ** 2
        if inv_var is not None:
            sq = sq * inv_var
        n_bt = prediction.shape[0] * prediction.shape[1]
        return (sq * self.interior_mask).sum() / float(n_bt * self._num_interior)
