"""Losses / metrics with the reference's names, signatures and reductions
(/root/reference/neural_lam/metrics.py:5-237).  `training_step` only uses
`wmse` / `mse` with a grid mask (ar_model.py:294-298)."""
import torch


def mask_and_reduce_metric(metric_entry_vals, mask, average_grid, sum_vars):
    """metrics.py:21-53: keep masked grid nodes, mean over grid (dim -2), then
    sum over variables (dim -1)."""
    if mask is not None:
        if mask.dtype == torch.bool:
            metric_entry_vals = metric_entry_vals[..., mask, :]
        else:  # int64 node indices (== mask.nonzero()): same rows, no device->host sync
            metric_entry_vals = metric_entry_vals.index_select(-2, mask)
    if average_grid:
        metric_entry_vals = torch.mean(metric_entry_vals, dim=-2)
    if sum_vars:
        metric_entry_vals = torch.sum(metric_entry_vals, dim=-1)
    return metric_entry_vals


def wmse(pred, target, pred_std, mask=None, average_grid=True, sum_vars=True):
    """Weighted MSE: ((pred-target)/pred_std)^2 (metrics.py:56-84)."""
    entry = (pred - target) ** 2 / (pred_std**2)
    return mask_and_reduce_metric(entry, mask, average_grid, sum_vars)


def mse(pred, target, pred_std, mask=None, average_grid=True, sum_vars=True):
    """metrics.py:87-113."""
    return wmse(pred, target, torch.ones_like(pred_std), mask, average_grid, sum_vars)


def wmae(pred, target, pred_std, mask=None, average_grid=True, sum_vars=True):
    """metrics.py:116-144."""
    entry = torch.abs(pred - target) / pred_std
    return mask_and_reduce_metric(entry, mask, average_grid, sum_vars)


def mae(pred, target, pred_std, mask=None, average_grid=True, sum_vars=True):
    """metrics.py:147-173."""
    return wmae(pred, target, torch.ones_like(pred_std), mask, average_grid, sum_vars)


def nll(pred, target, pred_std, mask=None, average_grid=True, sum_vars=True):
    """Gaussian negative log-likelihood (metrics.py:176-201)."""
    dist = torch.distributions.Normal(pred, pred_std)
    return mask_and_reduce_metric(-dist.log_prob(target), mask, average_grid, sum_vars)


def crps_gauss(pred, target, pred_std, mask=None, average_grid=True, sum_vars=True):
    """Closed-form Gaussian CRPS (metrics.py:204-227)."""
    std_normal = torch.distributions.Normal(
        torch.zeros((), device=pred.device), torch.ones((), device=pred.device))
    diff = (target - pred) / pred_std
    entry = -pred_std * (
        torch.pi ** (-0.5)
        - 2 * torch.exp(std_normal.log_prob(diff))
        - diff * (2 * std_normal.cdf(diff) - 1))
    return mask_and_reduce_metric(entry, mask, average_grid, sum_vars)


DEFINED_METRICS = {"mse": mse, "mae": mae, "wmse": wmse, "wmae": wmae, "nll": nll,
                   "crps_gauss": crps_gauss}


def get_metric(metric_name):
    """metrics.py:5-18."""
    name = metric_name.lower()
    assert name in DEFINED_METRICS, f"Unknown metric: {metric_name}"
    return DEFINED_METRICS[name]
