"""Synthetic, seeded stand-ins for the reference's data layer (which is out of
scope: xarray/zarr I/O on CPU workers).

`SyntheticDatastore` is duck-typed to exactly what the model constructors read
(/root/reference/neural_lam/models/ar_model.py:40-48, loss_weighting.py:69,
models/base_graph_model.py:24, create_graph.py:546): `root_path`,
`get_num_data_vars`, `get_vars_names`, `get_dataarray("static")`,
`get_standardization_dataarray("state")`, `boundary_mask`, `get_xy`.

Shapes follow the two fixtures the reference defines:
* "dummy"  = tests/dummy_datastore.py:22-155: 100 x 100 grid over 500 km,
  5 state / 2 forcing / 1 static features, stats all ones, random 0/1 mask;
* "meps"   = README.md:277-351 + datastore/npyfilesmeps/store.py:124-131,
  586-595: 238 x 268 grid at 2.5 km, 17 state / 6 forcing / 4 static
  features, boundary = outer 10 cells.
`synthetic_batch` returns the 4-tuple `WeatherDataset.__getitem__` collates to
(weather_dataset.py:443-496): init_states (B,2,N,d_f), target_states
(B,T,N,d_f), forcing (B,T,N,d_forcing*window), times (B,T) int64.
"""
import pathlib
import types

import numpy as np
import torch


class _DA:
    """Minimal DataArray look-alike: `.values` and a no-op `.transpose`."""

    def __init__(self, values):
        self.values = values

    def transpose(self, *dims):
        return self


class SyntheticDatastore:
    def __init__(self, root_path, nx, ny, n_state, n_forcing, n_static,
                 border=None, spacing=2500.0, random_stats=False, seed=0):
        self.root_path = pathlib.Path(root_path)
        self.nx, self.ny = nx, ny
        self.n = dict(state=n_state, forcing=n_forcing, static=n_static)
        rng = np.random.default_rng(seed)
        self.num_grid_points = nx * ny
        self._static = rng.standard_normal((nx * ny, n_static))
        if border is None:  # dummy: random 0/1 mask
            mask = rng.integers(0, 2, size=(nx, ny))
        else:
            mask = np.ones((nx, ny), dtype=np.int64)
            mask[border:-border, border:-border] = 0
        # grid_index stacks ("x", "y"), x-major (datastore/base.py:428,529)
        self._mask = mask.reshape(-1)
        if random_stats:
            stats = {k: rng.uniform(0.5, 1.5, n_state) for k in
                     ("state_mean", "state_std", "state_diff_mean", "state_diff_std")}
        else:
            stats = {k: np.ones(n_state) for k in
                     ("state_mean", "state_std", "state_diff_mean", "state_diff_std")}
        self._stats = types.SimpleNamespace(**{k: _DA(v) for k, v in stats.items()})
        x = spacing * np.arange(nx)
        y = spacing * np.arange(ny)
        self._xy = np.zeros((nx, ny, 2))
        self._xy[:, :, 0] = x[:, None]
        self._xy[:, :, 1] = y[None, :]

    # ---- the attributes the models use
    def get_num_data_vars(self, category):
        return self.n[category]

    def get_vars_names(self, category):
        return [f"{category}_feat_{i}" for i in range(self.n[category])]

    def get_dataarray(self, category, split=None):
        assert category == "static"
        return _DA(self._static)

    def get_standardization_dataarray(self, category):
        assert category == "state"
        return self._stats

    @property
    def boundary_mask(self):
        return _DA(self._mask)

    def get_xy(self, category="state", stacked=False):
        return self._xy.reshape(-1, 2) if stacked else self._xy


def dummy_datastore(root_path, n_1d=100, seed=0):
    """Shapes of tests/dummy_datastore.py (5/2/1 features, stats = 1)."""
    return SyntheticDatastore(root_path, n_1d, n_1d, 5, 2, 1, border=None,
                              spacing=500e3 / n_1d, seed=seed)


def meps_datastore(root_path, scale=1, seed=0, random_stats=True):
    """MEPS-shaped domain (238 x 268, 17/6/4 features, border 10);
    `scale=2` is the 4x-area domain 476 x 536 of BASELINE config 5."""
    return SyntheticDatastore(root_path, 238 * scale, 268 * scale, 17, 6, 4,
                              border=10, seed=seed, random_stats=random_stats)


def synthetic_batch(datastore, batch_size, ar_steps, num_past_forcing_steps=1,
                    num_future_forcing_steps=1, seed=0, device="cpu",
                    pin_memory=False):
    g = torch.Generator().manual_seed(seed)
    n = datastore.num_grid_points
    d_f = datastore.get_num_data_vars("state")
    d_forc = datastore.get_num_data_vars("forcing") * (
        num_past_forcing_steps + num_future_forcing_steps + 1)
    init_states = torch.randn(batch_size, 2, n, d_f, generator=g)
    target_states = torch.randn(batch_size, ar_steps, n, d_f, generator=g)
    forcing = torch.randn(batch_size, ar_steps, n, d_forc, generator=g)
    times = torch.arange(batch_size * ar_steps, dtype=torch.int64).view(
        batch_size, ar_steps)
    batch = (init_states, target_states, forcing, times)
    if pin_memory:
        batch = tuple(t.pin_memory() for t in batch)
    if str(device) != "cpu":
        batch = tuple(t.to(device) for t in batch)
    return batch


def synthetic_series(datastore, n_time, seed=0, pin_memory=False):
    """Raw (un-standardised) analysis time series of the datastore's shape for the device
    feed: state (T, N, d_state), forcing (T, N, d_forcing), times (T,) int64 ns, and the
    standardisation statistics (state_mean, state_std, forcing_mean, forcing_std)."""
    g = torch.Generator().manual_seed(seed)
    n = datastore.num_grid_points
    d_s = datastore.get_num_data_vars("state")
    d_f = datastore.get_num_data_vars("forcing")
    stats = (torch.randn(d_s, generator=g), torch.rand(d_s, generator=g) + 0.5,
             torch.randn(d_f, generator=g), torch.rand(d_f, generator=g) + 0.5)
    state = torch.randn(n_time, n, d_s, generator=g) * stats[1] + stats[0]
    forcing = torch.randn(n_time, n, d_f, generator=g) * stats[3] + stats[2]
    times = torch.arange(n_time, dtype=torch.int64) * (3 * 3600 * 10 ** 9)
    if pin_memory:
        state, forcing, times = state.pin_memory(), forcing.pin_memory(), times.pin_memory()
    return state, forcing, times, stats


class ModelArgs:
    """Duck-typed `args` with the reference CLI defaults
    (train_model.py:29-209; tests/test_training.py:71-87)."""

    def __init__(self, **kw):
        self.graph = "multiscale"
        self.hidden_dim = 64
        self.hidden_layers = 1
        self.processor_layers = 4
        self.mesh_aggr = "sum"
        self.output_std = False
        self.loss = "wmse"
        self.lr = 1.0e-3
        self.restore_opt = False
        self.n_example_pred = 1
        self.val_steps_to_log = [1, 2, 3]
        self.metrics_watch = []
        self.num_past_forcing_steps = 1
        self.num_future_forcing_steps = 1
        self.precision = 32
        for k, v in kw.items():
            setattr(self, k, v)
