"""Device-side data feed: the GPU counterpart of
/root/reference/neural_lam/weather_dataset.py:13-496 (`WeatherDataset` for analysis data).

The reference builds every sample on CPU workers with xarray (time slicing :163-222, forcing
windowing :224-330, standardisation :399-420, tensor conversion :466-496) and ships whole
batches through pinned host memory (70 MB per step of 4 MEPS samples).  Here the time series
is standardised ONCE on the GPU when it is uploaded and stays resident in HBM (one MEPS time
step is 5.9 MB; a year of hourly data is 51 GB of the 180 GB); a batch is assembled from it
by one kernel (`nlam_feed_batch_run`).  Per training step the host sends only the sample
indices -- or, when streaming, the newest time slice (`append`).

Same item semantics as the reference: `feed.batch([i, j, ...])` equals the DataLoader's
default collation of `dataset[i], dataset[j], ...` (init_states (B, 2, N, d), target_states
(B, ar_steps, N, d), forcing (B, ar_steps, N, d_forcing * window), target_times (B, ar_steps)).
Forecast / ensemble datastores (weather_dataset.py:116-143) are not covered.
"""
import ctypes

import torch

from . import lib as L


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class DeviceWeatherFeed:
    def __init__(self, n_grid, d_state, d_forcing, state_mean=None, state_std=None,
                 forcing_mean=None, forcing_std=None, ar_steps=3, num_past_forcing_steps=1,
                 num_future_forcing_steps=1, standardize=True, capacity=1024, ring=False,
                 device="cuda"):
        """capacity: time steps kept on the device.  ring=False: `append` fails when full
        (a fixed dataset); ring=True: the oldest steps are overwritten (streaming)."""
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceWeatherFeed lives on a CUDA device (there is no CPU fallback; "
                               "the CPU path is the reference's WeatherDataset)")
        self.n_grid, self.d_state, self.d_forcing = int(n_grid), int(d_state), int(d_forcing)
        self.ar_steps = int(ar_steps)
        self.past, self.future = int(num_past_forcing_steps), int(num_future_forcing_steps)
        self.standardize = bool(standardize)
        self.capacity, self.ring = int(capacity), bool(ring)
        f32 = dict(device=self.device, dtype=torch.float32)
        self.state = torch.empty((self.capacity, self.n_grid, self.d_state), **f32)
        self.forcing = (torch.empty((self.capacity, self.n_grid, self.d_forcing), **f32)
                        if self.d_forcing > 0 else None)
        self.times = torch.zeros((self.capacity,), device=self.device, dtype=torch.int64)
        self.t_lo, self.t_hi = 0, 0
        if self.standardize:
            mk = lambda v, d, fill: (torch.as_tensor(v, dtype=torch.float32).reshape(-1).to(self.device)
                                     if v is not None else torch.full((d,), fill, **f32))
            self.state_mean, self.state_std = mk(state_mean, d_state, 0.0), mk(state_std, d_state, 1.0)
            if self.d_forcing > 0:
                self.forcing_mean = mk(forcing_mean, d_forcing, 0.0)
                self.forcing_std = mk(forcing_std, d_forcing, 1.0)
        self._stage = {}

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_arrays(cls, state, forcing=None, times=None, chunk=64, **kw):
        """state (T, N, d_state), forcing (T, N, d_forcing) or None, times (T,) int64 -- host
        (or device) tensors of the raw, un-standardised series; uploaded in chunks."""
        T, N, ds = state.shape
        df = forcing.shape[2] if forcing is not None else 0
        kw.setdefault("capacity", T)
        feed = cls(N, ds, df, **kw)
        for t0 in range(0, T, chunk):
            t1 = min(T, t0 + chunk)
            feed.append(state[t0:t1], forcing[t0:t1] if forcing is not None else None,
                        times[t0:t1] if times is not None else None)
        return feed

    def __len__(self):
        """weather_dataset.py:144-160 (analysis data)."""
        return (self.t_hi - self.t_lo) - self.ar_steps - max(2, self.past) - self.future

    # ------------------------------------------------------------------ upload
    def append(self, state, forcing=None, times=None, stream=None):
        """Add k new time steps (k, N, d_state) [+ (k, N, d_forcing), (k,)]: host tensors
        (pinned for an asynchronous copy) or device tensors.  Standardised on the device.
        With `stream` the copy + standardisation run there (overlapping the training step on
        the current stream); the caller orders later `batch` calls after it."""
        k = state.shape[0]
        if state.shape[1:] != (self.n_grid, self.d_state):
            raise ValueError(f"append: state {tuple(state.shape)} != (k, {self.n_grid}, {self.d_state})")
        if (forcing is None) != (self.d_forcing == 0):
            raise ValueError("append: forcing must be given exactly when d_forcing > 0")
        if not self.ring and self.t_hi + k > self.capacity:
            raise RuntimeError(f"DeviceWeatherFeed is full ({self.capacity} time steps); "
                               "create it with a larger capacity or ring=True")
        if k > self.capacity:
            raise ValueError("append: more time steps than the capacity")
        lib = L.load()
        ctx = torch.cuda.stream(stream) if stream is not None else _NullCtx()
        with ctx:
            st = _stream()
            done = 0
            while done < k:  # a ring write may wrap around
                slot = (self.t_hi + done) % self.capacity
                n = min(k - done, self.capacity - slot)
                self._put(lib, state[done:done + n], self.state[slot:slot + n], "state", st)
                if forcing is not None:
                    self._put(lib, forcing[done:done + n], self.forcing[slot:slot + n], "forcing", st)
                if times is not None:
                    self.times[slot:slot + n].copy_(torch.as_tensor(times[done:done + n]),
                                                    non_blocking=True)
                done += n
        self.t_hi += k
        self.t_lo = max(self.t_lo, self.t_hi - self.capacity)

    def _put(self, lib, src, dst, which, st):
        src = torch.as_tensor(src, dtype=torch.float32)
        dst.copy_(src, non_blocking=True)  # H2D (or D2D) into the resident series
        if self.standardize:
            mean = self.state_mean if which == "state" else self.forcing_mean
            std = self.state_std if which == "state" else self.forcing_std
            d = dst.shape[-1]
            L.check(lib.nlam_feed_standardize(dst.data_ptr(), mean.data_ptr(), std.data_ptr(),
                                              dst.data_ptr(), dst.numel() // d, d, st),
                    "nlam_feed_standardize")

    # ------------------------------------------------------------------ batches
    def batch(self, indices, out=None):
        """Sample indices (host ints, like a DataLoader sampler yields) -> (init_states,
        target_states, forcing, target_times) on the device, assembled by one kernel.
        out: optional tuple of preallocated tensors to fill (CUDA-graph static batch)."""
        idx = [int(i) for i in indices]
        B = len(idx)
        if not 1 <= B <= L.FEED_MAX_BATCH:
            raise ValueError(f"batch: 1..{L.FEED_MAX_BATCH} samples per call")
        W = self.past + self.future + 1
        dev = self.device
        if out is None:
            out = (torch.empty((B, 2, self.n_grid, self.d_state), device=dev),
                   torch.empty((B, self.ar_steps, self.n_grid, self.d_state), device=dev),
                   torch.empty((B, self.ar_steps, self.n_grid, self.d_forcing * W), device=dev),
                   torch.empty((B, self.ar_steps), device=dev, dtype=torch.int64))
        init, target, forc, times = out
        d = L.FeedBatch()
        d.state = self.state.data_ptr()
        d.forcing = self.forcing.data_ptr() if self.forcing is not None else None
        d.times = self.times.data_ptr()
        for b, i in enumerate(idx):
            d.sample_idx[b] = i
        d.batch, d.n_grid, d.d_state, d.d_forcing = B, self.n_grid, self.d_state, self.d_forcing
        d.ar_steps, d.past, d.future = self.ar_steps, self.past, self.future
        d.ring_cap = self.capacity if self.ring else 0
        d.t_lo, d.t_hi = self.t_lo, self.t_hi
        d.init_states, d.target_states = init.data_ptr(), target.data_ptr()
        d.forcing_out = forc.data_ptr() if self.forcing is not None else None
        d.target_times = times.data_ptr()
        L.check(L.load().nlam_feed_batch_run(ctypes.byref(d), _stream()), "nlam_feed_batch_run")
        return out

    def bytes_per_time_step(self):
        return 4 * self.n_grid * (self.d_state + self.d_forcing) + 8


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
