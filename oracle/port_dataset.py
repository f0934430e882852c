"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement (plain torch indexing) of the sample construction of
/root/reference/neural_lam/weather_dataset.py for ANALYSIS data:
  __len__                    :144-160
  _slice_state_time          :206-222 (else-branch)
  _slice_forcing_time        :286-328 (else-branch)
  _build_item_dataarrays     :385-440 (standardisation :399-420, feature/window stacking
                              :417-420 -- `stack(forcing_feature_windowed=("forcing_feature",
                              "window"))` puts the feature index outermost)
  __getitem__                :466-496
The reference file itself cannot run here (it needs xarray, absent and not installable), and
its tests (tests/test_datasets.py) only check shapes, so this restatement is NOT pinned
against reference outputs ("parity unpinned" for the data feed, stated in DESIGN.md); it is
pinned against hand-computed expectations in tests/test_feed_cpu.py.
"""
import torch


class WeatherDatasetPort:
    def __init__(self, state, forcing, times, state_mean, state_std, forcing_mean, forcing_std,
                 ar_steps=3, num_past_forcing_steps=1, num_future_forcing_steps=1,
                 standardize=True):
        self.state, self.forcing, self.times = state, forcing, times
        self.sm, self.ss, self.fm, self.fs = state_mean, state_std, forcing_mean, forcing_std
        self.ar_steps = ar_steps
        self.past, self.future = num_past_forcing_steps, num_future_forcing_steps
        self.standardize = standardize

    def __len__(self):  # :144-160
        return self.state.shape[0] - self.ar_steps - max(2, self.past) - self.future

    def __getitem__(self, idx):
        init_steps = 2
        # _slice_state_time, analysis branch (:216-222)
        start = idx + max(0, self.past - init_steps)
        end = idx + max(init_steps, self.past) + self.ar_steps
        st = self.state[start:end]
        # _slice_forcing_time, analysis branch (:289-326)
        offset = idx + max(init_steps, self.past)
        windows = []
        for step in range(self.ar_steps):
            s0 = offset + step - self.past
            s1 = offset + step + self.future
            windows.append(self.forcing[s0:s1 + 1])  # (window, N, f)
        fw = torch.stack(windows)  # (time, window, N, f)
        init, target = st[:2], st[2:]
        times = self.times[start + 2:end]
        if self.standardize:  # :399-420
            init = (init - self.sm) / self.ss
            target = (target - self.sm) / self.ss
            fw = (fw - self.fm) / self.fs
        # stack ("forcing_feature", "window") -> feature-major (:417-420)
        fw = fw.permute(0, 2, 3, 1).reshape(fw.shape[0], fw.shape[2], -1)
        return init.float(), target.float(), fw.float(), times
