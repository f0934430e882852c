"""CPU: the restatement of WeatherDataset's sample construction (oracle/port_dataset.py)
against hand-computed expectations (the reference file needs xarray and its own tests only
check shapes -- tests/test_datasets.py)."""
import torch

import helpers  # noqa: F401
from oracle.port_dataset import WeatherDatasetPort


def _series(T=12, N=3, ds=2, df=2):
    # value encodes (time, node, feature): t*100 + n*10 + f
    t = torch.arange(T).view(T, 1, 1) * 100.0
    n = torch.arange(N).view(1, N, 1) * 10.0
    state = t + n + torch.arange(ds).view(1, 1, ds)
    forcing = 1000 + t + n + torch.arange(df).view(1, 1, df)
    return state, forcing, torch.arange(T) * 3600


def test_item_layout_default_window():
    state, forcing, times = _series()
    ds = WeatherDatasetPort(state, forcing, times, 0.0, 1.0, 0.0, 1.0, ar_steps=3,
                            num_past_forcing_steps=1, num_future_forcing_steps=1)
    assert len(ds) == 12 - 3 - 2 - 1
    init, target, fw, tt = ds[2]
    assert init.shape == (2, 3, 2) and target.shape == (3, 3, 2) and fw.shape == (3, 3, 6)
    assert init[:, 0, 0].tolist() == [200.0, 300.0]
    assert target[:, 0, 0].tolist() == [400.0, 500.0, 600.0]
    assert tt.tolist() == [4 * 3600, 5 * 3600, 6 * 3600]
    # target step s = time 4 + s; window = times (3+s, 4+s, 5+s); feature-major stacking
    assert fw[0, 1].tolist() == [1310.0, 1410.0, 1510.0, 1311.0, 1411.0, 1511.0]
    assert fw[2, 2].tolist() == [1520.0, 1620.0, 1720.0, 1521.0, 1621.0, 1721.0]


def test_item_layout_long_past_window_and_standardisation():
    state, forcing, times = _series(T=14)
    sm, ss = torch.tensor([1.0, 2.0]), torch.tensor([2.0, 4.0])
    ds = WeatherDatasetPort(state, forcing, times, sm, ss, torch.zeros(2), torch.ones(2),
                            ar_steps=2, num_past_forcing_steps=3, num_future_forcing_steps=0)
    assert len(ds) == 14 - 2 - 3 - 0
    init, target, fw, tt = ds[1]
    # past = 3 > 2: states start at idx + 1 (weather_dataset.py:219), targets at idx + 3
    assert torch.allclose(init[:, 0, 0], (torch.tensor([200.0, 300.0]) - 1) / 2)
    assert torch.allclose(target[:, 1, 1], (torch.tensor([411.0, 511.0]) - 2) / 4)
    assert tt.tolist() == [4 * 3600, 5 * 3600]
    # window of step 0 = times 1..4 (offset 4, past 3), 4 entries per feature
    assert fw.shape == (2, 3, 8)
    assert fw[0, 0, :4].tolist() == [1100.0, 1200.0, 1300.0, 1400.0]
