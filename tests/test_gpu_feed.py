"""GPU: DeviceWeatherFeed (device-resident series, one-kernel batch assembly) equals the CPU
restatement of the reference's WeatherDataset item construction + default collation,
bit for bit (it is pure data movement plus one fp32 (x - mean) / std)."""
import pytest
import torch

import helpers  # noqa: F401
from oracle.port_dataset import WeatherDatasetPort

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    import __graft_entry__ as entry
    entry.build()
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _data(T, N, ds, df, seed=0):
    g = torch.Generator().manual_seed(seed)
    state = torch.randn(T, N, ds, generator=g) * 3 + 1
    forcing = torch.randn(T, N, df, generator=g)
    times = torch.arange(T, dtype=torch.int64) * 3600 * 10 ** 9
    stats = (torch.randn(ds, generator=g), torch.rand(ds, generator=g) + 0.5,
             torch.randn(df, generator=g), torch.rand(df, generator=g) + 0.5)
    return state, forcing, times, stats


@pytest.mark.parametrize("T,N,ds,df,ar,past,future,standardize", [
    (20, 1000, 17, 6, 3, 1, 1, True),     # MEPS feature counts, reference defaults
    (16, 333, 5, 2, 1, 3, 0, True),       # past window longer than the two initial states
    (12, 64, 4, 3, 2, 0, 2, False),
])
def test_feed_equals_dataset_port(dev, T, N, ds, df, ar, past, future, standardize):
    from neural_lam_b200.device_feed import DeviceWeatherFeed
    state, forcing, times, (sm, ss, fm, fs) = _data(T, N, ds, df)
    ref = WeatherDatasetPort(state, forcing, times, sm, ss, fm, fs, ar, past, future, standardize)
    feed = DeviceWeatherFeed.from_arrays(state, forcing, times, chunk=7, state_mean=sm, state_std=ss,
                                         forcing_mean=fm, forcing_std=fs, ar_steps=ar,
                                         num_past_forcing_steps=past,
                                         num_future_forcing_steps=future, standardize=standardize,
                                         device=dev)
    assert len(feed) == len(ref)
    idx = [len(ref) - 1, 0, 2, len(ref) // 2]
    got = feed.batch(idx)
    want = [torch.stack(x) for x in zip(*[ref[i] for i in idx])]  # default collation
    for g, w, what in zip(got, want, ("init", "target", "forcing", "times")):
        assert g.shape == w.shape, what
        assert torch.equal(g.cpu(), w), what
    with pytest.raises(RuntimeError, match="resident"):
        feed.batch([len(ref) + 1])


def test_streaming_ring(dev):
    """Streaming: a ring of 8 time steps, one new slice per step; every batch equals the
    dataset over the whole series."""
    from neural_lam_b200.device_feed import DeviceWeatherFeed
    T, N, ds, df = 30, 500, 17, 6
    state, forcing, times, (sm, ss, fm, fs) = _data(T, N, ds, df, seed=1)
    ref = WeatherDatasetPort(state, forcing, times, sm, ss, fm, fs, 1, 1, 1, True)
    feed = DeviceWeatherFeed(N, ds, df, sm, ss, fm, fs, ar_steps=1, capacity=8, ring=True, device=dev)
    feed.append(state[:4].pin_memory(), forcing[:4].pin_memory(), times[:4])
    for t in range(4, T):
        feed.append(state[t:t + 1].pin_memory(), forcing[t:t + 1].pin_memory(), times[t:t + 1])
        i = t + 1 - 4  # newest sample whose window [i, i+4) is complete
        got = feed.batch([i, max(i - 3, feed.t_lo)])
        want = [torch.stack(x) for x in zip(*[ref[i], ref[max(i - 3, feed.t_lo)]])]
        for g, w in zip(got, want):
            assert torch.equal(g.cpu(), w)
    with pytest.raises(RuntimeError, match="resident"):
        feed.batch([0])  # overwritten long ago


def test_trainer_fit_from_feed(dev):
    """fit_from_feed (indices in, device-assembled batches) trains exactly like step() on the
    same batches built by the dataset restatement."""
    import tempfile

    from helpers import build_model_case, load_golden
    from neural_lam_b200 import config as nl_config
    from neural_lam_b200 import models, train
    from neural_lam_b200.device_feed import DeviceWeatherFeed
    case = load_golden("models.pt")["graphlam_multiscale_mean_d16"]

    def model():
        with tempfile.TemporaryDirectory() as root:
            ds_, args, batch = build_model_case(case["case"], root)
            m = models.GraphLAM(args, nl_config.default_config(), ds_)
        m.load_state_dict(case["state_dict"])
        return m.to(dev), batch

    ma, batch = model()
    mb, _ = model()
    N, d_s = batch[0].shape[2], batch[0].shape[3]
    d_fw = batch[2].shape[3]
    assert d_fw % 3 == 0
    T, ar = 12, batch[1].shape[1]
    state, forcing, times, (sm, ss, fm, fs) = _data(T, N, d_s, d_fw // 3, seed=2)
    ref = WeatherDatasetPort(state, forcing, times, sm, ss, fm, fs, ar, 1, 1, True)
    feed = DeviceWeatherFeed.from_arrays(state, forcing, times, state_mean=sm, state_std=ss,
                                         forcing_mean=fm, forcing_std=fs, ar_steps=ar, device=dev)
    index_batches = [[0, 3], [5, 1], [2, 4]]
    ta, tb = train.DataParallelTrainer(ma), train.DataParallelTrainer(mb)
    la = []
    for ib in index_batches:
        b = [torch.stack(x).to(dev) for x in zip(*[ref[i] for i in ib])]
        la.append(ta.step(tuple(b)).item())
    lb = tb.fit_from_feed(feed, index_batches)
    assert la == pytest.approx(lb, rel=1e-6)
    for p, q in zip(ma.parameters(), mb.parameters()):
        torch.testing.assert_close(p, q, rtol=1e-6, atol=1e-8)
