"""CPU: integer side tables of the InteractionNet plan that are built on the host --
receiver-aligned tiles and the sender pre-reduction tables of the fused backward."""
import numpy as np

import helpers  # noqa: F401
from neural_lam_b200.interaction_net import _aligned_tiles, sender_partial_tables


def _random_plan(seed, n_rec, n_send, M):
    rng = np.random.default_rng(seed)
    recv = np.sort(rng.integers(0, n_rec, size=M))  # receiver-sorted order
    send = rng.integers(0, n_send, size=M)
    rowptr = np.concatenate(([0], np.cumsum(np.bincount(recv, minlength=n_rec))))
    return recv, send, rowptr


def test_aligned_tiles_cover_whole_segments():
    recv, send, rowptr = _random_plan(0, 900, 50, 7000)
    tile_seg = _aligned_tiles(rowptr)
    assert tile_seg[0] == 0 and tile_seg[-1] == 900 and np.all(np.diff(tile_seg) > 0)
    rows = np.diff(rowptr[tile_seg])
    assert rows.max() <= 128 and np.diff(tile_seg).max() <= 128
    # in-degree above a tile: not alignable
    big = np.concatenate(([0], np.cumsum([3, 200, 5])))
    assert _aligned_tiles(big) is None


def test_sender_partial_tables():
    recv, send, rowptr = _random_plan(1, 700, 40, 9000)
    tile_ptr = rowptr[_aligned_tiles(rowptr)]
    t = sender_partial_tables(send, tile_ptr, 40)
    n_tiles = len(tile_ptr) - 1
    assert t["row_ptr"][0] == 0 and t["row_ptr"][-1] == 9000 and t["tile_ptr"][-1] == t["n_sp"]
    vals = np.random.default_rng(2).standard_normal(9000)
    total = np.zeros(40)
    for tile in range(n_tiles):
        r0, r1 = tile_ptr[tile], tile_ptr[tile + 1]
        q0, q1 = t["tile_ptr"][tile], t["tile_ptr"][tile + 1]
        # the row lists of a tile start at the tile's first row and cover it exactly once
        assert t["row_ptr"][q0] == r0 and t["row_ptr"][q1] == r1
        local = t["rows"][r0:r1]
        assert sorted(local.tolist()) == list(range(r1 - r0))
        senders = t["sender"][q0:q1]
        assert np.all(np.diff(senders) > 0)  # one partial per distinct sender of the tile
        for q in range(q0, q1):
            rows = r0 + t["rows"][t["row_ptr"][q]:t["row_ptr"][q + 1]]
            assert np.all(np.diff(rows) > 0) and np.all(send[rows] == t["sender"][q])
            total[t["sender"][q]] += vals[rows].sum()
    want = np.zeros(40)
    np.add.at(want, send, vals)
    assert np.allclose(total, want)
    # m2g-like locality shrinks the row count; senders drawn from a large set do not
    wide = np.random.default_rng(3).integers(0, 100000, size=9000)
    assert sender_partial_tables(wide, tile_ptr, 100000)["n_sp"] > 0.9 * 9000
    send_local = (recv // 20) % 40
    assert sender_partial_tables(send_local, tile_ptr, 40)["n_sp"] < 0.2 * 9000


def test_sender_partial_tables_empty():
    t = sender_partial_tables(np.zeros(0, np.int64), np.zeros(1, np.int64), 5)
    assert t["n_sp"] == 0 and t["row_ptr"].tolist() == [0] and t["tile_ptr"].tolist() == [0]
