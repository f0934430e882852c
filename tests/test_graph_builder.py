"""CPU: create_graph (vectorised builder) against hashes of the reference's
own create_graph output (tests/golden/graphs.pt, oracle/make_golden.py) and
the file-format checks of /root/reference/tests/test_graph_creation.py:16-119."""
import hashlib
import os
import tempfile

import numpy as np
import pytest
import torch

from helpers import load_golden
from neural_lam_b200 import create_graph, utils

GRAPHS = load_golden("graphs.pt")


def _sha(t):
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def _xy(nx, ny, dx=2500.0):
    xy = np.zeros((nx, ny, 2))
    xy[:, :, 0] = (dx * np.arange(nx))[:, None]
    xy[:, :, 1] = (dx * np.arange(ny))[None, :]
    return xy


@pytest.mark.parametrize("name", sorted(GRAPHS))
def test_edge_index_bit_exact_vs_reference(name):
    e = GRAPHS[name]
    g = create_graph.build_graph(_xy(e["nx"], e["ny"]), e["n_max_levels"], e["hierarchical"])
    for key, shas in e["sha"].items():
        ts = g[key] if isinstance(g[key], list) else [g[key]]
        assert [tuple(t.shape) for t in ts] == e["shapes"][key]
        assert [_sha(t) for t in ts] == shas, key
    for key, sums in e["feat_sum"].items():
        ts = g[key] if isinstance(g[key], list) else [g[key]]
        assert [tuple(t.shape) for t in ts] == e["shapes"][key]
        for t, s in zip(ts, sums):
            assert t.dtype == torch.float32
            assert abs(t.double().abs().sum().item() - s) <= 1e-6 * max(1.0, abs(s)), key


@pytest.mark.parametrize("graph_name,n_max_levels,hierarchical",
                         [("1level", 1, False), ("multiscale", 3, False), ("hierarchical", 3, True)])
def test_graph_files_format(graph_name, n_max_levels, hierarchical):
    """Mirrors tests/test_graph_creation.py: required files, list lengths,
    edge_index.shape[0]==2, feature dim 3, mesh static dim 2."""
    required = ["m2m_edge_index.pt", "g2m_edge_index.pt", "m2g_edge_index.pt",
                "m2m_features.pt", "g2m_features.pt", "m2g_features.pt", "mesh_features.pt"]
    if hierarchical:
        required += ["mesh_up_edge_index.pt", "mesh_down_edge_index.pt",
                     "mesh_up_features.pt", "mesh_down_features.pt"]
    with tempfile.TemporaryDirectory() as d:
        create_graph.create_graph(d, _xy(100, 100, 5000.0), n_max_levels, hierarchical)
        assert sorted(os.listdir(d)) == sorted(required)
        for fn in required:
            r = torch.load(os.path.join(d, fn), weights_only=True)
            if fn.startswith("g2m") or fn.startswith("m2g"):
                assert isinstance(r, torch.Tensor)
                assert r.shape[0] == 2 if "edge_index" in fn else r.shape[1] == 3
            else:
                assert isinstance(r, list)
                want = n_max_levels if not ("up" in fn or "down" in fn) else n_max_levels - 1
                if not hierarchical:
                    want = 1
                assert len(r) == want
                for t in r:
                    if "edge_index" in fn:
                        assert t.shape[0] == 2
                    elif fn == "mesh_features.pt":
                        assert t.shape[1] == 2
                    else:
                        assert t.shape[1] == 3
        hier, g = utils.load_graph(d)
        assert hier == hierarchical
        # longest m2m edge normalises to 1 (utils.py:105-113)
        feats = list(g["m2m_features"]) if hierarchical else [g["m2m_features"]]
        assert abs(max(f[:, 0].max() for f in feats).item() - 1.0) < 1e-6
