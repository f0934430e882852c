"""Host-side helpers mirroring /root/reference/neural_lam/utils.py:
`BufferList` (:11-33), `load_graph` (:36-188) and `make_mlp` (:191-214).

`make_mlp` returns an `nn.Sequential` with the SAME children (and therefore
the same `state_dict` keys: `0.weight`, `0.bias`, `2.weight`, `2.bias`,
`3.weight`, `3.bias`) as the reference, so reference checkpoints load, but its
`forward` runs ONE fused sm_100a kernel (Linear -> SiLU -> Linear -> LayerNorm)
through the C-ABI instead of four ATen ops.
"""
import os

import torch
from torch import nn


class BufferList(nn.Module):
    """Module holding an indexable list of buffers `b0, b1, ...`
    (utils.py:11-33)."""

    def __init__(self, buffer_tensors, persistent=True):
        super().__init__()
        self.n_buffers = len(buffer_tensors)
        for i, tensor in enumerate(buffer_tensors):
            self.register_buffer(f"b{i}", tensor, persistent=persistent)

    def __getitem__(self, key):
        return getattr(self, f"b{key}")

    def __len__(self):
        return self.n_buffers

    def __iter__(self):
        return (self[i] for i in range(len(self)))


_GRAPH_FILES = ("m2m", "g2m", "m2g")


def load_graph(graph_dir_path, device="cpu"):
    """Load the graph directory written by `create_graph` (utils.py:36-188).

    Edge features of every edge set are divided by the longest m2m edge
    (column 0 = length, utils.py:105-113,145-158).  Non-hierarchical graphs
    are unwrapped from their one-element lists (utils.py:163-174).
    Returns `(hierarchical, dict)` with the reference's keys.
    """

    def load(name):
        return torch.load(
            os.path.join(graph_dir_path, name),
            map_location=device,
            weights_only=True,
        )

    m2m_edge_index = load("m2m_edge_index.pt")
    m2m_features = load("m2m_features.pt")
    mesh_static = load("mesh_features.pt")
    n_levels = len(m2m_edge_index)
    assert len(m2m_features) == n_levels, "Inconsistent number of levels in mesh"
    assert len(mesh_static) == n_levels, "Inconsistent number of levels in mesh"
    hierarchical = n_levels > 1

    longest_edge = max(torch.max(f[:, 0]) for f in m2m_features)

    def scaled(feats):
        return [f / longest_edge for f in feats]

    graph = {
        "g2m_edge_index": load("g2m_edge_index.pt"),
        "m2g_edge_index": load("m2g_edge_index.pt"),
        "g2m_features": load("g2m_features.pt") / longest_edge,
        "m2g_features": load("m2g_features.pt") / longest_edge,
    }
    if hierarchical:
        graph["m2m_edge_index"] = BufferList(m2m_edge_index, persistent=False)
        graph["m2m_features"] = BufferList(scaled(m2m_features), persistent=False)
        graph["mesh_static_features"] = BufferList(mesh_static, persistent=False)
        for direction in ("up", "down"):
            graph[f"mesh_{direction}_edge_index"] = BufferList(
                load(f"mesh_{direction}_edge_index.pt"), persistent=False
            )
            graph[f"mesh_{direction}_features"] = BufferList(
                scaled(load(f"mesh_{direction}_features.pt")), persistent=False
            )
    else:
        graph["m2m_edge_index"] = m2m_edge_index[0]
        graph["m2m_features"] = scaled(m2m_features)[0]
        graph["mesh_static_features"] = mesh_static[0]
        for direction in ("up", "down"):
            graph[f"mesh_{direction}_edge_index"] = []
            graph[f"mesh_{direction}_features"] = []
    return hierarchical, graph


class FusedMLP(nn.Sequential):
    """`nn.Sequential(Linear, SiLU, [Linear, SiLU]*, Linear, [LayerNorm])`
    whose forward is one fused CUDA kernel (see ops.mlp)."""

    def forward(self, x):  # pylint: disable=arguments-renamed
        from . import ops

        return ops.mlp_forward(self, x)


def make_mlp(blueprint, layer_norm=True):
    """MLP `blueprint[0] -> ... -> blueprint[-1]`, SiLU between layers,
    optional output LayerNorm (utils.py:191-214)."""
    hidden_layers = len(blueprint) - 2
    assert hidden_layers >= 0, "Invalid MLP blueprint"
    if hidden_layers == 0:  # fail at construction, not at the first forward
        raise NotImplementedError(
            "make_mlp: hidden_layers = 0 (a single Linear) has no fused kernel; the kernels "
            "compute Linear -> SiLU -> Linear blocks (hidden_layers >= 1)")
    layers = []
    for i, (d_in, d_out) in enumerate(zip(blueprint[:-1], blueprint[1:])):
        layers.append(nn.Linear(d_in, d_out))
        if i != hidden_layers:
            layers.append(nn.SiLU())
    if layer_norm:
        layers.append(nn.LayerNorm(blueprint[-1]))
    return FusedMLP(*layers)
