"""Stand-in for `torch_geometric.nn.Sequential("mesh_rep, edge_rep",
[(net, "mesh_rep, mesh_rep, edge_rep -> mesh_rep, edge_rep"), ...])` as used at
/root/reference/neural_lam/models/graph_lam.py:51-57 and hi_lam_parallel.py:
47-53: children are named `module_{i}` so `state_dict` keys
(`processor.module_0.edge_mlp.0.weight`, ...) match reference checkpoints."""
from torch import nn


class ProcessorSequential(nn.Module):
    def __init__(self, nets):
        super().__init__()
        self._n = len(nets)
        for i, net in enumerate(nets):
            self.add_module(f"module_{i}", net)

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        return getattr(self, f"module_{i}")

    def forward(self, mesh_rep, edge_rep):
        for i in range(self._n):
            mesh_rep, edge_rep = self[i](mesh_rep, mesh_rep, edge_rep)
        return mesh_rep, edge_rep
