"""Hi-LAM: sequential down / up sweeps through the mesh hierarchy in every
processor layer (/root/reference/neural_lam/models/hi_lam.py:11-207)."""
from torch import nn

from ..interaction_net import InteractionNet
from .base_hi_graph_model import BaseHiGraphModel


class HiLAM(BaseHiGraphModel):
    def __init__(self, args, config, datastore):
        super().__init__(args, config=config, datastore=datastore)
        layers = range(args.processor_layers)
        # creation order as hi_lam.py:21-35 (matters for seed-matched init)
        self.mesh_down_gnns = nn.ModuleList([self.make_down_gnns(args) for _ in layers])
        self.mesh_down_same_gnns = nn.ModuleList([self.make_same_gnns(args) for _ in layers])
        self.mesh_up_gnns = nn.ModuleList([self.make_up_gnns(args) for _ in layers])
        self.mesh_up_same_gnns = nn.ModuleList([self.make_same_gnns(args) for _ in layers])

    @staticmethod
    def _gnns(args, edge_indices):
        return nn.ModuleList([
            InteractionNet(ei, args.hidden_dim, hidden_layers=args.hidden_layers)
            for ei in edge_indices])

    def make_same_gnns(self, args):
        return self._gnns(args, self.m2m_edge_index)

    def make_up_gnns(self, args):
        return self._gnns(args, self.mesh_up_edge_index)

    def make_down_gnns(self, args):
        return self._gnns(args, self.mesh_down_edge_index)

    def mesh_down_step(self, mesh_rep_levels, mesh_same_rep, mesh_down_rep, down_gnns,
                       same_gnns):
        """Top level same-level step, then (down, same) per level (hi_lam.py:82-124)."""
        top = self.num_levels - 1
        mesh_rep_levels[top], mesh_same_rep[top] = same_gnns[top](
            mesh_rep_levels[top], mesh_rep_levels[top], mesh_same_rep[top])
        for level in range(top - 1, -1, -1):
            new_node_rep, mesh_down_rep[level] = down_gnns[level](
                mesh_rep_levels[level + 1], mesh_rep_levels[level], mesh_down_rep[level])
            mesh_rep_levels[level], mesh_same_rep[level] = same_gnns[level](
                new_node_rep, new_node_rep, mesh_same_rep[level])
        return mesh_rep_levels, mesh_same_rep, mesh_down_rep

    def mesh_up_step(self, mesh_rep_levels, mesh_same_rep, mesh_up_rep, up_gnns, same_gnns):
        """Bottom level same-level step, then (up, same) per level (hi_lam.py:126-163)."""
        mesh_rep_levels[0], mesh_same_rep[0] = same_gnns[0](
            mesh_rep_levels[0], mesh_rep_levels[0], mesh_same_rep[0])
        for level in range(1, self.num_levels):
            new_node_rep, mesh_up_rep[level - 1] = up_gnns[level - 1](
                mesh_rep_levels[level - 1], mesh_rep_levels[level], mesh_up_rep[level - 1])
            mesh_rep_levels[level], mesh_same_rep[level] = same_gnns[level](
                new_node_rep, new_node_rep, mesh_same_rep[level])
        return mesh_rep_levels, mesh_same_rep, mesh_up_rep

    def hi_processor_step(self, mesh_rep_levels, mesh_same_rep, mesh_up_rep, mesh_down_rep):
        """hi_lam.py:165-207."""
        for down_gnns, down_same_gnns, up_gnns, up_same_gnns in zip(
                self.mesh_down_gnns, self.mesh_down_same_gnns, self.mesh_up_gnns,
                self.mesh_up_same_gnns):
            mesh_rep_levels, mesh_same_rep, mesh_down_rep = self.mesh_down_step(
                mesh_rep_levels, mesh_same_rep, mesh_down_rep, down_gnns, down_same_gnns)
            mesh_rep_levels, mesh_same_rep, mesh_up_rep = self.mesh_up_step(
                mesh_rep_levels, mesh_same_rep, mesh_up_rep, up_gnns, up_same_gnns)
        return mesh_rep_levels, mesh_same_rep, mesh_up_rep, mesh_down_rep
