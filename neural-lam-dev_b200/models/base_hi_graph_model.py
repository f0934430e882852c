"""Hierarchical mesh base: per-level embedders, init sweep up, read-out sweep
down (/root/reference/neural_lam/models/base_hi_graph_model.py:12-235)."""
from torch import nn

from .. import utils
from ..interaction_net import InteractionNet
from .base_graph_model import BaseGraphModel


class BaseHiGraphModel(BaseGraphModel):
    def __init__(self, args, config, datastore):
        super().__init__(args, config=config, datastore=datastore)
        self.num_levels = len(self.mesh_static_features)
        self.level_mesh_sizes = [f.shape[0] for f in self.mesh_static_features]
        d, h = args.hidden_dim, args.hidden_layers

        def embedders(in_dim, count):
            return nn.ModuleList(
                [utils.make_mlp([in_dim] + self.mlp_blueprint_end) for _ in range(count)])

        L = self.num_levels
        self.mesh_embedders = embedders(self.mesh_static_features[0].shape[1], L)
        self.mesh_same_embedders = embedders(self.m2m_features[0].shape[1], L)
        self.mesh_up_embedders = embedders(self.mesh_up_features[0].shape[1], L - 1)
        self.mesh_down_embedders = embedders(self.mesh_down_features[0].shape[1], L - 1)
        self.mesh_init_gnns = nn.ModuleList(
            [InteractionNet(ei, d, hidden_layers=h) for ei in self.mesh_up_edge_index])
        self.mesh_read_gnns = nn.ModuleList(
            [InteractionNet(ei, d, hidden_layers=h, update_edges=False)
             for ei in self.mesh_down_edge_index])

    def get_num_mesh(self):
        total = sum(f.shape[0] for f in self.mesh_static_features)
        return total, total - self.mesh_static_features[0].shape[0]

    def embedd_mesh_nodes(self):
        """Only the bottom level; the rest is embedded in process_step
        (base_hi_graph_model.py:115-122)."""
        return self.embed_mesh_static(self.mesh_embedders[0], self.mesh_static_features[0])

    def process_step(self, mesh_rep):
        """base_hi_graph_model.py:124-217."""
        batch_size = mesh_rep.shape[0]

        def embed(embs, feats):
            return [self.embed_static(e, f, batch_size) for e, f in zip(embs, feats)]

        mesh_rep_levels = [mesh_rep] + embed(list(self.mesh_embedders)[1:],
                                             list(self.mesh_static_features)[1:])
        mesh_same_rep = embed(self.mesh_same_embedders, self.m2m_features)
        mesh_up_rep = embed(self.mesh_up_embedders, self.mesh_up_features)
        mesh_down_rep = embed(self.mesh_down_embedders, self.mesh_down_features)

        # mesh init: sweep up, level l-1 -> l
        for level, gnn in enumerate(self.mesh_init_gnns, start=1):
            mesh_rep_levels[level], mesh_up_rep[level - 1] = gnn(
                mesh_rep_levels[level - 1], mesh_rep_levels[level], mesh_up_rep[level - 1])

        mesh_rep_levels, _, _, mesh_down_rep = self.hi_processor_step(
            mesh_rep_levels, mesh_same_rep, mesh_up_rep, mesh_down_rep)

        # read out: sweep down, level l+1 -> l
        for level in range(self.num_levels - 2, -1, -1):
            mesh_rep_levels[level] = self.mesh_read_gnns[level](
                mesh_rep_levels[level + 1], mesh_rep_levels[level], mesh_down_rep[level])
        return mesh_rep_levels[0]

    def hi_processor_step(self, mesh_rep_levels, mesh_same_rep, mesh_up_rep, mesh_down_rep):
        raise NotImplementedError("hi_process_step not implemented")
