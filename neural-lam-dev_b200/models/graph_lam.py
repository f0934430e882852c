"""GraphLAM: encode-process-decode on a flat (1-level or multiscale) mesh
(/root/reference/neural_lam/models/graph_lam.py:12-91)."""
from .. import utils
from ..interaction_net import InteractionNet
from ..sequential import ProcessorSequential
from .base_graph_model import BaseGraphModel


class GraphLAM(BaseGraphModel):
    def __init__(self, args, config, datastore):
        super().__init__(args, config=config, datastore=datastore)
        assert not self.hierarchical, "GraphLAM does not use a hierarchical mesh graph"
        mesh_dim = self.mesh_static_features.shape[1]
        _, m2m_dim = self.m2m_features.shape
        self.mesh_embedder = utils.make_mlp([mesh_dim] + self.mlp_blueprint_end)
        self.m2m_embedder = utils.make_mlp([m2m_dim] + self.mlp_blueprint_end)
        self.processor = ProcessorSequential([
            InteractionNet(self.m2m_edge_index, args.hidden_dim,
                           hidden_layers=args.hidden_layers, aggr=args.mesh_aggr)
            for _ in range(args.processor_layers)])

    def get_num_mesh(self):
        return self.mesh_static_features.shape[0], 0

    def embedd_mesh_nodes(self):
        return self.embed_mesh_static(self.mesh_embedder, self.mesh_static_features)

    def process_step(self, mesh_rep):
        """graph_lam.py:73-91: embed m2m edges, run the processor layers."""
        m2m_emb = self.embed_static(self.m2m_embedder, self.m2m_features, mesh_rep.shape[0])
        mesh_rep, _ = self.processor(mesh_rep, m2m_emb)
        return mesh_rep
