"""TEST INFRASTRUCTURE ONLY -- the product path never imports this file.

CPU oracle: a plain-PyTorch (fp32, CPU) restatement of the reference hot path,
used as the checker in tests/, __graft_entry__.smoke() and bench.py's
`cpu_baseline` / `--impl reference` legs.  It exists because /root/reference
(and PyG) do not exist on the GPU box.

PINNING: oracle/make_golden.py runs the UNMODIFIED reference files behind
oracle/ref_stubs.py in the build container and (a) asserts this port reproduces
the reference's outputs and gradients on seeded inputs, (b) writes the vectors
to tests/golden/*.pt, which tests/test_oracle_golden.py re-checks everywhere.
The reference's own tests hold NO golden values for this path,
so reference outputs generated here are the pin.

Third-party arithmetic restated: torch-geometric==2.3.1 (pyproject.toml:26)
`MessagePassing.propagate` with node_dim=-2, flow source_to_target:
x_j = x.index_select(-2, edge_index[0]), x_i = x.index_select(-2,
edge_index[1]); sum = scatter_add over receivers, mean = sum /
count.clamp(min=1); `torch_geometric.nn.Sequential` child naming `module_{i}`.
"""
import os

import torch
from torch import nn


# ------------------------------------------------------------------ utils.py
def make_mlp(blueprint, layer_norm=True):
    """utils.py:191-214."""
    n_hidden = len(blueprint) - 2
    layers = []
    for i, (a, b) in enumerate(zip(blueprint[:-1], blueprint[1:])):
        layers.append(nn.Linear(a, b))
        if i != n_hidden:
            layers.append(nn.SiLU())
    if layer_norm:
        layers.append(nn.LayerNorm(blueprint[-1]))
    return nn.Sequential(*layers)


class SplitMLPs(nn.Module):
    """interaction_net.py:134-163."""

    def __init__(self, mlps, chunk_sizes):
        super().__init__()
        assert len(mlps) == len(chunk_sizes)
        self.mlps = nn.ModuleList(mlps)
        self.chunk_sizes = list(chunk_sizes)

    def forward(self, x):
        parts = torch.split(x, self.chunk_sizes, dim=-2)
        return torch.cat([m(p) for m, p in zip(self.mlps, parts)], dim=-2)


def reindex_edge_index(edge_index):
    """interaction_net.py:55-61 -> (local edge_index with senders offset by
    num_rec, num_rec)."""
    ei = edge_index - edge_index.min(dim=1, keepdim=True)[0]
    num_rec = int(ei[1].max()) + 1
    ei[0] = ei[0] + num_rec
    return ei, num_rec


class InteractionNet(nn.Module):
    """interaction_net.py:10-131 with PyG's propagate written out."""

    def __init__(self, edge_index, input_dim, update_edges=True,
                 hidden_layers=1, hidden_dim=None, edge_chunk_sizes=None,
                 aggr_chunk_sizes=None, aggr="sum"):
        super().__init__()
        assert aggr in ("sum", "mean")
        self.aggr = aggr
        hidden_dim = input_dim if hidden_dim is None else hidden_dim
        ei, self.num_rec = reindex_edge_index(edge_index)
        self.register_buffer("edge_index", ei, persistent=False)
        e_bp = [3 * input_dim] + [hidden_dim] * (hidden_layers + 1)
        a_bp = [2 * input_dim] + [hidden_dim] * (hidden_layers + 1)
        if edge_chunk_sizes is None:
            self.edge_mlp = make_mlp(e_bp)
        else:
            self.edge_mlp = SplitMLPs(
                [make_mlp(e_bp) for _ in edge_chunk_sizes], edge_chunk_sizes)
        if aggr_chunk_sizes is None:
            self.aggr_mlp = make_mlp(a_bp)
        else:
            self.aggr_mlp = SplitMLPs(
                [make_mlp(a_bp) for _ in aggr_chunk_sizes], aggr_chunk_sizes)
        self.update_edges = update_edges

    def forward(self, send_rep, rec_rep, edge_rep):
        nodes = torch.cat((rec_rep, send_rep), dim=-2)  # :102
        x_j = nodes.index_select(-2, self.edge_index[0])
        x_i = nodes.index_select(-2, self.edge_index[1])
        msg = self.edge_mlp(torch.cat((edge_rep, x_j, x_i), dim=-1))  # :121
        shape = list(msg.shape)
        shape[-2] = self.num_rec
        aggr = msg.new_zeros(shape).index_add_(-2, self.edge_index[1], msg)
        if self.aggr == "mean":
            cnt = torch.bincount(self.edge_index[1], minlength=self.num_rec)
            aggr = aggr / cnt.clamp(min=1).to(msg.dtype).unsqueeze(-1)
        rec_rep = rec_rep + self.aggr_mlp(torch.cat((rec_rep, aggr), dim=-1))
        if self.update_edges:
            return rec_rep, edge_rep + msg  # :109-113
        return rec_rep


class _Chain(nn.Module):
    """pyg.nn.Sequential("mesh_rep, edge_rep", [(net, "mesh_rep, mesh_rep,
    edge_rep -> mesh_rep, edge_rep")...]) -- graph_lam.py:51-57."""

    def __init__(self, nets):
        super().__init__()
        for i, net in enumerate(nets):
            self.add_module(f"module_{i}", net)
        self.n = len(nets)

    def forward(self, mesh_rep, edge_rep):
        for i in range(self.n):
            mesh_rep, edge_rep = getattr(self, f"module_{i}")(
                mesh_rep, mesh_rep, edge_rep)
        return mesh_rep, edge_rep


# ------------------------------------------------------------------ metrics.py
def _reduce(v, mask, average_grid, sum_vars):
    """metrics.py:21-53: mask the grid nodes, mean over the grid, sum over variables."""
    if mask is not None:
        v = v[..., mask, :]
    if average_grid:
        v = v.mean(dim=-2)
    if sum_vars:
        v = v.sum(dim=-1)
    return v


def wmse(pred, target, pred_std, mask=None, average_grid=True, sum_vars=True):
    """metrics.py:56-84."""
    return _reduce((pred - target) ** 2 / (pred_std**2), mask, average_grid, sum_vars)


def mse(pred, target, pred_std, mask=None, average_grid=True, sum_vars=True):
    """metrics.py:87-113 (weights replaced by ones)."""
    return wmse(pred, target, torch.ones_like(pred_std), mask, average_grid, sum_vars)


def mae(pred, target, pred_std, mask=None, average_grid=True, sum_vars=True):
    """metrics.py:142-163."""
    return _reduce((pred - target).abs(), mask, average_grid, sum_vars)


def nll(pred, target, pred_std, mask=None, average_grid=True, sum_vars=True):
    """metrics.py:176-201: Gaussian negative log-likelihood."""
    v = -torch.distributions.Normal(pred, pred_std).log_prob(target)
    return _reduce(v, mask, average_grid, sum_vars)


class _BufList(nn.Module):
    def __init__(self, tensors):
        super().__init__()
        self.n = len(tensors)
        for i, t in enumerate(tensors):
            self.register_buffer(f"b{i}", t, persistent=False)

    def __getitem__(self, i):
        return getattr(self, f"b{i}")

    def __len__(self):
        return self.n

    def __iter__(self):
        return (self[i] for i in range(self.n))


def load_graph(path):
    """utils.py:36-188."""
    ld = lambda fn: torch.load(os.path.join(path, fn), weights_only=True)
    m2m_ei, m2m_f, mesh_f = ld("m2m_edge_index.pt"), ld("m2m_features.pt"), ld("mesh_features.pt")
    longest = max(torch.max(f[:, 0]) for f in m2m_f)
    g = dict(g2m_edge_index=ld("g2m_edge_index.pt"), m2g_edge_index=ld("m2g_edge_index.pt"),
             g2m_features=ld("g2m_features.pt") / longest, m2g_features=ld("m2g_features.pt") / longest)
    hier = len(m2m_ei) > 1
    if hier:
        g.update(m2m_edge_index=m2m_ei, m2m_features=[f / longest for f in m2m_f],
                 mesh_static_features=mesh_f,
                 mesh_up_edge_index=ld("mesh_up_edge_index.pt"),
                 mesh_down_edge_index=ld("mesh_down_edge_index.pt"),
                 mesh_up_features=[f / longest for f in ld("mesh_up_features.pt")],
                 mesh_down_features=[f / longest for f in ld("mesh_down_features.pt")])
    else:
        g.update(m2m_edge_index=m2m_ei[0], m2m_features=m2m_f[0] / longest,
                 mesh_static_features=mesh_f[0])
    return hier, g


# ------------------------------------------------------------------ models
class ARModel(nn.Module):
    """models/ar_model.py:31-131 (buffers), :204-309 (rollout + loss)."""

    def __init__(self, args, config, datastore):
        super().__init__()
        self.args = args
        d_state = datastore.get_num_data_vars("state")
        d_forc = datastore.get_num_data_vars("forcing")
        f32 = torch.float32
        static = datastore.get_dataarray(category="static", split=None).transpose(
            "grid_index", "static_feature").values
        self.register_buffer("grid_static_features", torch.tensor(static, dtype=f32), persistent=False)
        st = datastore.get_standardization_dataarray(category="state")
        for name, da in (("state_mean", st.state_mean), ("state_std", st.state_std),
                         ("diff_mean", st.state_diff_mean), ("diff_std", st.state_diff_std)):
            self.register_buffer(name, torch.tensor(da.values, dtype=f32), persistent=False)
        n_vars = len(datastore.get_vars_names(category="state"))
        # loss_weighting.py:8-106: manual per-variable weights or uniform 1/n
        wcfg = getattr(getattr(config, "training", None), "state_feature_weighting", None)
        if wcfg is not None and hasattr(wcfg, "weights"):
            names = datastore.get_vars_names(category="state")
            assert set(wcfg.weights) == set(names)
            weights = [wcfg.weights[n] for n in names]
        else:
            weights = [1.0 / n_vars] * n_vars
        self.feature_weights = torch.tensor(weights, dtype=f32)
        self.output_std = bool(args.output_std)
        if self.output_std:
            self.grid_output_dim = 2 * d_state
        else:
            self.grid_output_dim = d_state
            self.register_buffer("per_var_std", self.diff_std / torch.sqrt(self.feature_weights),
                                 persistent=False)
        self.num_grid_nodes, d_static = self.grid_static_features.shape
        self.grid_dim = (2 * self.grid_output_dim + d_static + d_forc * (
            args.num_past_forcing_steps + args.num_future_forcing_steps + 1))
        self.loss = {"wmse": wmse, "mse": mse, "nll": nll}[args.loss.lower()]
        bm = torch.tensor(datastore.boundary_mask.values, dtype=f32).unsqueeze(1)
        self.register_buffer("boundary_mask", bm, persistent=False)
        self.register_buffer("interior_mask", 1.0 - bm, persistent=False)

    @staticmethod
    def expand_to_batch(x, b):
        return x.unsqueeze(0).expand(b, -1, -1)

    def unroll_prediction(self, init_states, forcing_features, true_states):
        prev_prev, prev = init_states[:, 0], init_states[:, 1]
        preds, stds = [], []
        for i in range(forcing_features.shape[1]):
            pred, std = self.predict_step(prev, prev_prev, forcing_features[:, i])
            new = self.boundary_mask * true_states[:, i] + self.interior_mask * pred
            preds.append(new)
            if self.output_std:
                stds.append(std)
            prev_prev, prev = prev, new
        pred_std = torch.stack(stds, dim=1) if self.output_std else self.per_var_std
        return torch.stack(preds, dim=1), pred_std

    def training_step(self, batch):
        init_states, target, forcing, _ = batch
        pred, pred_std = self.unroll_prediction(init_states, forcing, target)
        mask = self.interior_mask[:, 0].to(torch.bool)
        return torch.mean(self.loss(pred, target, pred_std, mask=mask))

    def validation_step(self, batch):
        """ar_model.py:324-361 without the Lightning logging: ({name: value}, entry MSEs
        (B, pred_steps, d_f))."""
        init_states, target, forcing, _ = batch
        pred, pred_std = self.unroll_prediction(init_states, forcing, target)
        mask = self.interior_mask[:, 0].to(torch.bool)
        time_step_loss = torch.mean(self.loss(pred, target, pred_std, mask=mask), dim=0)
        log = {f"val_loss_unroll{step}": time_step_loss[step - 1]
               for step in self.args.val_steps_to_log if step <= len(time_step_loss)}
        log["val_mean_loss"] = torch.mean(time_step_loss)
        return log, mse(pred, target, pred_std, mask=mask, sum_vars=False)

    def test_step(self, batch):
        """ar_model.py:375-435 without logging / plotting: ({name: value}, {"mse", "mae":
        (B, pred_steps, d_f)}, spatial loss maps (B, N_log, num_grid_nodes))."""
        init_states, target, forcing, _ = batch
        pred, pred_std = self.unroll_prediction(init_states, forcing, target)
        mask = self.interior_mask[:, 0].to(torch.bool)
        time_step_loss = torch.mean(self.loss(pred, target, pred_std, mask=mask), dim=0)
        log = {f"test_loss_unroll{step}": time_step_loss[step - 1]
               for step in self.args.val_steps_to_log}
        log["test_mean_loss"] = torch.mean(time_step_loss)
        entry = {"mse": mse(pred, target, pred_std, mask=mask, sum_vars=False),
                 "mae": mae(pred, target, pred_std, mask=mask, sum_vars=False)}
        spatial = self.loss(pred, target, pred_std, average_grid=False)
        return log, entry, spatial[:, [step - 1 for step in self.args.val_steps_to_log]]

    def configure_optimizers(self):
        return torch.optim.AdamW(self.parameters(), lr=self.args.lr, betas=(0.9, 0.95))


class BaseGraphModel(ARModel):
    """models/base_graph_model.py:18-80, 106-177."""

    def __init__(self, args, config, datastore):
        super().__init__(args, config, datastore)
        self.hierarchical, g = load_graph(str(datastore.root_path / "graph" / args.graph))
        for k, v in g.items():
            if torch.is_tensor(v):
                self.register_buffer(k, v, persistent=False)
            else:
                setattr(self, k, _BufList(v))
        d, h = args.hidden_dim, args.hidden_layers
        self.bp_end = [d] * (h + 1)
        self.grid_embedder = make_mlp([self.grid_dim] + self.bp_end)
        self.g2m_embedder = make_mlp([self.g2m_features.shape[1]] + self.bp_end)
        self.m2g_embedder = make_mlp([self.m2g_features.shape[1]] + self.bp_end)
        self.g2m_gnn = InteractionNet(self.g2m_edge_index, d, hidden_layers=h, update_edges=False)
        self.encoding_grid_mlp = make_mlp([d] + self.bp_end)
        self.m2g_gnn = InteractionNet(self.m2g_edge_index, d, hidden_layers=h, update_edges=False)
        self.output_map = make_mlp([d] * (h + 1) + [self.grid_output_dim], layer_norm=False)

    def predict_step(self, prev_state, prev_prev_state, forcing):
        b = prev_state.shape[0]
        feats = torch.cat((prev_state, prev_prev_state, forcing,
                           self.expand_to_batch(self.grid_static_features, b)), dim=-1)
        grid_emb = self.grid_embedder(feats)
        g2m_emb = self.g2m_embedder(self.g2m_features)
        m2g_emb = self.m2g_embedder(self.m2g_features)
        mesh_emb = self.embedd_mesh_nodes()
        mesh_rep = self.g2m_gnn(grid_emb, self.expand_to_batch(mesh_emb, b),
                                self.expand_to_batch(g2m_emb, b))
        grid_rep = grid_emb + self.encoding_grid_mlp(grid_emb)
        mesh_rep = self.process_step(mesh_rep)
        grid_rep = self.m2g_gnn(mesh_rep, grid_rep, self.expand_to_batch(m2g_emb, b))
        out = self.output_map(grid_rep)
        if self.output_std:
            mean, std_raw = out.chunk(2, dim=-1)
            std = torch.nn.functional.softplus(std_raw)
        else:
            mean, std = out, None
        return prev_state + mean * self.diff_std + self.diff_mean, std


class GraphLAM(BaseGraphModel):
    """models/graph_lam.py:20-91."""

    def __init__(self, args, config, datastore):
        super().__init__(args, config, datastore)
        assert not self.hierarchical
        self.mesh_embedder = make_mlp([self.mesh_static_features.shape[1]] + self.bp_end)
        self.m2m_embedder = make_mlp([self.m2m_features.shape[1]] + self.bp_end)
        self.processor = _Chain([
            InteractionNet(self.m2m_edge_index, args.hidden_dim,
                           hidden_layers=args.hidden_layers, aggr=args.mesh_aggr)
            for _ in range(args.processor_layers)])

    def embedd_mesh_nodes(self):
        return self.mesh_embedder(self.mesh_static_features)

    def process_step(self, mesh_rep):
        m2m_emb = self.m2m_embedder(self.m2m_features)
        mesh_rep, _ = self.processor(mesh_rep, self.expand_to_batch(m2m_emb, mesh_rep.shape[0]))
        return mesh_rep


class BaseHiGraphModel(BaseGraphModel):
    """models/base_hi_graph_model.py:17-217."""

    def __init__(self, args, config, datastore):
        super().__init__(args, config, datastore)
        self.num_levels = len(self.mesh_static_features)
        self.level_mesh_sizes = [f.shape[0] for f in self.mesh_static_features]
        d, h = args.hidden_dim, args.hidden_layers
        L = self.num_levels
        mk = lambda dim, n: nn.ModuleList([make_mlp([dim] + self.bp_end) for _ in range(n)])
        self.mesh_embedders = mk(self.mesh_static_features[0].shape[1], L)
        self.mesh_same_embedders = mk(self.m2m_features[0].shape[1], L)
        self.mesh_up_embedders = mk(self.mesh_up_features[0].shape[1], L - 1)
        self.mesh_down_embedders = mk(self.mesh_down_features[0].shape[1], L - 1)
        self.mesh_init_gnns = nn.ModuleList(
            [InteractionNet(ei, d, hidden_layers=h) for ei in self.mesh_up_edge_index])
        self.mesh_read_gnns = nn.ModuleList(
            [InteractionNet(ei, d, hidden_layers=h, update_edges=False)
             for ei in self.mesh_down_edge_index])

    def embedd_mesh_nodes(self):
        return self.mesh_embedders[0](self.mesh_static_features[0])

    def process_step(self, mesh_rep):
        b = mesh_rep.shape[0]
        ex = self.expand_to_batch
        levels = [mesh_rep] + [ex(emb(f), b) for emb, f in zip(
            list(self.mesh_embedders)[1:], list(self.mesh_static_features)[1:])]
        same = [ex(emb(f), b) for emb, f in zip(self.mesh_same_embedders, self.m2m_features)]
        up = [ex(emb(f), b) for emb, f in zip(self.mesh_up_embedders, self.mesh_up_features)]
        down = [ex(emb(f), b) for emb, f in zip(self.mesh_down_embedders, self.mesh_down_features)]
        for l, gnn in enumerate(self.mesh_init_gnns, start=1):
            levels[l], up[l - 1] = gnn(levels[l - 1], levels[l], up[l - 1])
        levels, _, _, down = self.hi_processor_step(levels, same, up, down)
        for l, gnn in zip(range(self.num_levels - 2, -1, -1), reversed(self.mesh_read_gnns)):
            levels[l] = gnn(levels[l + 1], levels[l], down[l])
        return levels[0]


class HiLAM(BaseHiGraphModel):
    """models/hi_lam.py:17-207."""

    def __init__(self, args, config, datastore):
        super().__init__(args, config, datastore)
        d, h, P = args.hidden_dim, args.hidden_layers, args.processor_layers
        mk = lambda eis: nn.ModuleList([InteractionNet(ei, d, hidden_layers=h) for ei in eis])
        self.mesh_down_gnns = nn.ModuleList([mk(self.mesh_down_edge_index) for _ in range(P)])
        self.mesh_down_same_gnns = nn.ModuleList([mk(self.m2m_edge_index) for _ in range(P)])
        self.mesh_up_gnns = nn.ModuleList([mk(self.mesh_up_edge_index) for _ in range(P)])
        self.mesh_up_same_gnns = nn.ModuleList([mk(self.m2m_edge_index) for _ in range(P)])

    def hi_processor_step(self, levels, same, up, down):
        L = self.num_levels
        for dn, dn_same, upg, up_same in zip(self.mesh_down_gnns, self.mesh_down_same_gnns,
                                             self.mesh_up_gnns, self.mesh_up_same_gnns):
            levels[-1], same[-1] = dn_same[-1](levels[-1], levels[-1], same[-1])
            for l in range(L - 2, -1, -1):
                new, down[l] = dn[l](levels[l + 1], levels[l], down[l])
                levels[l], same[l] = dn_same[l](new, new, same[l])
            levels[0], same[0] = up_same[0](levels[0], levels[0], same[0])
            for l in range(1, L):
                new, up[l - 1] = upg[l - 1](levels[l - 1], levels[l], up[l - 1])
                levels[l], same[l] = up_same[l](new, new, same[l])
        return levels, same, up, down


class HiLAMParallel(BaseHiGraphModel):
    """models/hi_lam_parallel.py:20-99."""

    def __init__(self, args, config, datastore):
        super().__init__(args, config, datastore)
        eis = list(self.m2m_edge_index) + list(self.mesh_up_edge_index) + list(self.mesh_down_edge_index)
        total = torch.cat(eis, dim=1)
        self.edge_split_sections = [e.shape[1] for e in eis]
        if args.processor_layers == 0:
            self.processor = lambda x, e: (x, e)
        else:
            self.processor = _Chain([
                InteractionNet(total, args.hidden_dim, hidden_layers=args.hidden_layers,
                               edge_chunk_sizes=self.edge_split_sections,
                               aggr_chunk_sizes=self.level_mesh_sizes)
                for _ in range(args.processor_layers)])

    def hi_processor_step(self, levels, same, up, down):
        L = self.num_levels
        mesh = torch.cat(levels, dim=1)
        edges = torch.cat(list(same) + list(up) + list(down), dim=1)
        mesh, edges = self.processor(mesh, edges)
        levels = list(torch.split(mesh, self.level_mesh_sizes, dim=1))
        parts = torch.split(edges, self.edge_split_sections, dim=1)
        return levels, parts[:L], parts[L:2 * L - 1], parts[2 * L - 1:]


MODELS = {"graph_lam": GraphLAM, "hi_lam": HiLAM, "hi_lam_parallel": HiLAMParallel}
