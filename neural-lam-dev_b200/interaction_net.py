"""Drop-in `InteractionNet` (Battaglia et al. 2016) with the constructor /
forward signature, child-module names, `state_dict` keys, non-persistent
`edge_index` buffer and `num_rec` attribute of
/root/reference/neural_lam/interaction_net.py:10-163 -- but `forward` is a
single autograd op over the sm_100a kernels in csrc/ (gather, fused edge MLP,
deterministic CSR segment sum, fused node MLP; backward recomputes edge
activations and scatters sender gradients through a transposed CSR) instead of
PyG's MessagePassing.propagate + torch_scatter.

Index semantics are reproduced literally: senders and
receivers are re-based by their own minimum id, `num_rec = max(receiver)+1`,
senders index rows of `send_rep` after that re-basing.
"""
import torch
from torch import nn

from . import ops, utils


import numpy as np  # noqa: E402


class _GraphPlan:
    """Per-device integer side tables derived from `edge_index` (built on the
    GPU by nlam_csr_build; never saved, never trained)."""

    def __init__(self, edge_index, num_rec, edge_chunk_sizes, aggr_chunk_sizes):
        dev = edge_index.device
        self.device = dev
        self.num_rec = num_rec
        self.n_edges = edge_index.shape[1]
        self.send32 = (edge_index[0] - num_rec).to(torch.int32).contiguous()
        self.recv32 = edge_index[1].to(torch.int32).contiguous()
        self.n_send_idx = int(self.send32.max().item()) + 1 if self.n_edges > 0 else 0
        # receiver-sorted CSR (aggregation) and sender-sorted CSR (grad of x_j gather)
        self.rowptr, self.perm, self.inv_deg = ops.csr_build(self.recv32, num_rec, True)
        self.t_rowptr, self.t_perm, _ = ops.csr_build(self.send32, self.n_send_idx, False)
        # receiver-sorted views + receiver-aligned tiles (fused aggregation path)
        self._build_aligned()
        self.edge_tiles = (ops.TileTable(edge_chunk_sizes, dev)
                           if edge_chunk_sizes is not None else None)
        self.aggr_tiles = (ops.TileTable(aggr_chunk_sizes, dev)
                           if aggr_chunk_sizes is not None else None)

    def _build_aligned(self):
        dev = self.device
        rowptr = self.rowptr.cpu().numpy().astype(np.int64)
        tile_seg = _aligned_tiles(rowptr)
        self.alignable = tile_seg is not None and self.n_edges > 0
        if not self.alignable:
            return
        perm64 = self.perm.long()
        self.send_sorted = self.send32[perm64].contiguous()
        self.recv_sorted = self.recv32[perm64].contiguous()
        mk = lambda a: torch.tensor(np.asarray(a, dtype=np.int32), device=dev)
        self.a_tile_seg = mk(tile_seg)
        self.a_tile_ptr = mk(rowptr[tile_seg])
        self.a_n_tiles = len(tile_seg) - 1
        # sender CSR over the receiver-sorted positions (backward of the x_j gather)
        self.ts_rowptr, self.ts_perm, _ = ops.csr_build(self.send_sorted, self.n_send_idx, False)
        self._build_sender_partials(rowptr[tile_seg])

    def _build_sender_partials(self, tile_ptr):
        """Sender pre-reduction tables of the fused backward (nlam_rowmlp_bwd.sp_*): inside
        a receiver-aligned tile the edges of one sender are summed on chip, so the kernel
        writes one PARTIAL gradient row per (tile, distinct sender) instead of one row per
        edge; a CSR over the partials (by sender, ascending tile order = deterministic)
        finishes the per-sender sum.  m2g (4 nearest mesh nodes per grid node,
        create_graph.py:506-519): ~11x fewer rows; m2m: ~2.4x; g2m (every grid node sends to
        ~1.2 mesh nodes): nothing to gain, tables not built."""
        self.sp = None
        tab = sender_partial_tables(self.send_sorted.cpu().numpy(), tile_ptr, self.n_send_idx)
        if tab["n_sp"] > 0.6 * self.n_edges:
            return
        mk = lambda a: torch.tensor(np.asarray(a, dtype=np.int32), device=self.device)
        rowptr, perm, _ = ops.csr_build(mk(tab["sender"]), self.n_send_idx, False)
        self.sp = {"n_sp": tab["n_sp"], "tile_ptr": mk(tab["tile_ptr"]), "row_ptr": mk(tab["row_ptr"]),
                   "rows": mk(tab["rows"]), "csr_rowptr": rowptr, "csr_perm": perm}

    def aligned_tables(self, aggr):
        return {"tile_ptr": self.a_tile_ptr, "n_tiles": self.a_n_tiles,
                "tile_seg": self.a_tile_seg, "seg_ptr": self.rowptr, "n_seg": self.num_rec,
                "scale": self.inv_deg if aggr == "mean" else None, "out_idx": self.perm}


def sender_partial_tables(send_sorted, tile_ptr, n_send):
    """Host tables of the sender pre-reduction (see _GraphPlan._build_sender_partials).
    send_sorted[M]: sender of every edge in receiver-sorted order; tile_ptr[n_tiles + 1]: row
    ranges of the receiver-aligned tiles.  Partial q = one (tile, sender) pair; partials are
    numbered tile by tile (tile t owns [tile_ptr_q[t], tile_ptr_q[t+1])), inside a tile by
    sender id; its rows are rows[row_ptr[q] : row_ptr[q+1]] (tile-local ids, ascending), and
    the row lists of tile t start at position tile_ptr[t] of `rows`."""
    send = np.asarray(send_sorted, dtype=np.int64)
    tile_ptr = np.asarray(tile_ptr, dtype=np.int64)
    M, n_tiles = send.shape[0], tile_ptr.shape[0] - 1
    tile_of_row = np.repeat(np.arange(n_tiles, dtype=np.int64), np.diff(tile_ptr))
    key = tile_of_row * max(int(n_send), 1) + send
    order = np.argsort(key, kind="stable")  # rows grouped by (tile, sender), ascending inside
    ks = key[order]
    starts = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]]) if M > 0 else np.zeros(0, np.int64)
    return {
        "n_sp": int(starts.shape[0]),
        "sender": send[order[starts]],
        "tile_ptr": np.searchsorted(tile_of_row[order[starts]], np.arange(n_tiles + 1), "left"),
        "row_ptr": np.r_[starts, M],
        "rows": order - tile_ptr[tile_of_row[order]],
    }


def _aligned_tiles(rowptr, max_rows=128):
    """Greedy packing of consecutive receivers into tiles of <= max_rows edges
    (and <= max_rows receivers); returns tile_seg or None if some in-degree
    exceeds a tile."""
    n_rec = rowptr.shape[0] - 1
    if n_rec == 0:
        return np.zeros(1, dtype=np.int64)
    if int((rowptr[1:] - rowptr[:-1]).max()) > max_rows:
        return None
    tile_seg = [0]
    i = 0
    while i < n_rec:
        j = int(np.searchsorted(rowptr, rowptr[i] + max_rows, side="right")) - 1
        j = max(min(j, i + max_rows, n_rec), i + 1)
        tile_seg.append(j)
        i = j
    return np.asarray(tile_seg, dtype=np.int64)


class InteractionNet(nn.Module):
    """See module docstring; arguments as interaction_net.py:19-47."""

    def __init__(self, edge_index, input_dim, update_edges=True, hidden_layers=1,
                 hidden_dim=None, edge_chunk_sizes=None, aggr_chunk_sizes=None, aggr="sum"):
        assert aggr in ("sum", "mean"), f"Unknown aggregation method: {aggr}"
        super().__init__()
        self.aggr = aggr
        if hidden_dim is None:
            hidden_dim = input_dim

        # interaction_net.py:55-62 -- integer arithmetic kept bit-exact
        edge_index = edge_index - edge_index.min(dim=1, keepdim=True)[0]
        self.num_rec = edge_index[1].max() + 1
        edge_index[0] = edge_index[0] + self.num_rec
        self.register_buffer("edge_index", edge_index, persistent=False)

        edge_mlp_recipe = [3 * input_dim] + [hidden_dim] * (hidden_layers + 1)
        aggr_mlp_recipe = [2 * input_dim] + [hidden_dim] * (hidden_layers + 1)
        if edge_chunk_sizes is None:
            self.edge_mlp = utils.make_mlp(edge_mlp_recipe)
        else:
            self.edge_mlp = SplitMLPs(
                [utils.make_mlp(edge_mlp_recipe) for _ in edge_chunk_sizes], edge_chunk_sizes)
        if aggr_chunk_sizes is None:
            self.aggr_mlp = utils.make_mlp(aggr_mlp_recipe)
        else:
            self.aggr_mlp = SplitMLPs(
                [utils.make_mlp(aggr_mlp_recipe) for _ in aggr_chunk_sizes], aggr_chunk_sizes)
        if hidden_layers < 1:
            raise NotImplementedError("InteractionNet: hidden_layers must be >= 1 (the fused "
                                      "kernels compute Linear -> SiLU -> Linear blocks)")
        self.hidden_layers = hidden_layers
        self.update_edges = update_edges
        self._edge_chunk_sizes = list(edge_chunk_sizes) if edge_chunk_sizes is not None else None
        self._aggr_chunk_sizes = list(aggr_chunk_sizes) if aggr_chunk_sizes is not None else None
        self._num_rec_int = int(self.num_rec)  # no host read-back per call
        self._plan = None

    def _get_plan(self):
        dev = self.edge_index.device
        if self._plan is None or self._plan.device != dev:
            if dev.type != "cuda":
                raise RuntimeError(
                    "neural_lam_b200.InteractionNet runs on CUDA only (module is on "
                    f"{dev}); there is no CPU fallback")
            self._plan = _GraphPlan(self.edge_index, self._num_rec_int,
                                    self._edge_chunk_sizes, self._aggr_chunk_sizes)
        return self._plan

    def forward(self, send_rep, rec_rep, edge_rep):
        """send_rep (B, N_send, d), rec_rep (B, N_rec, d), edge_rep (B, M, d)
        [the batch dim is optional, as in the reference] ->
        rec_rep or (rec_rep, edge_rep) (interaction_net.py:86-115)."""
        squeeze = rec_rep.dim() == 2
        if squeeze:
            send_rep, rec_rep, edge_rep = (t.unsqueeze(0) for t in (send_rep, rec_rep, edge_rep))
        if self.hidden_layers != 1:  # deeper MLPs: chained launches (ops.interaction_net_deep)
            out = ops.interaction_net_deep(self._get_plan(), self.edge_mlp, self.aggr_mlp, send_rep,
                                           rec_rep, edge_rep, self.aggr, self.update_edges)
            if self.update_edges:
                return (out[0][0], out[1][0]) if squeeze else out
            return out[0] if squeeze else out
        We = ops.weights_of(self.edge_mlp)
        Wa = ops.weights_of(self.aggr_mlp)
        meta = {
            "plan": self._get_plan(),
            "edge_chunks": We.n_chunks,
            "aggr_chunks": Wa.n_chunks,
            "aggr": self.aggr,
            "update_edges": self.update_edges,
            "precision": ops.get_precision(),
            "edge_params": We.t,
            "aggr_params": Wa.t,
        }
        out = ops._InteractionNetFn.apply(meta, *We.t, *Wa.t, send_rep, rec_rep, edge_rep)
        sh = meta.get("_sh", {})
        if self.update_edges:
            rec_out, edge_out = out
            ops.attach_shadow(rec_out, sh.get("rec_out"))
            ops.attach_shadow(edge_out, sh.get("new_edge"))
            return (rec_out[0], edge_out[0]) if squeeze else (rec_out, edge_out)
        ops.attach_shadow(out, sh.get("rec_out"))
        return out[0] if squeeze else out


class SplitMLPs(nn.Module):
    """Feeds chunks of the input (split along dim -2) through separate MLPs
    (interaction_net.py:134-163); one kernel launch, per-tile weight set."""

    def __init__(self, mlps, chunk_sizes):
        super().__init__()
        assert len(mlps) == len(chunk_sizes), "Number of MLPs must match the number of chunks"
        self.mlps = nn.ModuleList(mlps)
        self.chunk_sizes = chunk_sizes
        self._tiles = None

    def _tile_table(self, device):
        if self._tiles is None or self._tiles.tile_ptr.device != device:
            self._tiles = ops.TileTable(list(self.chunk_sizes), device)
        return self._tiles

    def forward(self, x):
        return ops.split_mlp_forward(self, x)
