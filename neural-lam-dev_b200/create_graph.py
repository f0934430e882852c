"""Graph builder for regular grids: writes the on-disk graph directory that
`utils.load_graph` reads (hot-path INPUT format).

Produces the same files, tensors, node numbering and EDGE ORDER as the
reference's networkx/PyG pipeline (/root/reference/neural_lam/create_graph.py:
157-535), but derived in closed form with numpy instead of replaying graph
mutations:

* a level is an n x n 8-neighbour lattice, nodes numbered i-major; the edge
  list is "for every node in index order, its neighbours in the order
  [W, E, S, N, SW, NE, SE, NW]" in (i, j) terms: (i-1,j), (i+1,j), (i,j-1),
  (i,j+1), (i-1,j-1), (i+1,j+1), (i+1,j-1), (i-1,j+1)  (that is the adjacency
  insertion order produced by create_graph.py:111-147);
* multiscale: coarser levels are re-labelled onto the level-0 nodes
  [1::3, 1::3] and appended after each node's finer-level neighbours
  (create_graph.py:371-405);
* hierarchical: 1-nearest-neighbour down edges sorted by (upper node, lower
  node); up edges are the same list flipped (create_graph.py:264-369);
* g2m: grid nodes within 0.67*dm of a level-0 mesh node, sorted by
  (grid node, mesh node); m2g: the 4 nearest level-0 mesh nodes of every grid
  node, sorted by (mesh node, grid node) (create_graph.py:419-525).
  scipy's KDTree answers the neighbour queries exactly as in the reference.

Edge features are [len, dx, dy] with vdiff = pos[sender] - pos[receiver]
(for up edges the reference keeps the DOWN edge's features, create_graph.py:
336-340), saved as float32; `load_graph` normalises them later.
"""
import os

import numpy as np
import scipy.spatial
import torch

# neighbour offsets (di, dj) in adjacency order, see module docstring
_NEIGH = np.array(
    [(-1, 0), (1, 0), (0, -1), (0, 1), (-1, -1), (1, 1), (1, -1), (-1, 1)],
    dtype=np.int64,
)
DM_SCALE = 0.67  # create_graph.py:424


def _level_positions(xy, n):
    """Mesh node coordinates of an n x n level (create_graph.py:112-121)."""
    xm, xM = np.amin(xy[:, :, 0][:, 0]), np.amax(xy[:, :, 0][:, 0])
    ym, yM = np.amin(xy[:, :, 1][0, :]), np.amax(xy[:, :, 1][0, :])
    dx = (xM - xm) / n
    dy = (yM - ym) / n
    lx = np.linspace(xm + dx / 2, xM - dx / 2, n)
    ly = np.linspace(ym + dy / 2, yM - dy / 2, n)
    pos = np.empty((n, n, 2), dtype=np.float64)
    pos[:, :, 0] = lx[:, None]
    pos[:, :, 1] = ly[None, :]
    return pos  # pos[i, j]


def _lattice_edges(n):
    """(src_i, src_j, dst_i, dst_j, slot) for all directed lattice edges, in
    node-major / adjacency order."""
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    ii = ii.reshape(-1, 1)
    jj = jj.reshape(-1, 1)
    di = ii + _NEIGH[None, :, 0]
    dj = jj + _NEIGH[None, :, 1]
    ok = (di >= 0) & (di < n) & (dj >= 0) & (dj < n)
    si = np.broadcast_to(ii, ok.shape)[ok]
    sj = np.broadcast_to(jj, ok.shape)[ok]
    return si, sj, di[ok], dj[ok]


def _edge_feats(pos_u, pos_v):
    vdiff = pos_u - pos_v
    length = np.sqrt(np.sum(vdiff**2, axis=1))
    return torch.from_numpy(
        np.concatenate((length[:, None], vdiff), axis=1)
    ).to(torch.float32)


def build_graph(xy, n_max_levels=None, hierarchical=False):
    """Compute all graph tensors for grid coordinates `xy` (Nx, Ny, 2).

    Returns a dict with the tensors/lists that `save_graph` writes.
    """
    xy = np.asarray(xy, dtype=np.float64)
    Nx, Ny = xy.shape[:2]
    pos_max = torch.max(torch.abs(torch.tensor(xy)))

    nx = 3
    nlev = int(np.log(max(xy.shape[:2])) / np.log(nx))
    nleaf = nx**nlev
    mesh_levels = nlev - 1
    if n_max_levels:
        mesh_levels = min(mesh_levels, n_max_levels)
    sizes = [int(nleaf / (nx**lev)) for lev in range(1, mesh_levels + 1)]
    level_pos = [_level_positions(xy, n) for n in sizes]

    out = {}
    if hierarchical:
        n_nodes = np.array([n * n for n in sizes])
        first = np.concatenate((np.zeros(1, dtype=int), np.cumsum(n_nodes[:-1])))
        m2m_ei, m2m_f, mesh_pos = [], [], []
        for n, pos, start in zip(sizes, level_pos, first):
            si, sj, di, dj = _lattice_edges(n)
            ei = np.stack((si * n + sj, di * n + dj)) + start
            m2m_ei.append(torch.from_numpy(ei))
            m2m_f.append(_edge_feats(pos[si, sj], pos[di, dj]))
            mesh_pos.append(torch.from_numpy(pos.reshape(-1, 2)).to(torch.float32))
        up_ei, down_ei, updown_f = [], [], []
        for lvl in range(mesh_levels - 1):
            lo = level_pos[lvl].reshape(-1, 2)
            hi = level_pos[lvl + 1].reshape(-1, 2)
            nearest = scipy.spatial.KDTree(hi).query(lo, 1)[1]
            order = np.lexsort((np.arange(lo.shape[0]), nearest))
            u = nearest[order]  # upper-level (sender of down edge)
            v = order  # lower-level node
            start = first[lvl]
            down = np.stack((u + start + lo.shape[0], v + start))
            down_ei.append(torch.from_numpy(down))
            up_ei.append(torch.from_numpy(down[::-1].copy()))
            updown_f.append(_edge_feats(hi[u], lo[v]))
        out["mesh_up_edge_index"] = up_ei
        out["mesh_down_edge_index"] = down_ei
        out["mesh_up_features"] = updown_f
        out["mesh_down_features"] = [f.clone() for f in updown_f]
        n_mesh_total = int(n_nodes.sum())
    else:
        n0 = sizes[0]
        src, dst, feats, lev_id, slot = [], [], [], [], []
        stride = 1
        offset = 0
        for lev, (n, pos) in enumerate(zip(sizes, level_pos)):
            # level-`lev` node (i, j) lives on level-0 node (offset+stride*i, ...)
            si, sj, di, dj = _lattice_edges(n)
            s0 = (offset + stride * si) * n0 + (offset + stride * sj)
            d0 = (offset + stride * di) * n0 + (offset + stride * dj)
            src.append(s0)
            dst.append(d0)
            feats.append(_edge_feats(pos[si, sj], pos[di, dj]))
            lev_id.append(np.full(s0.shape, lev))
            slot.append(np.arange(s0.shape[0]))
            offset = offset + stride  # [1::3] of the previous level
            stride = stride * nx
        src = np.concatenate(src)
        dst = np.concatenate(dst)
        feats = torch.cat(feats)
        # node-major; within a node finer levels first, then adjacency order
        order = np.lexsort(
            (np.concatenate(slot), np.concatenate(lev_id), src)
        )
        m2m_ei = [torch.from_numpy(np.stack((src[order], dst[order])))]
        m2m_f = [feats[torch.from_numpy(order)]]
        # merged nodes take the COARSEST level's coordinates (compose() keeps
        # the last writer, create_graph.py:384)
        pos0 = level_pos[0].copy()
        offset, stride = 0, 1
        for lev in range(1, len(sizes)):
            offset = offset + stride
            stride = stride * nx
            n = sizes[lev]
            idx = offset + stride * np.arange(n)
            pos0[np.ix_(idx, idx)] = level_pos[lev]
        level_pos = [pos0]
        mesh_pos = [torch.from_numpy(pos0.reshape(-1, 2)).to(torch.float32)]
        n_mesh_total = n0 * n0

    out["m2m_edge_index"] = m2m_ei
    out["m2m_features"] = m2m_f
    out["mesh_features"] = [p / pos_max for p in mesh_pos]

    # ---- grid <-> mesh (bottom level only) ----
    vm_xy = level_pos[0].reshape(-1, 2)
    dm = np.sqrt(np.sum((level_pos[0][1, 0] - level_pos[0][0, 0]) ** 2))
    # grid node k = a * Nx + b  <->  xy[b, a]   (create_graph.py:437-456)
    vg_xy = np.transpose(xy, (1, 0, 2)).reshape(-1, 2)
    kdt_g = scipy.spatial.KDTree(vg_xy)
    neigh = kdt_g.query_ball_point(vm_xy, dm * DM_SCALE)
    cnt = np.array([len(x) for x in neigh])
    v = np.repeat(np.arange(vm_xy.shape[0]), cnt)
    u = np.concatenate([np.asarray(x, dtype=np.int64) for x in neigh])
    order = np.lexsort((v, u))
    u, v = u[order], v[order]
    out["g2m_edge_index"] = torch.from_numpy(np.stack((u + n_mesh_total, v)))
    out["g2m_features"] = _edge_feats(vg_xy[u], vm_xy[v])

    kdt_m = scipy.spatial.KDTree(vm_xy)
    nn4 = kdt_m.query(vg_xy, 4)[1]  # (N_grid, 4)
    v = np.repeat(np.arange(vg_xy.shape[0]), 4)
    u = nn4.reshape(-1)
    order = np.lexsort((v, u))
    u, v = u[order], v[order]
    out["m2g_edge_index"] = torch.from_numpy(np.stack((u, v + n_mesh_total)))
    out["m2g_features"] = _edge_feats(vm_xy[u], vg_xy[v])
    return out


def save_graph(graph, graph_dir_path):
    """Write the `.pt` files in the layout utils.load_graph expects
    (/root/reference/neural_lam/utils.py:36-188; README.md:470-512)."""
    os.makedirs(graph_dir_path, exist_ok=True)
    for key in ("m2m_edge_index", "m2m_features", "mesh_features"):
        torch.save(graph[key], os.path.join(graph_dir_path, f"{key}.pt"))
    for key in ("g2m_edge_index", "g2m_features", "m2g_edge_index", "m2g_features"):
        torch.save(graph[key], os.path.join(graph_dir_path, f"{key}.pt"))
    if "mesh_up_edge_index" in graph:
        for key in (
            "mesh_up_edge_index",
            "mesh_down_edge_index",
            "mesh_up_features",
            "mesh_down_features",
        ):
            torch.save(graph[key], os.path.join(graph_dir_path, f"{key}.pt"))


def create_graph(graph_dir_path, xy, n_max_levels=None, hierarchical=False,
                 create_plot=False):
    """Same call signature as the reference's create_graph
    (create_graph.py:157-163); plotting is out of scope."""
    save_graph(build_graph(xy, n_max_levels, hierarchical), graph_dir_path)


def create_graph_from_datastore(datastore, output_root_path, n_max_levels=None,
                                hierarchical=False, create_plot=False):
    """create_graph.py:538-558."""
    xy = datastore.get_xy(category="state", stacked=False)
    create_graph(output_root_path, xy, n_max_levels, hierarchical, create_plot)
