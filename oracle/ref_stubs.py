"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Stub modules that let the reference files under /root/reference run VERBATIM in
this container, where its third-party dependencies (torch_geometric,
pytorch_lightning, xarray, cartopy, matplotlib, tueplots, dataclass_wizard,
mllam_data_prep, dask, parse, plotly) are not installed and cannot be
(no network).  Recipe: SURVEY.md Appendix F.

The ONLY arithmetic restated here is PyG 2.3.1 `MessagePassing.propagate`
(`index_select` along dim -2 for `x_j`/`x_i`, `scatter_add_` for sum, sum /
count.clamp(min=1) for mean) and `torch_geometric.nn.Sequential`
(child naming `module_{i}`); everything else that executes is the reference's
own code (call sites: /root/reference/neural_lam/interaction_net.py:10,49,
103-105,130; graph_lam.py:51-57; hi_lam_parallel.py:47-53).

/root/reference does not exist on the GPU box: this file is only used here,
by oracle/make_golden.py, to (a) validate oracle/port.py against the real
reference and (b) generate the golden vectors committed under tests/golden/.
"""
import inspect
import sys
import types
from unittest import mock

import torch
from torch import nn

REFERENCE_ROOT = "/root/reference"


# ---------------------------------------------------------------- torch_geometric
class _MessagePassing(nn.Module):
    """Restates PyG 2.3.1 MessagePassing for node_dim=-2, flow=source_to_target."""

    def __init__(self, aggr="add", **kwargs):
        super().__init__()
        assert aggr in ("add", "sum", "mean")
        self.aggr = aggr
        self.node_dim = -2
        self._msg_params = [
            p for p in inspect.signature(self.message).parameters
        ]

    def propagate(self, edge_index, size=None, **kwargs):
        msg_kwargs = {}
        for name in self._msg_params:
            if name.endswith("_j"):
                msg_kwargs[name] = kwargs[name[:-2]].index_select(
                    self.node_dim, edge_index[0]
                )
            elif name.endswith("_i"):
                msg_kwargs[name] = kwargs[name[:-2]].index_select(
                    self.node_dim, edge_index[1]
                )
            else:
                msg_kwargs[name] = kwargs[name]
        out = self.message(**msg_kwargs)
        dim_size = kwargs["x"].shape[self.node_dim]
        out = self.aggregate(out, edge_index[1], None, dim_size)
        return self.update(out)

    def message(self, x_j):  # pragma: no cover - always overridden
        return x_j

    def aggregate(self, inputs, index, ptr=None, dim_size=None):
        dim_size = int(dim_size)
        shape = list(inputs.shape)
        shape[self.node_dim] = dim_size
        bshape = [1] * inputs.dim()
        bshape[self.node_dim] = -1
        idx = index.view(bshape).expand_as(inputs)
        out = inputs.new_zeros(shape).scatter_add_(self.node_dim, idx, inputs)
        if self.aggr == "mean":
            count = inputs.new_zeros(dim_size).scatter_add_(
                0, index, inputs.new_ones(index.shape[0])
            )
            out = out / count.clamp(min=1).view(bshape)
        return out

    def update(self, inputs):
        return inputs


class _PygSequential(nn.Module):
    """Restates torch_geometric.nn.Sequential(input_args, [(module, desc), ...])."""

    def __init__(self, input_args, modules):
        super().__init__()
        self._in = [a.strip() for a in input_args.split(",")]
        self._descs = []
        for i, (mod, desc) in enumerate(modules):
            ins, outs = desc.split("->")
            self._descs.append(
                (
                    [a.strip() for a in ins.split(",")],
                    [a.strip() for a in outs.split(",")],
                )
            )
            self.add_module(f"module_{i}", mod)

    def forward(self, *args):
        env = dict(zip(self._in, args))
        out = None
        for i, (ins, outs) in enumerate(self._descs):
            out = getattr(self, f"module_{i}")(*[env[a] for a in ins])
            if not isinstance(out, tuple):
                out = (out,)
            env.update(zip(outs, out))
        return out if len(out) > 1 else out[0]


def _from_networkx(G):
    """Restates torch_geometric.utils.convert.from_networkx (2.3.1) for the
    attributes create_graph.py uses (node 'pos'; edge 'len','vdiff')."""
    import networkx as nx
    import numpy as np

    G = G.to_directed() if not nx.is_directed(G) else G
    mapping = dict(zip(G.nodes(), range(G.number_of_nodes())))
    edge_index = torch.empty((2, G.number_of_edges()), dtype=torch.long)
    for i, (src, dst) in enumerate(G.edges()):
        edge_index[0, i] = mapping[src]
        edge_index[1, i] = mapping[dst]
    data = types.SimpleNamespace()
    data.edge_index = edge_index

    def _stack(vals):
        arr = np.array(vals)
        return torch.from_numpy(arr)

    node_attrs = {}
    for _, feat in G.nodes(data=True):
        for k, v in feat.items():
            node_attrs.setdefault(k, []).append(v)
    for k, v in node_attrs.items():
        setattr(data, k, _stack(v))
    edge_attrs = {}
    for _, _, feat in G.edges(data=True):
        for k, v in feat.items():
            edge_attrs.setdefault(k, []).append(v)
    for k, v in edge_attrs.items():
        setattr(data, k, _stack(v))

    def clone():
        c = types.SimpleNamespace(**{
            k: (v.clone() if torch.is_tensor(v) else v)
            for k, v in vars(data).items() if k != "clone"
        })
        return c

    data.clone = clone
    return data


# ---------------------------------------------------------------- lightning
class _LightningModule(nn.Module):
    def save_hyperparameters(self, *a, **k):
        pass

    def log_dict(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass

    def all_gather(self, t, *a, **k):
        return t.unsqueeze(0)

    @property
    def trainer(self):
        return types.SimpleNamespace(is_global_zero=True, sanity_checking=False)


class _LightningDataModule:
    def __init__(self, *a, **k):
        pass


def _mock_module(name):
    m = mock.MagicMock(name=name)
    m.__path__ = []
    m.__spec__ = None
    m.__name__ = name
    return m


def install_stubs():
    """Register the stub modules in sys.modules (idempotent)."""
    if "torch_geometric" in sys.modules and getattr(
        sys.modules["torch_geometric"], "_nlam_stub", False
    ):
        return

    pyg = types.ModuleType("torch_geometric")
    pyg._nlam_stub = True
    pyg.__path__ = []
    pyg_nn = types.ModuleType("torch_geometric.nn")
    pyg_nn.MessagePassing = _MessagePassing
    pyg_nn.Sequential = _PygSequential
    pyg.nn = pyg_nn
    pyg_utils = types.ModuleType("torch_geometric.utils")
    pyg_utils.__path__ = []
    pyg_convert = types.ModuleType("torch_geometric.utils.convert")
    pyg_convert.from_networkx = _from_networkx
    pyg_utils.convert = pyg_convert
    pyg.utils = pyg_utils
    sys.modules["torch_geometric"] = pyg
    sys.modules["torch_geometric.nn"] = pyg_nn
    sys.modules["torch_geometric.utils"] = pyg_utils
    sys.modules["torch_geometric.utils.convert"] = pyg_convert

    pl = types.ModuleType("pytorch_lightning")
    pl.__path__ = []
    pl.LightningModule = _LightningModule
    pl.LightningDataModule = _LightningDataModule
    pl.Trainer = mock.MagicMock()
    pl.callbacks = mock.MagicMock()
    sys.modules["pytorch_lightning"] = pl
    for sub in ("utilities", "loggers", "callbacks"):
        sys.modules[f"pytorch_lightning.{sub}"] = _mock_module(
            f"pytorch_lightning.{sub}"
        )
    sys.modules["lightning_fabric"] = _mock_module("lightning_fabric")
    sys.modules["lightning_fabric.utilities"] = _mock_module(
        "lightning_fabric.utilities"
    )
    sys.modules["lightning_fabric.utilities.seed"] = _mock_module(
        "lightning_fabric.utilities.seed"
    )

    tue = types.ModuleType("tueplots")
    tue.__path__ = []
    bundles = types.ModuleType("tueplots.bundles")
    bundles.neurips2023 = lambda **k: {"figure.figsize": (5.5, 3.4)}
    figsizes = types.ModuleType("tueplots.figsizes")
    figsizes.neurips2023 = lambda **k: {"figure.figsize": (5.5, 3.4)}
    tue.bundles = bundles
    tue.figsizes = figsizes
    sys.modules["tueplots"] = tue
    sys.modules["tueplots.bundles"] = bundles
    sys.modules["tueplots.figsizes"] = figsizes

    mpl = _mock_module("matplotlib")

    def rc_context(*a, **k):
        def deco(f):
            return f

        return deco

    mpl.rc_context = rc_context
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = _mock_module("matplotlib.pyplot")
    mpl.pyplot = sys.modules["matplotlib.pyplot"]
    for name in (
        "xarray",
        "xarray.core",
        "xarray.core.dataarray",
        "cartopy",
        "cartopy.crs",
        "mllam_data_prep",
        "dask",
        "dask.array",
        "dask.delayed",
        "parse",
        "plotly",
        "plotly.graph_objects",
    ):
        sys.modules[name] = _mock_module(name)

    dw = types.ModuleType("dataclass_wizard")
    dw.__path__ = []

    class JSONWizard:
        class Meta:
            pass

        def __init_subclass__(cls, **kw):
            super().__init_subclass__()

    class YAMLWizard:
        def __init_subclass__(cls, **kw):
            super().__init_subclass__()

    class JSONFileWizard:
        pass

    dw.JSONWizard = JSONWizard
    dw.YAMLWizard = YAMLWizard
    dw.JSONFileWizard = JSONFileWizard
    dw_err = types.ModuleType("dataclass_wizard.errors")

    class UnknownJSONKey(Exception):
        pass

    dw_err.UnknownJSONKey = UnknownJSONKey
    dw.errors = dw_err
    sys.modules["dataclass_wizard"] = dw
    sys.modules["dataclass_wizard.errors"] = dw_err


STAGED_ARCHIVE = __import__("os").path.join(
    __import__("os").path.dirname(__import__("os").path.abspath(__file__)), "_ref",
    "neural_lam_ref.zip")


def reference_location():
    """Where the unmodified reference can be imported from: its directory in the build
    container, else the archive oracle/stage_ref.py packed from it (GPU box), else None."""
    import os

    if os.path.isdir(os.path.join(REFERENCE_ROOT, "neural_lam")):
        return REFERENCE_ROOT
    if os.path.exists(STAGED_ARCHIVE):
        return STAGED_ARCHIVE
    return None


def import_reference():
    """Import the unmodified reference package; returns the module."""
    where = reference_location()
    if where is None:
        raise RuntimeError(
            f"neither {REFERENCE_ROOT} nor {STAGED_ARCHIVE} is present; run "
            "oracle/stage_ref.py in the build container, or use oracle.port instead"
        )
    install_stubs()
    if where not in sys.path:
        sys.path.insert(0, where)
    import neural_lam  # noqa: F401

    return neural_lam
