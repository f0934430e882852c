"""CPU: the C-ABI library builds, loads and exports every symbol that
include/nlam_b200.h declares (no compute without a GPU), and the host logic
fails loudly instead of falling back."""
import os
import re

import pytest
import torch

import __graft_entry__ as entry
from helpers import ROOT


@pytest.fixture(scope="module")
def built():
    entry.build()
    from neural_lam_b200 import lib
    return lib


def test_header_symbols_exported(built):
    header = open(os.path.join(ROOT, "include", "nlam_b200.h")).read()
    declared = set(re.findall(r"\b(nlam_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(built.SYMBOLS), (declared ^ set(built.SYMBOLS))
    l = built.load()
    for name in declared:
        assert hasattr(l, name)
    assert l.nlam_version() >= 1


def test_struct_sizes_match_header(built):
    # LP64 layout of the by-pointer descriptors (catches ctypes/header drift)
    import ctypes
    assert ctypes.sizeof(built.Src) == 56
    assert ctypes.sizeof(built.MlpWeights) == 48
    assert ctypes.sizeof(built.Agg) == 48
    assert ctypes.sizeof(built.RowMlp) == 8 + 3 * 56 + 16 + 48 + 8 + 24 + 8 + 8 + 8 + 8 + 48 + 8 + 16
    assert ctypes.sizeof(built.SegSum) == 64


def test_no_cpu_fallback(built):
    from neural_lam_b200.interaction_net import InteractionNet
    ei = torch.tensor([[3, 4, 5], [0, 1, 2]])
    net = InteractionNet(ei, 16)
    assert torch.equal(net.edge_index, torch.tensor([[3, 4, 5], [0, 1, 2]]))
    assert int(net.num_rec) == 3
    x = torch.randn(1, 3, 16)
    with pytest.raises(RuntimeError, match="CUDA only"):
        net(x, x, x)


def test_state_dict_keys_match_reference_contract(built):
    """SURVEY.md Appendix C."""
    from neural_lam_b200.interaction_net import InteractionNet
    ei = torch.tensor([[3, 4, 5, 5], [0, 1, 2, 2]])
    keys = list(InteractionNet(ei, 8).state_dict())
    want = [f"{m}.{i}.{p}" for m in ("edge_mlp", "aggr_mlp") for i in (0, 2, 3)
            for p in ("weight", "bias")]
    assert keys == want
    net = InteractionNet(ei, 8, edge_chunk_sizes=[3, 1], aggr_chunk_sizes=[2, 1])
    keys = list(net.state_dict())
    assert "edge_mlp.mlps.1.2.weight" in keys and "aggr_mlp.mlps.0.3.bias" in keys
    assert "edge_index" not in keys  # non-persistent buffer


def test_kernel_family_selection(built):
    """nlam_rowmlp_path is host arithmetic (widths, shared-memory budgets): which kernel family
    takes an MLP in each precision mode, and that option fp32_split switches the fp32 mode
    between the split-operand tensor-core kernels and the FFMA kernels."""
    from neural_lam_b200 import ops
    fam = ops.kernel_family
    assert fam((64, 64, 64), 64, 64, "bf16") == 1 and fam((128,) * 3, 128, 128, "bf16") == 1
    # fp32 on the tensor cores: every d = 64 shape of the models, narrow embedder inputs, the
    # d_out = 17 output map; hi + lo tiles of a d = 128 MLP do not fit shared memory -> FFMA
    assert fam((64, 64, 64), 64, 64, "fp32") == 2 and fam((64, 64), 64, 64, "fp32") == 2
    assert fam((3,), 64, 64, "fp32") == 2 and fam((64,), 64, 17, "fp32") == 2
    assert fam((16, 16, 16), 16, 16, "fp32") == 2
    assert fam((128,) * 3, 128, 128, "fp32") == 0 and fam((128, 128), 128, 128, "fp32") == 0
    l = built.load()
    try:
        assert l.nlam_set_option(b"fp32_split", 0) == 0
        assert fam((64, 64, 64), 64, 64, "fp32") == 0
    finally:
        l.nlam_set_option(b"fp32_split", 1)
    assert l.nlam_set_option(b"no_such_option", 1) == 1


def test_documented_options_are_accepted(built):
    """Every option name the header documents for nlam_set_option is accepted (set to its
    documented default here), and nothing else is."""
    defaults = {"fwd_mc": -1, "dgrad_mc": 0, "bwd_fused": -1, "pdl": 1, "tma": 0, "bwd_nh": 2,
                "wide128": 1, "fp32_split": 1, "small512": 1, "bwd_spread": 2}
    header = open(os.path.join(ROOT, "include", "nlam_b200.h")).read()
    doc = header[header.index("Kernel-selection knobs"):header.index("int nlam_set_option")]
    named = set(re.findall(r'"([a-z0-9_]+)"', doc))
    assert named == set(defaults), named ^ set(defaults)
    l = built.load()
    for name, value in defaults.items():
        assert l.nlam_set_option(name.encode(), value) == 0, name
    assert l.nlam_set_option(b"bwd_spred", 1) == 1
    assert b"unknown option" in l.nlam_last_error()
