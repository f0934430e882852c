"""CPU: ops.blocks_of cuts a make_mlp blueprint with h >= 1 hidden layers into the
[LN](W_b . SiLU(W_a . x + b_a) + b_b) pieces one fused-kernel launch computes
(utils.py:191-214, train_model.py:94).  Evaluating the pieces with plain torch must reproduce
the module's own forward -- this pins the identity-first-layer construction of the deeper
blocks, and the parameter bookkeeping of Weights, without a GPU."""
import pytest
import torch
import torch.nn.functional as F

import helpers  # noqa: F401
from neural_lam_b200 import ops, utils
from oracle import port


def _run_blocks(blocks, x):
    for W in blocks:
        w1, b1, w2, b2, g, b = W.t
        assert W.n_chunks == 1 and w1.shape == (W.d_hidden, W.k) and w2.shape == (W.d_out, W.d_hidden)
        x = F.silu(x @ w1.T + b1) @ w2.T + b2
        if g is not None:
            x = F.layer_norm(x, (W.d_out,), g, b, eps=1e-5)
    return x


@pytest.mark.parametrize("blueprint,ln", [([56, 64, 64], True), ([56, 64, 64, 64], True),
                                          ([64, 64, 64, 64, 17], False), ([3, 16, 16, 16, 16], True)])
def test_blocks_reproduce_the_sequential(blueprint, ln):
    torch.manual_seed(0)
    ref = port.make_mlp(blueprint, layer_norm=ln)  # the reference's blueprint, restated
    mlp = utils.make_mlp(blueprint, layer_norm=ln)
    mlp.load_state_dict(ref.state_dict())
    blocks = ops.blocks_of(mlp)
    h = len(blueprint) - 2
    assert len(blocks) == h
    # LayerNorm only on the last block; deeper blocks ride on an identity first layer
    assert [W.has_ln for W in blocks] == [False] * (h - 1) + [ln]
    for W in blocks[1:]:
        assert torch.equal(W.t[0], torch.eye(W.k)) and not W.t[1].any()
    n_params = sum(p.numel() for p in mlp.parameters())
    extra = sum(W.k * W.k + W.k for W in blocks[1:])  # the identity layers are not parameters
    assert sum(W.param_floats() for W in blocks) == n_params + extra
    x = torch.randn(5, 37, blueprint[0])
    with torch.no_grad():
        torch.testing.assert_close(_run_blocks(blocks, x), ref(x), rtol=1e-5, atol=1e-6)


def test_single_block_is_weights_of():
    mlp = utils.make_mlp([192, 64, 64], layer_norm=True)
    (W,), V = ops.blocks_of(mlp), ops.weights_of(mlp)
    assert all(a is b for a, b in zip(W.t, V.t))


def test_zero_hidden_layers_raises():
    with pytest.raises((ValueError, NotImplementedError)):
        utils.make_mlp([8, 8], layer_norm=True)
