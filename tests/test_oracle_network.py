"""Pin the network oracle: independent torch.nn restatement (nn.LSTM, nn.Conv2d), hand-checked padding, goldens."""

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import network_oracle as no
from orcai_b200 import runtime
from orcai_b200.weights import expected_shapes, synthetic_weights


@pytest.fixture(scope="module")
def PSW():
    P, S = runtime.bundled_parameters()
    return P, S, synthetic_weights(P, S, seed=1234)


def test_parameter_count(PSW):
    P, S, W = PSW
    assert sum(v.size for v in W.values()) == 996039
    assert set(W) == set(expected_shapes(P, S))


def test_maxpool_same_padding():
    x = torch.arange(1 * 1 * 6 * 3, dtype=torch.float32).reshape(1, 1, 6, 3)
    y = no._maxpool_3x2_s2_same(x)
    assert y.shape == (1, 1, 3, 2)
    # even H (every level of orcai-V1): windows start at even indices, one row / column of -inf padding at the END only
    assert y[0, 0, 0, 0] == x[0, 0, 0:3, 0:2].max() and y[0, 0, 2, 1] == x[0, 0, 4:6, 2:3].max()
    # odd H would split TF's "same" padding 1 before / 1 after; the CUDA kernels reject such shapes
    y5 = no._maxpool_3x2_s2_same(torch.arange(15, dtype=torch.float32).reshape(1, 1, 5, 3))
    assert y5[0, 0, 0, 0] == 4.0
    y2 = no._maxpool_3x2_s2_same(torch.zeros(1, 1, 736, 171))
    assert y2.shape == (1, 1, 368, 86)


def test_lstm_matches_torch_nn_lstm(PSW):
    _, _, W = PSW
    rng = np.random.default_rng(2)
    x = rng.standard_normal((3, 46, 396)).astype(np.float32)
    ours = no._bilstm(torch.as_tensor(x), W, "lstm1", torch.float32).numpy()
    lstm = torch.nn.LSTM(396, 128, batch_first=True, bidirectional=True)
    with torch.no_grad():
        for sfx, d in (("", "forward"), ("_reverse", "backward")):
            getattr(lstm, "weight_ih_l0" + sfx).copy_(torch.as_tensor(W[f"lstm1/{d}/kernel"]).T)
            getattr(lstm, "weight_hh_l0" + sfx).copy_(torch.as_tensor(W[f"lstm1/{d}/recurrent_kernel"]).T)
            getattr(lstm, "bias_ih_l0" + sfx).copy_(torch.as_tensor(W[f"lstm1/{d}/bias"]))
            getattr(lstm, "bias_hh_l0" + sfx).zero_()
        ref = lstm(torch.as_tensor(x))[0].numpy()
    np.testing.assert_allclose(ours, ref, atol=2e-6)


def test_forward_shapes_and_float64_agreement(PSW):
    _, _, W = PSW
    rng = np.random.default_rng(5)
    x = rng.random((1, 736, 171), dtype=np.float32)
    out32, inter = no.forward(x, W, return_intermediates=True)
    assert out32.shape == (1, 46, 7)
    assert inter["conv0"].shape == (1, 16, 736, 171) and inter["block1"].shape == (1, 30, 368, 86)
    assert inter["block2"].shape == (1, 40, 184, 43) and inter["block3"].shape == (1, 50, 92, 22)
    assert inter["block4"].shape == (1, 60, 46, 11) and inter["final"].shape == (1, 36, 46, 11)
    out64 = no.forward(x, W, dtype=torch.float64)
    assert np.abs(out32 - out64).max() < 2e-5


def test_block_residual_is_pre_relu(PSW):
    """The residual branch of block 2 must see block 1's un-rectified sum (architectures.py:170-196)."""
    _, _, W = PSW
    rng = np.random.default_rng(9)
    x = rng.random((1, 736, 171), dtype=np.float32)
    _, inter = no.forward(x, W, return_intermediates=True)
    b1 = torch.as_tensor(inter["block1"])
    assert (b1 < 0).any()  # the sum is not rectified
    rk = torch.as_tensor(W["block2/res/kernel"]).permute(3, 2, 0, 1)
    res = F.conv2d(b1, rk, torch.as_tensor(W["block2/res/bias"]), stride=2)
    assert res.shape == (1, 40, 184, 43)


def test_golden_network(PSW, golden_dir):
    _, _, W = PSW
    g = np.load(golden_dir / "network_seed1234.npz")
    assert int(g["n_params"]) == 996039
    assert abs(sum(float(np.abs(v).sum()) for v in W.values()) - float(g["weight_checksum"])) < 1e-6 * float(g["weight_checksum"])
    x = np.random.default_rng(5).random((2, 736, 171), dtype=np.float32)
    out, inter = no.forward(x, W, return_intermediates=True)
    np.testing.assert_allclose(out, g["probs"], atol=2e-6)
    for k in ("conv0", "block1", "block4", "final", "lstm2"):
        np.testing.assert_allclose(inter[k].mean(), g[f"{k}_mean"], rtol=1e-4)
