"""torch-CPU restatement of the orcai-V1 forward pass (oracle; see oracle/__init__.py).

Follows the Keras graph of ``src/orcAI/architectures.py:120-241`` (ResNetLSTM) in
inference mode with Keras 3.10 / TF 2.19 layer semantics:

* Conv2D / SeparableConv2D are cross-correlations, "same" padding, NHWC.
* BatchNormalization uses moving statistics, epsilon = 1e-3.
* MaxPooling2D((3,2), strides 2, "same") pads only at the END of H (and of W when odd) with -inf.
* The residual 1x1 stride-2 "same" convolution samples even indices, no padding.
* ``previous_block_activation`` is taken before each block's leading ReLU.
* Reshape flattens (W, C) as w*C + c.
* LSTM gates are ordered i, f, c, o; sigmoid recurrent activation, tanh cell/output.
* Bidirectional concatenates [forward, time-reversed backward].

Weights are a dict name -> numpy array in Keras variable layouts (see
``orcai_b200/weights.py`` for the naming scheme).
"""

from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3


def _t(w, dtype):
    return torch.as_tensor(np.asarray(w), dtype=dtype)


def _bn(x, W, prefix, dtype):
    g, b, m, v = (_t(W[f"{prefix}/{k}"], dtype) for k in ("gamma", "beta", "moving_mean", "moving_variance"))
    sh = (1, -1, 1, 1) if x.dim() == 4 else (1, 1, -1)
    return (x - m.view(sh)) / torch.sqrt(v.view(sh) + BN_EPS) * g.view(sh) + b.view(sh)


def _conv_same(x, k_hwio, bias, dtype):
    w = _t(k_hwio, dtype).permute(3, 2, 0, 1).contiguous()  # (O, I, kh, kw)
    return F.conv2d(x, w, None if bias is None else _t(bias, dtype), padding=(w.shape[2] // 2, w.shape[3] // 2))


def _sepconv(x, W, prefix, dtype):
    dw = _t(W[f"{prefix}/depthwise_kernel"], dtype)  # (3,3,C,1)
    C = dw.shape[2]
    x = F.conv2d(x, dw.permute(2, 3, 0, 1).contiguous(), None, padding=1, groups=C)
    pw = _t(W[f"{prefix}/pointwise_kernel"], dtype)  # (1,1,C,O)
    return F.conv2d(x, pw.permute(3, 2, 0, 1).contiguous(), _t(W[f"{prefix}/bias"], dtype))


def _maxpool_3x2_s2_same(x):
    H, Wd = x.shape[2], x.shape[3]
    oh, ow = -(-H // 2), -(-Wd // 2)
    ph = max((oh - 1) * 2 + 3 - H, 0)
    pw = max((ow - 1) * 2 + 2 - Wd, 0)
    # TF "same": pad_before = total // 2 (== 0 here), the rest after
    x = F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2), value=float("-inf"))
    return F.max_pool2d(x, kernel_size=(3, 2), stride=(2, 2))


def _lstm_dir(x, W, prefix, reverse, dtype):
    """x (B, Tn, I) -> (B, Tn, U); Keras LSTM cell, gate order i, f, c, o."""
    K = _t(W[f"{prefix}/kernel"], dtype)  # (I, 4U)
    R = _t(W[f"{prefix}/recurrent_kernel"], dtype)  # (U, 4U)
    b = _t(W[f"{prefix}/bias"], dtype)  # (4U,)
    U = R.shape[0]
    B, Tn, _ = x.shape
    h = torch.zeros(B, U, dtype=dtype)
    c = torch.zeros(B, U, dtype=dtype)
    xz = x @ K + b
    out = torch.empty(B, Tn, U, dtype=dtype)
    order = range(Tn - 1, -1, -1) if reverse else range(Tn)
    for t in order:
        z = xz[:, t] + h @ R
        i, f, g, o = z[:, :U], z[:, U : 2 * U], z[:, 2 * U : 3 * U], z[:, 3 * U :]
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[:, t] = h
    return out


def _bilstm(x, W, prefix, dtype):
    return torch.cat(
        [_lstm_dir(x, W, f"{prefix}/forward", False, dtype), _lstm_dir(x, W, f"{prefix}/backward", True, dtype)], dim=-1
    )


@torch.no_grad()
def forward(snippets: np.ndarray, W: dict, n_blocks: int = 4, dtype=torch.float32, return_intermediates: bool = False):
    """snippets (B, H, Wf) or (B, H, Wf, 1) in [0,1] -> probabilities (B, H // 2**n_blocks, n_labels)."""
    x = torch.as_tensor(np.asarray(snippets), dtype=dtype)
    if x.dim() == 4:
        x = x[..., 0]
    x = x[:, None]  # NCHW with C = 1, H = time, W = frequency
    inter = {}
    x = _conv_same(x, W["conv0/kernel"], W["conv0/bias"], dtype)
    x = torch.relu(_bn(x, W, "bn0", dtype))
    inter["conv0"] = x
    prev = x
    for b in range(1, n_blocks + 1):
        p = f"block{b}"
        x = torch.relu(x)
        x = _sepconv(x, W, f"{p}/sep1", dtype)
        x = torch.relu(_bn(x, W, f"{p}/bn1", dtype))
        x = _sepconv(x, W, f"{p}/sep2", dtype)
        x = _bn(x, W, f"{p}/bn2", dtype)
        x = _maxpool_3x2_s2_same(x)
        rk = _t(W[f"{p}/res/kernel"], dtype).permute(3, 2, 0, 1).contiguous()
        res = F.conv2d(prev, rk, _t(W[f"{p}/res/bias"], dtype), stride=2)
        x = x + res
        prev = x
        inter[p] = x
    x = _sepconv(x, W, "final/sep", dtype)
    x = torch.relu(_bn(x, W, "final/bn", dtype))
    inter["final"] = x
    B, C, H, Wd = x.shape
    x = x.permute(0, 2, 3, 1).reshape(B, H, Wd * C)  # feature index = w*C + c
    x = _bilstm(x, W, "lstm1", dtype)
    inter["lstm1"] = x
    x = _bilstm(x, W, "lstm2", dtype)
    inter["lstm2"] = x
    x = torch.relu(x @ _t(W["dense1/kernel"], dtype) + _t(W["dense1/bias"], dtype))
    x = _bn(x, W, "bn_dense", dtype)
    x = torch.sigmoid(x @ _t(W["dense2/kernel"], dtype) + _t(W["dense2/bias"], dtype))
    out = x.numpy()
    if return_intermediates:
        return out, {k: v.numpy() for k, v in inter.items()}
    return out
