#!/usr/bin/env python
"""CPU study of the 16-bit network path's rounding error (test infrastructure; uses the oracle).

Emulates the rounding points of the fused tensor-core path (folded fp16 weights, fp16 activations at conv0 / S1 / S2 /
block outputs, fp32 accumulation) inside the torch-CPU oracle graph and reports the probability error against the
fp32 oracle, for the plain seeded weights and for BatchNorm-calibrated weights (moving statistics set to the actual
activation statistics, as training would leave them).

    python tools/precision_study.py [--snippets 4]
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import network_oracle as no, postprocess_oracle as po, spectrogram_oracle as so  # noqa: E402
from orcai_b200 import runtime  # noqa: E402
from orcai_b200.synth import pcm16_to_float, synth_pcm16  # noqa: E402
from orcai_b200.weights import synthetic_weights  # noqa: E402


def calibrate_bn(W: dict, x: np.ndarray) -> dict:
    """Set every BatchNorm's moving statistics to the statistics its input has on `x` (sequentially, layer by layer)."""
    W = {k: np.array(v, copy=True) for k, v in W.items()}
    orig = no._bn

    def bn(t, Wd, prefix, dtype):
        dims = (0, 2, 3) if t.dim() == 4 else (0, 1)
        Wd[f"{prefix}/moving_mean"] = t.mean(dim=dims).numpy().astype(np.float32)
        Wd[f"{prefix}/moving_variance"] = t.var(dim=dims, unbiased=False).numpy().astype(np.float32)
        return orig(t, Wd, prefix, dtype)

    no._bn = bn
    try:
        no.forward(x, W)
    finally:
        no._bn = orig
    return W


def h16(t, fmt):
    return t.to(fmt).to(torch.float32)


@torch.no_grad()
def lstm_dir_16(x, W, prefix, reverse, fmt, proj16, rec16):
    f32 = torch.float32
    K = torch.as_tensor(W[f"{prefix}/kernel"], dtype=f32)
    R = torch.as_tensor(W[f"{prefix}/recurrent_kernel"], dtype=f32)
    b = torch.as_tensor(W[f"{prefix}/bias"], dtype=f32)
    U = R.shape[0]
    B, Tn, _ = x.shape
    h = torch.zeros(B, U); c = torch.zeros(B, U)
    xz = (h16(x, fmt) @ h16(K, fmt) + b) if proj16 else (x @ K + b)
    Rr = h16(R, fmt) if rec16 else R
    out = torch.empty(B, Tn, U)
    for t in (range(Tn - 1, -1, -1) if reverse else range(Tn)):
        hh = h16(h, fmt) if rec16 == 1 else h       # rec16 == 2: h split hi+lo (exact), weights fp16
        z = xz[:, t] + hh @ Rr
        i, f, g, o = z[:, :U], z[:, U:2 * U], z[:, 2 * U:3 * U], z[:, 3 * U:]
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[:, t] = h
    return out


def bilstm_16(x, W, prefix, fmt, proj16, rec16):
    return torch.cat([lstm_dir_16(x, W, f"{prefix}/forward", False, fmt, proj16, rec16),
                      lstm_dir_16(x, W, f"{prefix}/backward", True, fmt, proj16, rec16)], dim=-1)


def forward_16(x: np.ndarray, W: dict, fmt=torch.float16, conv0_fp32=False, s2_fp32=False, proj16=False, rec16=0):
    f32 = torch.float32
    T = lambda a: torch.as_tensor(np.asarray(a), dtype=f32)

    def bn_fold(prefix):
        g, b, m, v = (T(W[f"{prefix}/{k}"]).double() for k in ("gamma", "beta", "moving_mean", "moving_variance"))
        s = g / torch.sqrt(v + no.BN_EPS)
        return s, b - m * s

    def sep_folded(xin, sp, bnp):
        dw = T(W[f"{sp}/depthwise_kernel"]).double()[..., 0]        # (3,3,C)
        pw = T(W[f"{sp}/pointwise_kernel"]).double()[0, 0]          # (C,O)
        b = T(W[f"{sp}/bias"]).double()
        s, t = bn_fold(bnp)
        pws = (pw * s[None, :]).float()                              # fp32 like the host fold
        wt = h16(dw.float()[:, :, :, None] * pws[None, None], fmt)   # (3,3,C,O) folded, rounded
        bias = (b * s + t).float()
        return F.conv2d(xin, wt.permute(3, 2, 0, 1).contiguous(), bias, padding=1)

    x = torch.as_tensor(np.asarray(x), dtype=f32)[:, None]
    s, t = bn_fold("bn0")
    k0 = (T(W["conv0/kernel"]).double() * s[None, None, None, :]).float()
    b0 = (T(W["conv0/bias"]).double() * s + t).float()
    if conv0_fp32:
        y = F.conv2d(x, k0.permute(3, 2, 0, 1).contiguous(), b0, padding=1)
    else:
        y = F.conv2d(h16(x, fmt), h16(k0, fmt).permute(3, 2, 0, 1).contiguous(), b0, padding=1)
    y = h16(torch.relu(y), fmt)
    prev = y
    for b in range(1, 5):
        p = f"block{b}"
        a = h16(torch.relu(sep_folded(torch.relu(prev), f"{p}/sep1", f"{p}/bn1")), fmt)
        z = sep_folded(a, f"{p}/sep2", f"{p}/bn2")
        if not s2_fp32:
            z = h16(z, fmt)
        z = no._maxpool_3x2_s2_same(z)
        rk = h16(T(W[f"{p}/res/kernel"]), fmt).permute(3, 2, 0, 1).contiguous()
        res = F.conv2d(prev, rk, T(W[f"{p}/res/bias"]), stride=2)
        prev = h16(z + res, fmt)
    feat = torch.relu(sep_folded(prev, "final/sep", "final/bn"))
    B, C, H, Wd = feat.shape
    xx = feat.permute(0, 2, 3, 1).reshape(B, H, Wd * C)
    xx = bilstm_16(xx, W, "lstm1", fmt, proj16, rec16)
    xx = bilstm_16(xx, W, "lstm2", fmt, proj16, rec16)
    xx = torch.relu(xx @ T(W["dense1/kernel"]) + T(W["dense1/bias"]))
    xx = no._bn(xx, W, "bn_dense", f32)
    return torch.sigmoid(xx @ T(W["dense2/kernel"]) + T(W["dense2/bias"])).numpy(), feat.numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--snippets", type=int, default=4)
    a = ap.parse_args()
    P, S = runtime.bundled_parameters()
    pcm = synth_pcm16(30.0, seed=77, calls_per_minute=30.0)
    spec, _, _ = so.make_spectrogram(pcm16_to_float(pcm), P["spectrogram"])
    snips = po.cut_snippets(spec, 736)[: a.snippets]
    rnd = np.random.default_rng(5).random((a.snippets, 736, 171), dtype=np.float32)
    W0 = synthetic_weights(P, S, seed=1234)
    Wc = calibrate_bn(W0, po.cut_snippets(spec, 736)[a.snippets : a.snippets + 4])
    for wname, W in (("seed1234", W0), ("seed1234+BN-calibrated", Wc)):
        for xname, x in (("uniform-random", rnd), ("synthetic-audio", snips)):
            ref, inter = no.forward(x, W, return_intermediates=True)
            fscale = np.abs(inter["final"]).max()
            for fmt in (torch.float16,):
                for kw in ({}, {"conv0_fp32": True}, {"conv0_fp32": True, "proj16": True}, {"conv0_fp32": True, "proj16": True, "rec16": 1},
                           {"conv0_fp32": True, "proj16": True, "rec16": 2}):
                    out, feat = forward_16(x, W, fmt, **kw)
                    e = np.abs(out - ref)
                    fe = np.abs(feat - inter["final"]).max()
                    print(f"{wname:24s} {xname:16s} {str(fmt)[6:]:9s} {str(kw):42s} prob err max {e.max():.2e} mean {e.mean():.2e} | final feat max {fscale:8.2f} err {fe:.2e}", flush=True)


if __name__ == "__main__":
    main()
