#!/usr/bin/env python
"""CPU study: which 16-bit rounding points of the tensor-core path must go to meet 1e-3 (test infrastructure; uses the oracle).

Every GEMM of the network has a weight mode (fp16, or split fp16 hi+lo = exact to 2^-22) and every stored activation a
storage mode (fp16, or split hi+lo).  A plan names the layers whose WEIGHTS are split (`w`) and the activation points
that are kept SPLIT (`a`); everything else is fp16.  The script runs each plan over a whole recording and reports the
deviation of the per-snippet and of the overlap-averaged probabilities from the fp32 oracle.

    python tools/precision_plan.py [--minutes 10] [--bn-matched] [--plans name,name]
"""
from __future__ import annotations

import argparse
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from oracle import network_oracle as no, postprocess_oracle as po, spectrogram_oracle as so  # noqa: E402
from orcai_b200 import runtime  # noqa: E402
from orcai_b200.synth import pcm16_to_float, synth_pcm16  # noqa: E402
from orcai_b200.weights import synthetic_weights  # noqa: E402
import precision_study as ps  # noqa: E402

f16, f32 = torch.float16, torch.float32
GEMMS = ["conv0"] + [f"b{b}{s}" for b in range(1, 5) for s in ("s1", "s2", "res")] + ["final", "l1p", "l1r", "l2p", "l2r", "d1"]
ACTS = ["spec", "c0"] + [f"b{b}{s}" for b in range(1, 5) for s in ("S1", "S2", "out")] + ["feat", "h1", "o1", "h2", "o2"]


def h(t, on=True):
    return t.to(f16).to(f32) if on else t


@torch.no_grad()
def forward_plan(x, W, wsplit: set, asplit: set):
    """fp16 emulation; layers in `wsplit` use exact weights, activation points in `asplit` keep fp32-grade storage."""
    T = lambda a: torch.as_tensor(np.asarray(a), dtype=f32)
    wq = lambda w, name: w if name in wsplit else h(w)
    aq = lambda a, name: a if name in asplit else h(a)

    def bn_fold(prefix):
        g, b, m, v = (T(W[f"{prefix}/{k}"]).double() for k in ("gamma", "beta", "moving_mean", "moving_variance"))
        s = g / torch.sqrt(v + no.BN_EPS)
        return s, b - m * s

    def sep(a, sp, bnp, name):
        dw = T(W[f"{sp}/depthwise_kernel"]).double()[..., 0]
        pw = T(W[f"{sp}/pointwise_kernel"]).double()[0, 0]
        b = T(W[f"{sp}/bias"]).double()
        s, t = bn_fold(bnp)
        pws = (pw * s[None, :]).float()
        wt = wq(dw.float()[:, :, :, None] * pws[None, None], name)
        return F.conv2d(a, wt.permute(3, 2, 0, 1).contiguous(), (b * s + t).float(), padding=1)

    def lstm_dir(xx, prefix, reverse, pn, rn, hn):
        K, R, b = (T(W[f"{prefix}/{k}"]) for k in ("kernel", "recurrent_kernel", "bias"))
        U = R.shape[0]
        B, Tn, _ = xx.shape
        hh = torch.zeros(B, U); c = torch.zeros(B, U)
        xz = xx @ wq(K, pn) + b
        Rr = wq(R, rn)
        out = torch.empty(B, Tn, U)
        for t in (range(Tn - 1, -1, -1) if reverse else range(Tn)):
            z = xz[:, t] + aq(hh, hn) @ Rr
            i, f, g, o = z[:, :U], z[:, U:2 * U], z[:, 2 * U:3 * U], z[:, 3 * U:]
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            hh = torch.sigmoid(o) * torch.tanh(c)
            out[:, t] = hh
        return out

    xx = aq(torch.as_tensor(np.asarray(x), dtype=f32)[:, None], "spec")
    s, t = bn_fold("bn0")
    k0 = (T(W["conv0/kernel"]).double() * s[None, None, None, :]).float()
    b0 = (T(W["conv0/bias"]).double() * s + t).float()
    prev = aq(torch.relu(F.conv2d(xx, wq(k0, "conv0").permute(3, 2, 0, 1).contiguous(), b0, padding=1)), "c0")
    for b in range(1, 5):
        p = f"block{b}"
        a = aq(torch.relu(sep(torch.relu(prev), f"{p}/sep1", f"{p}/bn1", f"b{b}s1")), f"b{b}S1")
        z = aq(sep(a, f"{p}/sep2", f"{p}/bn2", f"b{b}s2"), f"b{b}S2")
        z = no._maxpool_3x2_s2_same(z)
        rk = wq(T(W[f"{p}/res/kernel"]), f"b{b}res").permute(3, 2, 0, 1).contiguous()
        res = F.conv2d(prev, rk, T(W[f"{p}/res/bias"]), stride=2)
        prev = aq(z + res, f"b{b}out")
    feat = aq(torch.relu(sep(prev, "final/sep", "final/bn", "final")), "feat")
    B, C, H, Wd = feat.shape
    xx = feat.permute(0, 2, 3, 1).reshape(B, H, Wd * C)
    xx = aq(torch.cat([lstm_dir(xx, "lstm1/forward", False, "l1p", "l1r", "h1"), lstm_dir(xx, "lstm1/backward", True, "l1p", "l1r", "h1")], -1), "o1")
    xx = aq(torch.cat([lstm_dir(xx, "lstm2/forward", False, "l2p", "l2r", "h2"), lstm_dir(xx, "lstm2/backward", True, "l2p", "l2r", "h2")], -1), "o2")
    xx = torch.relu(xx @ wq(T(W["dense1/kernel"]), "d1") + T(W["dense1/bias"]))
    xx = no._bn(xx, W, "bn_dense", f32)
    return torch.sigmoid(xx @ T(W["dense2/kernel"]) + T(W["dense2/bias"])).numpy()


def batched(fn, x, bs=32):
    return np.concatenate([fn(x[i:i + bs]) for i in range(0, len(x), bs)])


TAIL_W = {"l1p", "l1r", "l2p", "l2r", "d1"}
TAIL_A = {"feat", "h1", "o1", "h2", "o2"}
B34_W = {"b3s1", "b3s2", "b3res", "b4s1", "b4s2", "b4res", "final"}
B34_A = {"b3S1", "b3S2", "b3out", "b4S1", "b4S2", "b4out", "b2out"}
B2_W = {"b2s1", "b2s2", "b2res"}
B2_A = {"b2S1", "b2S2", "b1out"}
B1_W = {"b1s1", "b1s2", "b1res"}
PLANS = {
    "fp16 everywhere (today, uncalibrated)": (set(), set()),
    "all weights split": (set(GEMMS), set()),
    "all weights split, fp32 spec+conv0 in": (set(GEMMS), {"spec"}),
    "all acts split (weights fp16)": (set(), set(ACTS)),
    "w: tail": (TAIL_W, set()),
    "w: tail+b34": (TAIL_W | B34_W, set()),
    "w: tail+b34+b2": (TAIL_W | B34_W | B2_W, set()),
    "w: tail+b34, a: tail": (TAIL_W | B34_W, TAIL_A),
    "w: tail+b34, a: tail+b34": (TAIL_W | B34_W, TAIL_A | B34_A),
    "w: all, a: tail": (set(GEMMS), TAIL_A),
    "w: all, a: tail+b34": (set(GEMMS), TAIL_A | B34_A),
    "w: all, a: tail+b34+b2": (set(GEMMS), TAIL_A | B34_A | B2_A),
    "w: all, a: tail+b34+b2+spec": (set(GEMMS), TAIL_A | B34_A | B2_A | {"spec"}),
    "w: all but b1, a: tail+b34": (set(GEMMS) - B1_W - {"conv0"}, TAIL_A | B34_A),
    # candidates for the shipped "precise" mode: every weight split; only the listed activations stay single fp16
    "P1: fp16 S1 of all blocks": (set(GEMMS), set(ACTS) - {"b1S1", "b2S1", "b3S1", "b4S1"}),
    "P2: fp16 S1 of all blocks + fp16 recurrence": (set(GEMMS) - {"l1r", "l2r"}, set(ACTS) - {"b1S1", "b2S1", "b3S1", "b4S1", "h1", "h2"}),
    "P3: fp16 S1 of blocks 2-4": (set(GEMMS), set(ACTS) - {"b2S1", "b3S1", "b4S1"}),
    "P4: fp16 S1 of blocks 2-4 + fp16 recurrence": (set(GEMMS) - {"l1r", "l2r"}, set(ACTS) - {"b2S1", "b3S1", "b4S1", "h1", "h2"}),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--minutes", type=float, default=10.0)
    ap.add_argument("--bn-matched", action="store_true")
    ap.add_argument("--plans", default="")
    ap.add_argument("--seed", type=int, default=20251018)
    ap.add_argument("--single", action="store_true", help="one rounding point at a time")
    a = ap.parse_args()
    torch.set_num_threads(8)
    P, S = runtime.bundled_parameters()
    pcm = synth_pcm16(a.minutes * 60.0, seed=a.seed)
    spec, _, _ = so.make_spectrogram(pcm16_to_float(pcm), P["spectrogram"])
    x = po.cut_snippets(spec, 736)
    W = synthetic_weights(P, S, seed=1234)
    if a.bn_matched:
        W = ps.calibrate_bn(W, x[:4])
    t0 = time.time()
    ref = batched(lambda xb: no.forward(xb, W), x)
    print(f"# {len(x)} snippets, bn_matched={a.bn_matched}, reference forward {time.time() - t0:.0f} s", flush=True)
    if a.single:   # everything exact except ONE rounding point: that point's own contribution
        for g in GEMMS:
            PLANS[f"only fp16: weights of {g}"] = (set(GEMMS) - {g}, set(ACTS))
        for g in ACTS:
            PLANS[f"only fp16: activation {g}"] = (set(GEMMS), set(ACTS) - {g})
    names = [n for n in PLANS if (a.single and n.startswith("only")) or (not a.single and (not a.plans or any(s in n for s in a.plans.split(";"))))]
    for name in names:
        ws, as_ = PLANS[name]
        out = batched(lambda xb: forward_plan(xb, W, ws, as_), x)
        e = np.abs(out - ref)
        print(f"{name:44s} per-snippet max {e.max():.2e} mean {e.mean():.2e} p99.9 {np.quantile(e, 0.999):.2e}", flush=True)


if __name__ == "__main__":
    main()
