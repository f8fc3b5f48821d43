#!/usr/bin/env python
"""Stage-by-stage bring-up check of liborcai_b200 against the CPU oracle (run on a B200 via gpurun).

    python tools/gpu_check.py [--seconds 60] [--snippets 4]

Prints one line per stage with the error against the oracle; exits non-zero on a failed gate.
"""

from __future__ import annotations

import argparse
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from oracle import network_oracle, postprocess_oracle as po, spectrogram_oracle as so  # noqa: E402
from orcai_b200 import runtime  # noqa: E402
from orcai_b200.synth import pcm16_to_float, synth_pcm16  # noqa: E402
from orcai_b200.weights import synthetic_weights  # noqa: E402


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--snippets", type=int, default=4)
    ap.add_argument("--skip-net", action="store_true")
    args = ap.parse_args()
    ok = True
    P, S = runtime.bundled_parameters()
    sp = P["spectrogram"]
    ctx = runtime.get_context(P, S, 0)

    pcm = synth_pcm16(args.seconds, seed=20251018)
    y = pcm16_to_float(pcm)
    t0 = time.time()
    db_ref, freqs, times = so.calculate_spectrogram(y, sp)
    spec_ref, lo_ref, hi_ref = so.preprocess_spectrogram(db_ref, freqs, sp)
    print(f"[oracle] spectrogram {time.time() - t0:.2f}s  T={spec_ref.shape[0]} lo={lo_ref:.6f} hi={hi_ref:.6f}")

    for name, arr, f64 in (("int16/f64", pcm, 1), ("float32/f64", y, 1), ("int16/f32", pcm, 0)):
        ctx.set_option("stft_f64", f64)
        spec, st = ctx.spectrogram(arr)
        spec, st = ctx.spectrogram(arr)  # second call: warm timings
        db = ctx.read_db(0, spec.shape[0])
        band = db_ref[:171].T
        err_db = np.abs(db - band)
        err_sp = np.abs(spec - spec_ref)
        pmax_ref = float(np.max(np.abs(so.stft_complex64(y)) ** 2))
        print(
            f"[K1/{name}] T={st.n_frames} ref_power={st.ref_power:.6e} (oracle {pmax_ref:.6e}) db_ref={st.db_ref:.5f} "
            f"dB max err {err_db.max():.3e} mean {err_db.mean():.3e} p99.99 {np.quantile(err_db, 0.9999):.3e} | "
            f"lo {st.lo:.6f} hi {st.hi:.6f} (oracle {lo_ref:.6f} {hi_ref:.6f}) | norm max err {err_sp.max():.3e}"
        )
        # the select must return the exact order statistics of the array the GPU itself produced
        flat = np.sort(db.ravel())
        exact = flat[st.rank_lo] == st.lo and flat[st.rank_hi] == st.hi
        print(f"[select/{name}] exact order statistics of the device array: {exact} (ranks {st.rank_lo}, {st.rank_hi})")
        ok &= bool(exact) and err_db.max() < (1e-3 if f64 else 5e-3) and err_sp.max() < 1e-4
        print(f"           timings {ctx.timings()}")
    ctx.set_option("stft_f64", 1)

    # post-processing on random predictions (bit-exact integers)
    rng = np.random.default_rng(3)
    T = spec_ref.shape[0]
    N = (T - 736) // 368 + 1
    preds = rng.random((N, 46, 7), dtype=np.float32)
    preds = (0.6 * preds + 0.4 * np.repeat(rng.random((N, 1, 7), dtype=np.float32), 46, axis=1)).astype(np.float32) * 0.6
    agg_r, cnt_r = po.aggregate_predictions(preds, T, 736, 4, 7)
    s_r, e_r, n_r = po.binary_predictions(agg_r, cnt_r, P["calls"])
    agg, cnt, lab, sta, sto = ctx.postprocess(preds, T)
    same = (
        np.array_equal(agg, agg_r)
        and np.array_equal(cnt, cnt_r)
        and list(sta) == [int(v) for v in s_r]
        and list(sto) == [int(v) for v in e_r]
        and [P["calls"][i] for i in lab] == n_r
    )
    print(f"[K7] N={N} segments={len(sta)} (oracle {len(s_r)}) bit-exact={same}")
    ok &= same
    lab2, sta2, sto2 = ctx.threshold_segments(agg_r, cnt_r)
    same2 = list(sta2) == [int(v) for v in s_r] and list(sto2) == [int(v) for v in e_r]
    print(f"[K7/threshold_segments] bit-exact={same2}")
    ok &= same2

    if not args.skip_net:
        W = synthetic_weights(P, S, seed=1234)
        ctx.load_weights(W)
        snips = po.cut_snippets(spec_ref, 736)[: args.snippets]
        t0 = time.time()
        ref, inter = network_oracle.forward(snips, W, return_intermediates=True)
        print(f"[oracle] network {time.time() - t0:.2f}s for {len(snips)} snippets; prob range [{ref.min():.4f}, {ref.max():.4f}] mean {ref.mean():.4f}")
        out = ctx.forward_host(snips)
        err = np.abs(out - ref)
        print(f"[net fp32/host] max abs err {err.max():.3e} mean {err.mean():.3e}")
        ok &= err.max() < 1e-3
        # resident path (normalise-on-load from the raw dB buffer)
        ctx.upload_pcm(pcm)
        ctx.spectrogram_resident(True)
        out2 = ctx.forward_resident(0, len(snips))
        err2 = np.abs(out2 - ref)
        print(f"[net fp32/resident] max abs err {err2.max():.3e}")
        ok &= err2.max() < 1e-3
        # fused predict
        st, agg, cnt, lab, sta, sto = ctx.predict_pcm(pcm)
        print(f"[predict_pcm] segments={len(sta)} agg range [{agg.min():.4f}, {agg.max():.4f}] timings {ctx.timings()}")
        allp = ctx.forward_resident(0, N)
        a2, c2, l2, s2, e2 = ctx.postprocess(allp, T)
        same3 = np.array_equal(a2, agg) and list(s2) == list(sta) and list(e2) == list(sto)
        print(f"[predict_pcm vs staged] identical={same3}")
        ok &= same3
    print("ALL OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
