#!/usr/bin/env python
"""BASELINE config 5: orcai-V1 snippet forward, batch-size sweep 1-4096 on a resident recording.

Reports per network path (0 fp32 CUDA cores, 3 fp16 tcgen05 fused, 2 bf16 tcgen05 layer-wise) the snippets/s, the achieved
algorithmic TFLOP/s (0.972 GFLOP per snippet) and the max probability deviation from the fp32 path.

    python tools/sweep_batch.py > profiles/r01_batch_sweep.json      (needs a B200)
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from orcai_b200 import runtime  # noqa: E402
from orcai_b200.synth import synth_pcm16  # noqa: E402
from orcai_b200.weights import synthetic_weights  # noqa: E402

FLOP = 0.972e9


def main() -> int:
    P, S = runtime.bundled_parameters()
    ctx = runtime.get_context(P, S, 0)
    ctx.load_weights(synthetic_weights(P, S, seed=1234))
    pcm = synth_pcm16(2.25 * 3600.0, seed=20251018)          # 4 126 snippets
    ctx.upload_pcm(pcm)
    ctx.spectrogram_resident(False)
    ctx.set_option("chunk", 4096)
    sizes = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096]
    ref = {}
    rows = []
    for path, name in ((0, "fp32 CUDA cores"), (3, "fp16 tcgen05 fused"), (2, "bf16 tcgen05 layer-wise")):
        ctx.set_option("net_path", path)
        for n in sizes:
            if path != 3 and n > 512:
                continue   # the slow paths only serve as tolerance references
            out = ctx.forward_resident(0, n)                 # warm-up + result
            best = 1e30
            for _ in range(3):
                ctx.forward_resident(0, n)
                best = min(best, ctx.timings()["network_ms"])
            if path == 0:
                ref[n] = out
            err = float(np.abs(out - ref[n]).max()) if n in ref else None
            rows.append({"path": name, "batch": n, "ms": best, "snippets_per_s": n / (best * 1e-3), "tflops": n * FLOP / (best * 1e-3) / 1e12,
                         "max_abs_dev_vs_fp32": err})
            print(json.dumps(rows[-1]), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
