#!/usr/bin/env python
"""`orcai predict TABLE.csv -o OUTDIR` through the public Python entry point, WAV files on disk -> label files on disk
(BASELINE configs[3] in miniature: a recording table whose rows reference K distinct seeded 1-hour files).

    python tools/bench_table.py [--rows 24] [--files 4] [--hours 1.0] [--dir /tmp/orcai_table]

Prints one JSON line: wall seconds, hours of audio per second, and the per-stage shares the run reports.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import pandas as pd

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from orcai_b200 import predict as opredict, runtime  # noqa: E402
from orcai_b200.synth import synth_pcm16  # noqa: E402
from orcai_b200.wavio import write_wav_pcm16  # noqa: E402


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=24)
    ap.add_argument("--files", type=int, default=4)
    ap.add_argument("--hours", type=float, default=1.0)
    ap.add_argument("--dir", default="/tmp/orcai_table")
    a = ap.parse_args()
    root = Path(a.dir)
    (root / "wav").mkdir(parents=True, exist_ok=True)
    (root / "out").mkdir(parents=True, exist_ok=True)
    for k in range(a.files):
        p = root / "wav" / f"rec{k}.wav"
        if not p.exists():
            write_wav_pcm16(p, synth_pcm16(a.hours * 3600.0, seed=20251018 + k), 48000)
    P, S = runtime.bundled_parameters()
    model_dir = root / "orcai-V1"
    model_dir.mkdir(exist_ok=True)
    (model_dir / "orcai_parameter.json").write_text(json.dumps(P))
    (model_dir / "model_shape.json").write_text(json.dumps(S))
    os.environ.setdefault("ORCAI_B200_SYNTHETIC_WEIGHTS", "1234")
    table = pd.DataFrame({
        "recording": [f"row{r:03d}" for r in range(a.rows)],
        "base_dir_recording": str(root / "wav"),
        "rel_recording_path": [f"rec{r % a.files}.wav" for r in range(a.rows)],
        "channel": 1,
    })
    table.to_csv(root / "table.csv", index=False)
    # warm-up: page cache, library load (a short table)
    table.iloc[: max(a.files, 8)].to_csv(root / "warm.csv", index=False)
    opredict.predict(root / "warm.csv", model_dir=model_dir, output_path=str(root / "out"), overwrite=True, verbosity=0)
    t0 = time.perf_counter()
    opredict.predict(root / "table.csv", model_dir=model_dir, output_path=str(root / "out"), overwrite=True, verbosity=0)
    dt = time.perf_counter() - t0
    n_out = len(list((root / "out").glob("*_predicted.txt")))
    print(json.dumps({"workload": f"orcai predict table: {a.rows} rows x {a.hours:g} h ({a.files} distinct WAV files on disk), label files written",
                      "seconds": dt, "h_audio_per_s": a.rows * a.hours / dt, "ms_per_recording": 1e3 * dt / a.rows, "label_files": n_out,
                      "devices": os.environ.get("ORCAI_B200_DEVICES", "(one)"), "host_threads": len(os.sched_getaffinity(0))}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
