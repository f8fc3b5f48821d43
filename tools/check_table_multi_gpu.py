#!/usr/bin/env python
"""`orcai predict TABLE.csv -o DIR` on a box with several GPUs: the default shards the table over every visible GPU with one worker
process per GPU; the label files must be byte-identical to the one-GPU run.

    python tools/check_table_multi_gpu.py [--rows 12] [--seconds 60]
"""
from __future__ import annotations

import argparse
import os
import sys
import tempfile
import time
from pathlib import Path

import pandas as pd

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from orcai_b200 import predict as opredict, runtime  # noqa: E402
from orcai_b200.auxiliary import Messenger  # noqa: E402
from orcai_b200.synth import synth_pcm16  # noqa: E402
from orcai_b200.wavio import write_wav_pcm16  # noqa: E402
from orcai_b200.weights import save_npz, synthetic_weights  # noqa: E402


class Collect(Messenger):
    def __init__(self):
        super().__init__(verbosity=0)
        self.errors = []

    def error(self, text, *a, **k):
        self.errors.append(str(text))


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=12)
    ap.add_argument("--seconds", type=float, default=60.0)
    a = ap.parse_args()
    import torch

    n_gpu = torch.cuda.device_count()
    d = Path(tempfile.mkdtemp(prefix="orcai_multi_"))
    P, S = runtime.bundled_parameters()
    model_dir = d / "orcai-V1"
    model_dir.mkdir()
    from importlib.resources import files as pkg_files

    src = pkg_files("orcai_b200.models").joinpath("orcai-V1")
    for name in ("orcai_parameter.json", "model_shape.json"):
        (model_dir / name).write_text(src.joinpath(name).read_text())
    save_npz(synthetic_weights(P, S, seed=1234), model_dir / "orcai-v1.weights.npz")
    for k in range(4):
        write_wav_pcm16(d / f"rec{k}.wav", synth_pcm16(a.seconds * (1 + 0.25 * k), seed=20251018 + k, calls_per_minute=30.0))
    names = [f"r{j:02d}" for j in range(a.rows)] + ["gone"]
    pd.DataFrame({"recording": names, "channel": 1, "base_dir_recording": str(d),
                  "rel_recording_path": [f"rec{j % 4}.wav" for j in range(a.rows)] + ["missing.wav"]}).to_csv(d / "t.csv", index=False)
    out = {}
    for tag, env, procs in (("all_gpus_worker_processes", None, "1"), ("all_gpus_default", None, None), ("one_gpu", "0", None)):
        if env is None:
            os.environ.pop("ORCAI_B200_DEVICES", None)
        else:
            os.environ["ORCAI_B200_DEVICES"] = env
        if procs is None:
            os.environ.pop("ORCAI_B200_TABLE_PROCESSES", None)
        else:
            os.environ["ORCAI_B200_TABLE_PROCESSES"] = procs
        o = d / tag
        o.mkdir()
        m = Collect()
        t0 = time.perf_counter()
        opredict.predict(d / "t.csv", model_dir=model_dir, output_path=str(o), verbosity=0, msgr=m)
        dt = time.perf_counter() - t0
        out[tag] = {f.name: f.read_bytes() for f in o.iterdir()}
        print(f"{tag}: {len(out[tag])} label files in {dt:.2f} s, errors reported: {m.errors}", flush=True)
        assert any("gone" in e for e in m.errors), "the missing file must be reported against its row"
    same = out["all_gpus_default"] == out["one_gpu"] == out["all_gpus_worker_processes"] and len(out["one_gpu"]) == a.rows
    print(f"{n_gpu} GPUs visible; default run == one-GPU run byte for byte: {same}")
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
