#!/usr/bin/env python
"""Check that a reference model directory (orcai_parameter.json, model_shape.json, <name>.keras or model_weights.h5) loads
through orcai_b200's own readers, and print what was found.  Needs neither keras, h5py nor a GPU.

    python tools/verify_keras.py path/to/orcai-V1 [--export]

--export additionally writes <name>.weights.npz (orcai_b200's container) next to the Keras file.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from orcai_b200.keras_weights import load_keras_archive, load_weights_h5  # noqa: E402
from orcai_b200.weights import expected_shapes, save_npz  # noqa: E402


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("model_dir")
    ap.add_argument("--export", action="store_true")
    a = ap.parse_args()
    d = Path(a.model_dir)
    P = json.loads((d / "orcai_parameter.json").read_text())
    S = json.loads((d / "model_shape.json").read_text())
    keras_file, legacy = d / (P["name"] + ".keras"), d / "model_weights.h5"
    if keras_file.exists():
        src, W = keras_file, load_keras_archive(keras_file, P, S)
    elif legacy.exists():
        src, W = legacy, load_weights_h5(legacy, P, S)
    else:
        print(f"neither {keras_file.name} nor {legacy.name} in {d}")
        return 1
    print(f"{src}  sha256 {hashlib.sha256(src.read_bytes()).hexdigest()[:16]}...  {src.stat().st_size} bytes")
    total = 0
    for name, shape in expected_shapes(P, S).items():
        w = W[name]
        total += w.size
        print(f"  {name:34s} {str(tuple(w.shape)):20s} mean {float(w.mean()):+.4e}  std {float(w.std()):.4e}  finite {bool(np.isfinite(w).all())}")
    print(f"{len(W)} tensors, {total} parameters")
    if a.export:
        out = d / (P["name"] + ".weights.npz")
        save_npz(W, out)
        print("wrote", out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
