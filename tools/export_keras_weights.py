#!/usr/bin/env python
"""Export the variables of a trained orcAI ResNetLSTM Keras model to orcai_b200's npz container.

Run on a machine that HAS keras/tensorflow (this repo's image does not):

    python tools/export_keras_weights.py path/to/model_dir

Reads `<model_dir>/<name>.keras` (or the legacy `model_weights.h5` through the reference's
`load_orcai_model`) and writes `<model_dir>/<name>.weights.npz` with the names of
`orcai_b200/weights.py`.  Layers are matched by TYPE ORDER of the functional graph built by
`res_net_LSTM_arch` (reference `src/orcAI/architectures.py:120-241`), not by auto-generated layer
names.  UNTESTED in this repository (no keras here) - verify with
`orcai_b200.weights.check_weights` which the loader runs on every import.
"""

from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np


def export(model, n_blocks: int) -> dict[str, np.ndarray]:
    import keras

    conv, sep, bn, bi, dense = [], [], [], [], []
    for layer in model.layers:
        if isinstance(layer, keras.layers.SeparableConv2D):
            sep.append(layer)
        elif isinstance(layer, keras.layers.Conv2D):
            conv.append(layer)
        elif isinstance(layer, keras.layers.BatchNormalization):
            bn.append(layer)
        elif isinstance(layer, keras.layers.Bidirectional):
            bi.append(layer)
        elif isinstance(layer, keras.layers.Dense):
            dense.append(layer)
    assert len(conv) == 1 + n_blocks and len(sep) == 2 * n_blocks + 1 and len(bn) == 2 * n_blocks + 3, "unexpected graph"
    W: dict[str, np.ndarray] = {}

    def put_bn(prefix, layer):
        g, b, m, v = layer.get_weights()
        W[f"{prefix}/gamma"], W[f"{prefix}/beta"], W[f"{prefix}/moving_mean"], W[f"{prefix}/moving_variance"] = g, b, m, v

    def put_sep(prefix, layer):
        d, p, b = layer.get_weights()
        W[f"{prefix}/depthwise_kernel"], W[f"{prefix}/pointwise_kernel"], W[f"{prefix}/bias"] = d, p, b

    W["conv0/kernel"], W["conv0/bias"] = conv[0].get_weights()
    put_bn("bn0", bn[0])
    for b in range(n_blocks):
        put_sep(f"block{b + 1}/sep1", sep[2 * b])
        put_bn(f"block{b + 1}/bn1", bn[1 + 2 * b])
        put_sep(f"block{b + 1}/sep2", sep[2 * b + 1])
        put_bn(f"block{b + 1}/bn2", bn[2 + 2 * b])
        W[f"block{b + 1}/res/kernel"], W[f"block{b + 1}/res/bias"] = conv[1 + b].get_weights()
    put_sep("final/sep", sep[2 * n_blocks])
    put_bn("final/bn", bn[1 + 2 * n_blocks])
    for i, layer in enumerate(bi, start=1):
        fk, fr, fb = layer.forward_layer.get_weights()
        bk, br, bb = layer.backward_layer.get_weights()
        W[f"lstm{i}/forward/kernel"], W[f"lstm{i}/forward/recurrent_kernel"], W[f"lstm{i}/forward/bias"] = fk, fr, fb
        W[f"lstm{i}/backward/kernel"], W[f"lstm{i}/backward/recurrent_kernel"], W[f"lstm{i}/backward/bias"] = bk, br, bb
    W["dense1/kernel"], W["dense1/bias"] = dense[0].get_weights()
    put_bn("bn_dense", bn[2 + 2 * n_blocks])
    W["dense2/kernel"], W["dense2/bias"] = dense[1].get_weights()
    return {k: np.asarray(v, np.float32) for k, v in W.items()}


def main():
    model_dir = Path(sys.argv[1])
    import keras

    P = json.loads((model_dir / "orcai_parameter.json").read_text())
    S = json.loads((model_dir / "model_shape.json").read_text())
    model = keras.saving.load_model(model_dir / (P["name"] + ".keras"), compile=False)
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from orcai_b200.weights import check_weights, save_npz

    W = export(model, len(P["model"]["filters"]))
    check_weights(W, P, S)
    save_npz(W, model_dir / (P["name"] + ".weights.npz"))
    print("wrote", model_dir / (P["name"] + ".weights.npz"))


if __name__ == "__main__":
    main()
