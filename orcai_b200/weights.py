"""orcai-V1 weight container: naming scheme, shapes, seeded synthetic weights, npz I/O.

The names mirror the Keras variables of the graph built by the reference's
``res_net_LSTM_arch`` (``src/orcAI/architectures.py:120-241``), in Keras layouts:

    conv0/kernel (3,3,1,16)  conv0/bias (16)          bn0/{gamma,beta,moving_mean,moving_variance}
    block{b}/sep1/{depthwise_kernel (3,3,Ci,1), pointwise_kernel (1,1,Ci,Co), bias (Co)}   block{b}/bn1/...
    block{b}/sep2/{depthwise_kernel (3,3,Co,1), pointwise_kernel (1,1,Co,Co), bias (Co)}   block{b}/bn2/...
    block{b}/res/{kernel (1,1,Ci,Co), bias (Co)}
    final/sep/{depthwise_kernel (3,3,60,1), pointwise_kernel (1,1,60,36), bias (36)}       final/bn/...
    lstm{1,2}/{forward,backward}/{kernel (I,4U), recurrent_kernel (U,4U), bias (4U)}       gate order i,f,c,o
    dense1/{kernel (2U,128), bias (128)}   bn_dense/...   dense2/{kernel (128,L), bias (L)}

The packaged ``orcai-v1.keras`` blob is absent from the reference mount
(``.MISSING_LARGE_BLOBS``), so benchmarks and parity tests run on seeded synthetic
weights drawn from the reference's initialisers (he_normal convolutions,
glorot_uniform / orthogonal LSTM kernels, ``architectures.py:127-128,214``).
"""

from __future__ import annotations

from pathlib import Path

import numpy as np

BN_KEYS = ("gamma", "beta", "moving_mean", "moving_variance")
FINAL_SEP_FILTERS = 36  # architectures.py:198-199
ENTRY_FILTERS = 16  # architectures.py:164
DENSE_UNITS = 128  # architectures.py:231


def expected_shapes(orcai_parameter: dict, shape: dict) -> dict[str, tuple[int, ...]]:
    """name -> shape for every variable of the ResNetLSTM graph described by the parameter dicts."""
    m = orcai_parameter["model"]
    filters = list(m["filters"])
    k = int(m["kernel_size"])
    U = int(m["lstm_units"])
    L = int(shape["num_labels"])
    Wf = int(shape["input_shape"][1])
    s: dict[str, tuple[int, ...]] = {}

    def bn(prefix, c):
        for key in BN_KEYS:
            s[f"{prefix}/{key}"] = (c,)

    def sep(prefix, ci, co):
        s[f"{prefix}/depthwise_kernel"] = (k, k, ci, 1)
        s[f"{prefix}/pointwise_kernel"] = (1, 1, ci, co)
        s[f"{prefix}/bias"] = (co,)

    s["conv0/kernel"] = (k, k, 1, ENTRY_FILTERS)
    s["conv0/bias"] = (ENTRY_FILTERS,)
    bn("bn0", ENTRY_FILTERS)
    ci = ENTRY_FILTERS
    w = Wf
    for b, co in enumerate(filters, start=1):
        sep(f"block{b}/sep1", ci, co)
        bn(f"block{b}/bn1", co)
        sep(f"block{b}/sep2", co, co)
        bn(f"block{b}/bn2", co)
        s[f"block{b}/res/kernel"] = (1, 1, ci, co)
        s[f"block{b}/res/bias"] = (co,)
        ci = co
        w = -(-w // 2)
    sep("final/sep", ci, FINAL_SEP_FILTERS)
    bn("final/bn", FINAL_SEP_FILTERS)
    feat = w * FINAL_SEP_FILTERS
    for layer, fin in (("lstm1", feat), ("lstm2", 2 * U)):
        for d in ("forward", "backward"):
            s[f"{layer}/{d}/kernel"] = (fin, 4 * U)
            s[f"{layer}/{d}/recurrent_kernel"] = (U, 4 * U)
            s[f"{layer}/{d}/bias"] = (4 * U,)
    s["dense1/kernel"] = (2 * U, DENSE_UNITS)
    s["dense1/bias"] = (DENSE_UNITS,)
    bn("bn_dense", DENSE_UNITS)
    s["dense2/kernel"] = (DENSE_UNITS, L)
    s["dense2/bias"] = (L,)
    return s


def _he_normal(rng, shape, fan_in):
    std = np.sqrt(2.0 / fan_in) / 0.87962566103423978  # keras truncated-normal correction
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2.0
    return (x * std).astype(np.float32)


def _glorot_uniform(rng, shape):
    lim = np.sqrt(6.0 / (shape[0] + shape[1]))
    return rng.uniform(-lim, lim, shape).astype(np.float32)


def _orthogonal(rng, shape):
    rows, cols = shape
    a = rng.standard_normal((max(rows, cols), min(rows, cols)))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))
    if rows < cols:
        q = q.T
    return q[:rows, :cols].astype(np.float32)


def synthetic_weights(orcai_parameter: dict, shape: dict, seed: int = 1234) -> dict[str, np.ndarray]:
    """Deterministic random weights with the reference's initialiser families."""
    rng = np.random.default_rng(seed)
    shapes = expected_shapes(orcai_parameter, shape)
    U = int(orcai_parameter["model"]["lstm_units"])
    W: dict[str, np.ndarray] = {}
    for name, shp in shapes.items():
        leaf = name.rsplit("/", 1)[1]
        if leaf == "gamma":
            W[name] = rng.uniform(0.5, 1.5, shp).astype(np.float32)
        elif leaf == "beta":
            W[name] = (0.1 * rng.standard_normal(shp)).astype(np.float32)
        elif leaf == "moving_mean":
            W[name] = (0.1 * rng.standard_normal(shp)).astype(np.float32)
        elif leaf == "moving_variance":
            W[name] = rng.uniform(0.5, 1.5, shp).astype(np.float32)
        elif leaf == "depthwise_kernel":
            W[name] = _he_normal(rng, shp, shp[0] * shp[1])
        elif leaf == "pointwise_kernel":
            W[name] = _he_normal(rng, shp, shp[2])
        elif leaf == "recurrent_kernel":
            W[name] = _orthogonal(rng, shp)
        elif leaf == "kernel" and "lstm" in name:
            W[name] = _glorot_uniform(rng, shp)
        elif leaf == "kernel" and name.startswith("dense2"):
            W[name] = _glorot_uniform(rng, shp)
        elif leaf == "kernel" and len(shp) == 4:
            W[name] = _he_normal(rng, shp, shp[0] * shp[1] * shp[2])
        elif leaf == "kernel":
            W[name] = _he_normal(rng, shp, shp[0])
        elif leaf == "bias" and "lstm" in name:
            b = np.zeros(shp, np.float32)
            b[U : 2 * U] = 1.0  # unit_forget_bias
            W[name] = b
        elif leaf == "bias" and name.startswith("dense2"):
            # spread the label logits so that averaged probabilities straddle the 0.25 decision level
            W[name] = np.linspace(-1.6, -0.6, shp[0]).astype(np.float32)
        elif leaf == "bias":
            W[name] = (0.05 * rng.standard_normal(shp)).astype(np.float32)
        else:  # pragma: no cover
            raise KeyError(name)
    return W


def check_weights(W: dict, orcai_parameter: dict, shape: dict) -> None:
    """Raise ValueError when a variable is missing or has the wrong shape."""
    for name, shp in expected_shapes(orcai_parameter, shape).items():
        if name not in W:
            raise ValueError(f"missing weight '{name}'")
        if tuple(np.shape(W[name])) != shp:
            raise ValueError(f"weight '{name}' has shape {tuple(np.shape(W[name]))}, expected {shp}")


def save_npz(W: dict, path: Path | str) -> None:
    np.savez(path, **{k.replace("/", "__"): np.asarray(v, np.float32) for k, v in W.items()})


def load_npz(path: Path | str) -> dict[str, np.ndarray]:
    with np.load(path) as z:
        return {k.replace("__", "/"): np.asarray(z[k], np.float32) for k in z.files}
