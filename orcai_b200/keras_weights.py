"""Read the reference's Keras model artefacts without keras / h5py (reference ``io.py:386-404``).

Two containers are understood, both through ``orcai_b200.hdf5_min``:

* ``<name>.keras`` — the Keras 3 zip archive (``config.json``, ``metadata.json``, ``model.weights.h5``); the weight file
  keeps every layer under ``layers/<layer>/vars/<i>`` (``Bidirectional``: ``forward_layer/cell/vars/<i>`` and
  ``backward_layer/cell/vars/<i>``), plus ``optimizer/vars`` which is ignored;
* ``model_weights.h5`` — the legacy Keras 2 layout ``<layer>/<layer>/kernel:0`` that ``model.load_weights`` accepts for the
  reference's legacy models.

Layers are identified by the STRUCTURE of their variables (count, rank, names where the legacy layout has them), never by
their auto-generated names, and ordered by ``config.json``'s layer list when the archive has one, else by the numeric
suffix of the group names (``conv2d``, ``conv2d_1``, ...), which follows creation order.  The k-th layer of each kind is
then the k-th layer of that kind in ``res_net_LSTM_arch`` (reference ``architectures.py:162-241``).  The packaged
``orcai-v1.keras`` blob is absent from the mount this was developed against: the path is tested on archives written by
``write_keras_archive`` below and the HDF5 reader on a libhdf5-written file; ``check_weights`` guards every load.
"""

from __future__ import annotations

import io as _io
import json
import re
import zipfile
from pathlib import Path

import numpy as np

from orcai_b200.hdf5_min import H5File, Hdf5Error, write_h5
from orcai_b200.weights import check_weights, expected_shapes

_SUFFIX = re.compile(r"^(.*?)(?:_(\d+))?$")


def _suffix_key(name: str):
    m = _SUFFIX.match(name)
    return (m.group(1), int(m.group(2)) if m.group(2) else 0)


def _collect(h5: H5File) -> dict[str, dict[str, np.ndarray]]:
    """{layer group name: {variable key: array}}; keys are '0','1',.. or 'kernel',.., prefixed 'forward/' / 'backward/'."""
    layers: dict[str, dict[str, np.ndarray]] = {}
    for path, ds in h5.datasets().items():
        parts = [p for p in path.split("/") if p]
        if not parts or parts[0] in ("optimizer", "optimizer_weights", "vars"):
            continue
        if parts[0] == "layers":      # Keras 3 layout
            if len(parts) < 4 or "vars" not in parts:
                continue
            lname = parts[1]
            mid = parts[2:-1]
        else:                          # legacy layout: <layer>/<layer>/.../<variable>:0
            if len(parts) < 2:
                continue
            lname = parts[0]
            mid = parts[1:-1]
        leaf = parts[-1].split(":")[0]
        direction = ""
        for m in mid:
            if m.startswith("forward"):
                direction = "forward/"
            elif m.startswith("backward"):
                direction = "backward/"
        layers.setdefault(lname, {})[direction + leaf] = ds.read().astype(np.float32)
    return layers


_LEGACY_ORDER = {
    "conv": ("kernel", "bias"),
    "dense": ("kernel", "bias"),
    "sep": ("depthwise_kernel", "pointwise_kernel", "bias"),
    "bn": ("gamma", "beta", "moving_mean", "moving_variance"),
    "lstm": ("kernel", "recurrent_kernel", "bias"),
}


def _kind(vars_: dict[str, np.ndarray]) -> str | None:
    keys = set(vars_)
    if any(k.startswith("forward/") for k in keys):
        return "bi"
    if keys >= {"depthwise_kernel"} or (keys == {"0", "1", "2"} and vars_["0"].ndim == 4 and vars_["1"].ndim == 4):
        return "sep"
    if keys >= {"gamma"} or (keys == {"0", "1", "2", "3"} and all(vars_[k].ndim == 1 for k in keys)):
        return "bn"
    first = vars_.get("kernel", vars_.get("0"))
    if first is not None and len(keys) == 2:
        return "conv" if first.ndim == 4 else "dense" if first.ndim == 2 else None
    return None


def _ordered(vars_: dict[str, np.ndarray], kind: str, prefix: str = "") -> list[np.ndarray]:
    names = _LEGACY_ORDER[kind]
    if prefix + names[0] in vars_:
        return [vars_[prefix + n] for n in names]
    return [vars_[prefix + str(i)] for i in range(len(names))]


def weights_from_h5(h5: H5File, orcai_parameter: dict, shape: dict, config: dict | None = None) -> dict[str, np.ndarray]:
    """Map the variables of a Keras weight file onto orcai_b200's names (``orcai_b200/weights.py``)."""
    layers = _collect(h5)
    if not layers:
        raise Hdf5Error("no Keras layer variables found in the HDF5 file")
    order = sorted(layers, key=_suffix_key)
    cfg_names: list = []
    if config is not None:
        try:
            cfg_names = [l["name"] if "name" in l else l["config"]["name"] for l in config["config"]["layers"]]
        except (KeyError, TypeError):
            cfg_names = []
    ordered_by_config = bool(cfg_names) and all(n in cfg_names for n in layers)
    if ordered_by_config:
        order = sorted(layers, key=cfg_names.index)
    by_kind: dict[str, list[str]] = {"conv": [], "sep": [], "bn": [], "bi": [], "dense": []}
    for name in order:
        k = _kind(layers[name])
        if k is None:
            raise Hdf5Error(f"layer group {name!r}: unrecognised variable structure {sorted(layers[name])}")
        by_kind[k].append(name)
    if not ordered_by_config:
        # Without a layer list the only order information is the numeric suffix of auto-generated names (conv2d, conv2d_1, ..: creation
        # order; this reader does not parse the legacy file's `layer_names` attribute).  Several layers of one kind have identical
        # shapes (the two BatchNormalizations of a block), so a renamed or re-created layer could be swapped silently: refuse any
        # family that is not one base name with consecutive suffixes.
        for kind, names in by_kind.items():
            keys = [_suffix_key(n) for n in names]
            bases = {b for b, _ in keys}
            nums = [k for _, k in keys]
            if nums and (len(bases) > 1 or nums != list(range(nums[0], nums[0] + len(nums)))):
                raise Hdf5Error(
                    f"cannot establish the order of the {kind!r} layers from their names {names}: expected one auto-generated family with "
                    "consecutive suffixes (e.g. conv2d, conv2d_1, ..). Export the weights with tools/export_keras_weights.py (-> .weights.npz) "
                    "or provide the .keras archive, whose config.json lists the layers in order.")
    nb = len(orcai_parameter["model"]["filters"])
    want = {"conv": 1 + nb, "sep": 2 * nb + 1, "bn": 2 * nb + 3, "bi": 2, "dense": 2}
    got = {k: len(v) for k, v in by_kind.items()}
    if got != want:
        raise Hdf5Error(f"weight file does not hold a ResNetLSTM of {nb} blocks: layer counts {got}, expected {want}")
    W: dict[str, np.ndarray] = {}

    def put(prefix: str, names: tuple[str, ...], arrays: list[np.ndarray]):
        for n, a in zip(names, arrays):
            W[f"{prefix}/{n}"] = a

    conv, sep, bn, bi, dense = (by_kind[k] for k in ("conv", "sep", "bn", "bi", "dense"))
    put("conv0", _LEGACY_ORDER["conv"], _ordered(layers[conv[0]], "conv"))
    put("bn0", _LEGACY_ORDER["bn"], _ordered(layers[bn[0]], "bn"))
    for b in range(nb):
        put(f"block{b + 1}/sep1", _LEGACY_ORDER["sep"], _ordered(layers[sep[2 * b]], "sep"))
        put(f"block{b + 1}/bn1", _LEGACY_ORDER["bn"], _ordered(layers[bn[1 + 2 * b]], "bn"))
        put(f"block{b + 1}/sep2", _LEGACY_ORDER["sep"], _ordered(layers[sep[2 * b + 1]], "sep"))
        put(f"block{b + 1}/bn2", _LEGACY_ORDER["bn"], _ordered(layers[bn[2 + 2 * b]], "bn"))
        put(f"block{b + 1}/res", _LEGACY_ORDER["conv"], _ordered(layers[conv[1 + b]], "conv"))
    put("final/sep", _LEGACY_ORDER["sep"], _ordered(layers[sep[2 * nb]], "sep"))
    put("final/bn", _LEGACY_ORDER["bn"], _ordered(layers[bn[1 + 2 * nb]], "bn"))
    for i, name in enumerate(bi, start=1):
        for d in ("forward", "backward"):
            put(f"lstm{i}/{d}", _LEGACY_ORDER["lstm"], _ordered(layers[name], "lstm", d + "/"))
    put("dense1", _LEGACY_ORDER["dense"], _ordered(layers[dense[0]], "dense"))
    put("bn_dense", _LEGACY_ORDER["bn"], _ordered(layers[bn[2 + 2 * nb]], "bn"))
    put("dense2", _LEGACY_ORDER["dense"], _ordered(layers[dense[1]], "dense"))
    W = {k: np.ascontiguousarray(v, np.float32) for k, v in W.items()}
    check_weights(W, orcai_parameter, shape)
    return W


def load_keras_archive(path: Path | str, orcai_parameter: dict, shape: dict) -> dict[str, np.ndarray]:
    """``<name>.keras`` (zip) -> weight dict."""
    with zipfile.ZipFile(path) as z:
        names = z.namelist()
        wname = next((n for n in names if n.endswith("model.weights.h5")), None)
        if wname is None:
            raise Hdf5Error(f"{path}: no model.weights.h5 inside the archive (members: {names})")
        config = json.loads(z.read("config.json")) if "config.json" in names else None
        return weights_from_h5(H5File(z.read(wname)), orcai_parameter, shape, config)


def load_weights_h5(path: Path | str, orcai_parameter: dict, shape: dict) -> dict[str, np.ndarray]:
    """``model_weights.h5`` (legacy layout) or a bare Keras 3 ``*.weights.h5`` -> weight dict."""
    return weights_from_h5(H5File(path), orcai_parameter, shape)


# -------------------------------------------------------------------------------------------------
# writers — produce the two containers from a weight dict (tests; shipping synthetic model directories)
# -------------------------------------------------------------------------------------------------
def _layer_plan(orcai_parameter: dict):
    """[(class_name, our prefix)] in the creation order of res_net_LSTM_arch (weighted layers only)."""
    nb = len(orcai_parameter["model"]["filters"])
    plan = [("Conv2D", "conv0"), ("BatchNormalization", "bn0")]
    for b in range(1, nb + 1):
        plan += [("SeparableConv2D", f"block{b}/sep1"), ("BatchNormalization", f"block{b}/bn1"),
                 ("SeparableConv2D", f"block{b}/sep2"), ("BatchNormalization", f"block{b}/bn2"), ("Conv2D", f"block{b}/res")]
    plan += [("SeparableConv2D", "final/sep"), ("BatchNormalization", "final/bn"), ("Bidirectional", "lstm1"),
             ("Bidirectional", "lstm2"), ("Dense", "dense1"), ("BatchNormalization", "bn_dense"), ("Dense", "dense2")]
    return plan


_SNAKE = {"Conv2D": "conv2d", "SeparableConv2D": "separable_conv2d", "BatchNormalization": "batch_normalization",
          "Bidirectional": "bidirectional", "Dense": "dense"}
_VARS = {"Conv2D": "conv", "Dense": "dense", "SeparableConv2D": "sep", "BatchNormalization": "bn"}


def write_keras_archive(path: Path | str, W: dict, orcai_parameter: dict, name_offset: int = 0) -> None:
    """Write W as a Keras-3-style ``.keras`` zip (config.json with the layer list + model.weights.h5)."""
    import tempfile

    counts: dict[str, int] = {}
    arrays: dict[str, np.ndarray] = {}
    cfg_layers = [{"class_name": "InputLayer", "name": "input_layer", "config": {"name": "input_layer"}}]
    for cls, prefix in _layer_plan(orcai_parameter):
        k = counts.get(cls, 0)
        counts[cls] = k + 1
        idx = k + name_offset
        lname = _SNAKE[cls] + (f"_{idx}" if idx else "")
        cfg_layers.append({"class_name": cls, "name": lname, "config": {"name": lname}})
        if cls == "Bidirectional":
            for d in ("forward", "backward"):
                for i, v in enumerate(_LEGACY_ORDER["lstm"]):
                    arrays[f"/layers/{lname}/{d}_layer/cell/vars/{i}"] = W[f"{prefix}/{d}/{v}"]
        else:
            for i, v in enumerate(_LEGACY_ORDER[_VARS[cls]]):
                arrays[f"/layers/{lname}/vars/{i}"] = W[f"{prefix}/{v}"]
    arrays["/optimizer/vars/0"] = np.zeros((), np.int64)  # iteration counter, ignored by the reader
    with tempfile.TemporaryDirectory() as td:
        h5p = Path(td) / "model.weights.h5"
        write_h5(h5p, arrays)
        config = {"class_name": "Functional", "config": {"name": "functional", "layers": cfg_layers}}
        with zipfile.ZipFile(path, "w", zipfile.ZIP_STORED) as z:
            z.writestr("metadata.json", json.dumps({"keras_version": "3.10.0", "written_by": "orcai_b200.keras_weights"}))
            z.writestr("config.json", json.dumps(config))
            z.write(h5p, "model.weights.h5")


def write_legacy_h5(path: Path | str, W: dict, orcai_parameter: dict) -> None:
    """Write W in the Keras 2 ``save_weights`` layout (``<layer>/<layer>/kernel:0``)."""
    counts: dict[str, int] = {}
    arrays: dict[str, np.ndarray] = {}
    for cls, prefix in _layer_plan(orcai_parameter):
        k = counts.get(cls, 0)
        counts[cls] = k + 1
        lname = _SNAKE[cls] + (f"_{k}" if k else "")
        if cls == "Bidirectional":
            for d in ("forward", "backward"):
                for v in _LEGACY_ORDER["lstm"]:
                    arrays[f"/{lname}/{lname}/{d}_lstm/lstm_cell/{v}:0"] = W[f"{prefix}/{d}/{v}"]
        else:
            for v in _LEGACY_ORDER[_VARS[cls]]:
                arrays[f"/{lname}/{lname}/{v}:0"] = W[f"{prefix}/{v}"]
    write_h5(path, arrays)


__all__ = ["load_keras_archive", "load_weights_h5", "weights_from_h5", "write_keras_archive", "write_legacy_h5", "expected_shapes"]
