"""File formats either side of the hot path.

* ``read_json`` / ``write_vector_to_json``           reference ``src/orcAI/io.py:221-238, 259-274``
* ``save_as_zarr`` (zarr v3 directory store, float32, chunks (2000, n_freq), bytes+gzip)
                                                      reference ``src/orcAI/io.py:296-331`` (zarr 3.0.8 is not
  installable here, so the store is written by hand from the zarr v3 specification; ``read_zarr`` is
  the matching reader used by the tests)
* ``load_orcai_model``                                reference ``src/orcAI/io.py:357-410``
"""

from __future__ import annotations

import gzip
import json
import os
import warnings
import zlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np


def read_json(filename: Path | str) -> dict:
    with open(filename, "r") as f:
        return json.load(f)


def _jsonable(v):
    if isinstance(v, np.generic):
        return v.item()
    return v


def write_vector_to_json(vector, filename: Path | str) -> None:
    """Equally spaced vector in short form: {"min", "max", "length"} (indent 4, like the reference)."""
    d = {"min": _jsonable(vector[0]), "max": _jsonable(vector[-1]), "length": len(vector)}
    with open(filename, "w") as f:
        json.dump(d, f, indent=4)


def generate_times_from_spectrogram(filename: Path | str) -> np.ndarray:
    d = read_json(filename)
    return np.linspace(d["min"], d["max"], d["length"])


# ---------------------------------------------------------------------------------------------
# zarr v3 (array at the store root)
# ---------------------------------------------------------------------------------------------
_GZIP_LEVEL = 5  # numcodecs / zarr default for {"name": "gzip", "configuration": {}}


def _gzip_chunk(buf: bytes) -> bytes:
    co = zlib.compressobj(_GZIP_LEVEL, zlib.DEFLATED, 31)  # wbits 31 = gzip container
    return co.compress(buf) + co.flush()


def save_as_zarr(obj: np.ndarray, filename: Path | str, chunk_rows: int = 2000, threads: int | None = None) -> None:
    """Write a 2-D array as a zarr v3 float32 array with chunks (chunk_rows, ncols), codecs bytes(little)+gzip."""
    a = np.ascontiguousarray(obj, dtype="<f4")
    if a.ndim != 2:
        raise ValueError("save_as_zarr expects a 2-D array")
    root = Path(filename)
    if root.exists():
        import shutil

        shutil.rmtree(root)  # zarr.open(mode="w") replaces an existing store
    root.mkdir(parents=True)
    rows, cols = a.shape
    meta = {
        "shape": [rows, cols],
        "data_type": "float32",
        "chunk_grid": {"name": "regular", "configuration": {"chunk_shape": [chunk_rows, cols]}},
        "chunk_key_encoding": {"name": "default", "configuration": {"separator": "/"}},
        "fill_value": 0.0,
        "codecs": [
            {"name": "bytes", "configuration": {"endian": "little"}},
            {"name": "gzip", "configuration": {"level": _GZIP_LEVEL}},
        ],
        "attributes": {},
        "zarr_format": 3,
        "node_type": "array",
        "storage_transformers": [],
    }
    (root / "zarr.json").write_text(json.dumps(meta, indent=2))
    n_chunks = -(-rows // chunk_rows) if rows else 0

    def write_chunk(i: int) -> None:
        blk = a[i * chunk_rows : (i + 1) * chunk_rows]
        if blk.shape[0] < chunk_rows:  # edge chunks are stored full-size, padded with the fill value
            pad = np.zeros((chunk_rows, cols), dtype="<f4")
            pad[: blk.shape[0]] = blk
            blk = pad
        if not blk.any():  # write_empty_chunks=False: all-fill chunks are omitted
            return
        d = root / "c" / str(i)
        d.mkdir(parents=True, exist_ok=True)
        (d / "0").write_bytes(_gzip_chunk(blk.tobytes()))

    workers = threads or gzip_workers()
    with ThreadPoolExecutor(max_workers=max(1, workers)) as ex:  # zlib releases the GIL
        list(ex.map(write_chunk, range(n_chunks)))


def gzip_workers() -> int:
    """Threads that compress the chunks of one store (zlib releases the interpreter lock): the cores this process may use, at most 32."""
    return min(32, len(os.sched_getaffinity(0)))


def read_zarr(filename: Path | str) -> np.ndarray:
    """Reader for stores written by save_as_zarr (float32, bytes+gzip)."""
    root = Path(filename)
    meta = read_json(root / "zarr.json")
    rows, cols = meta["shape"]
    cr, cc = meta["chunk_grid"]["configuration"]["chunk_shape"]
    assert cc == cols and meta["data_type"] == "float32"
    out = np.full((rows, cols), meta["fill_value"], dtype=np.float32)
    for i in range(-(-rows // cr)):
        p = root / "c" / str(i) / "0"
        if p.exists():
            blk = np.frombuffer(gzip.decompress(p.read_bytes()), dtype="<f4").reshape(cr, cols)
            out[i * cr : (i + 1) * cr] = blk[: min(cr, rows - i * cr)]
    return out


# ---------------------------------------------------------------------------------------------
# model artefact
# ---------------------------------------------------------------------------------------------
_MODEL_CACHE: dict = {}


def load_orcai_model(model_dir: Path | str, device: int | None = None):
    """Load a model directory -> (model, orcai_parameter, shape).

    The directory holds ``orcai_parameter.json`` and ``model_shape.json`` like the reference's.  Weights are
    looked up as ``<name>.weights.npz`` (this package's container, see ``orcai_b200/weights.py``), then - like the
    reference, ``io.py:386-404`` - as ``<name>.keras`` (Keras 3 archive) and as the legacy ``model_weights.h5``; both are
    read by ``orcai_b200.keras_weights`` on top of the pure-Python HDF5 reader ``orcai_b200.hdf5_min`` (no keras / h5py).
    Only when the directory holds NO weight file at all, ``ORCAI_B200_SYNTHETIC_WEIGHTS=<seed>`` substitutes seeded
    synthetic weights (benchmarks / smoke tests: the packaged orcai-v1.keras blob is absent from the reference mount) and
    says so with a warning; real weights are never shadowed by the variable.
    """
    from orcai_b200.model import OrcaiModel
    from orcai_b200.weights import load_npz, synthetic_weights

    from orcai_b200.model import precision_from_env
    from orcai_b200.runtime import default_device

    model_dir = Path(model_dir)
    # a process that annotates table after table (worker processes, services) loads a model directory once per (device, arithmetic):
    # the key holds the size and mtime of every file in the directory, so edited weights or parameters are re-read
    try:
        files_sig = tuple(sorted((f.name, f.stat().st_size, f.stat().st_mtime_ns) for f in model_dir.iterdir() if f.is_file()))
    except OSError:
        files_sig = ()
    cache_key = (str(model_dir.resolve()), files_sig, default_device() if device is None else int(device), precision_from_env(),
                 os.environ.get("ORCAI_B200_SYNTHETIC_WEIGHTS"), os.environ.get("ORCAI_B200_CALIBRATION"))
    hit = _MODEL_CACHE.get(cache_key)
    if hit is not None:
        return hit
    orcai_parameter = read_json(model_dir.joinpath("orcai_parameter.json"))
    shape = read_json(model_dir.joinpath("model_shape.json"))
    name = orcai_parameter["name"]
    npz = model_dir.joinpath(name + ".weights.npz")
    synth = os.environ.get("ORCAI_B200_SYNTHETIC_WEIGHTS")
    if npz.exists():
        W = load_npz(npz)
    elif model_dir.joinpath(name + ".keras").exists():
        from orcai_b200.keras_weights import load_keras_archive

        W = load_keras_archive(model_dir.joinpath(name + ".keras"), orcai_parameter, shape)
    elif model_dir.joinpath("model_weights.h5").exists():
        from orcai_b200.keras_weights import load_weights_h5

        W = load_weights_h5(model_dir.joinpath("model_weights.h5"), orcai_parameter, shape)
    elif synth is not None:
        seed = int(synth) if synth.strip().lstrip("-").isdigit() else 1234
        warnings.warn(f"{model_dir} holds no weight file: using SYNTHETIC random weights (ORCAI_B200_SYNTHETIC_WEIGHTS, seed {seed}); "
                      "the label files of this run are meaningless", RuntimeWarning, stacklevel=2)
        W = synthetic_weights(orcai_parameter, shape, seed=seed)
    else:
        raise ValueError(f"Couldn't find model weights ({name}.weights.npz, model_weights.h5) or keras model file in {model_dir}")
    model = OrcaiModel(orcai_parameter, shape, W, device=device)
    _MODEL_CACHE[cache_key] = (model, orcai_parameter, shape)
    return model, orcai_parameter, shape
