"""Keras artefact readers (reference io.py:386-404) on top of the pure-Python HDF5 reader: no keras / h5py in the image."""

import json
import struct
import zlib
from pathlib import Path

import numpy as np
import pytest

from orcai_b200 import runtime
from orcai_b200.hdf5_min import H5File, Hdf5Error, write_h5
from orcai_b200.keras_weights import load_keras_archive, load_weights_h5, write_keras_archive, write_legacy_h5
from orcai_b200.weights import synthetic_weights


@pytest.fixture(scope="module")
def PSW():
    P, S = runtime.bundled_parameters()
    return P, S, synthetic_weights(P, S, seed=4321)


def test_reader_on_a_libhdf5_written_file():
    """scipy ships one MATLAB v7.3 file = HDF5 written by libhdf5 itself (512-byte user block, superblock 0, version-1
    object header with attribute messages, symbol-table group, local heap, contiguous float64 dataset)."""
    import scipy.io

    p = Path(scipy.io.__file__).parent / "matlab" / "tests" / "data" / "testhdf5_7.4_GLNX86.mat"
    if not p.exists():
        pytest.skip("scipy test data not installed")
    f = H5File(p)
    assert f.sb_off == 512 and f.base == 512
    ds = f.datasets()
    assert list(ds) == ["/testdouble"]
    np.testing.assert_allclose(f.read("testdouble").ravel(), np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)


def test_write_read_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    arrays = {
        "/a/b/c": rng.standard_normal((3, 4, 5)).astype(np.float32),
        "/a/b/d": rng.standard_normal(7),
        "/a/e": np.arange(12, dtype=np.int32).reshape(3, 4),
        "/scalar": np.array(5, np.int64),
        "/empty": np.zeros((0, 3), np.float32),
    }
    arrays.update({f"/many/v{i}": np.full((2,), i, np.float32) for i in range(40)})  # several symbol-table nodes
    write_h5(tmp_path / "t.h5", arrays)
    f = H5File(tmp_path / "t.h5")
    got = f.datasets()
    assert set(got) == set(arrays)
    for k, a in arrays.items():
        b = f.read(k)
        assert b.dtype == a.dtype and b.shape == a.shape
        np.testing.assert_array_equal(a, b)
    with pytest.raises(Hdf5Error):
        H5File(b"not an hdf5 file at all" * 100)


def _chunked_file(a: np.ndarray, chunk, deflate: bool, shuffle: bool) -> bytes:
    """Hand-assembled file with ONE chunked dataset (version-3 layout, v1 chunk B-tree, filter pipeline v1) at the root."""
    buf = bytearray(96)

    def alloc(b):
        while len(buf) % 8:
            buf.append(0)
        o = len(buf)
        buf.extend(b)
        return o

    es = a.dtype.itemsize
    entries = []
    for i in range(0, a.shape[0], chunk[0]):
        for j in range(0, a.shape[1], chunk[1]):
            c = np.zeros(chunk, a.dtype)
            blk = a[i : i + chunk[0], j : j + chunk[1]]
            c[: blk.shape[0], : blk.shape[1]] = blk
            raw = c.tobytes()
            if shuffle:
                raw = np.frombuffer(raw, np.uint8).reshape(-1, es).T.tobytes()
            if deflate:
                raw = zlib.compress(raw, 6)
            entries.append(((i, j, 0), len(raw), alloc(raw)))
    node = bytearray(b"TREE" + struct.pack("<BBHQQ", 1, 0, len(entries), 2**64 - 1, 2**64 - 1))
    for offs, size, addr in entries:
        node += struct.pack("<II", size, 0) + struct.pack("<QQQ", *offs) + struct.pack("<Q", addr)
    node += struct.pack("<II", 0, 0) + struct.pack("<QQQ", a.shape[0], a.shape[1], 0)
    bt = alloc(bytes(node))
    space = struct.pack("<BBB5x", 1, 2, 0) + struct.pack("<QQ", *a.shape)
    dt = struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    layout = struct.pack("<BBB", 3, 2, 3) + struct.pack("<Q", bt) + struct.pack("<III", chunk[0], chunk[1], es)
    filt = []
    if shuffle:
        filt.append(struct.pack("<HHHH", 2, 0, 0, 1) + struct.pack("<I", es) + b"\0" * 4)
    if deflate:
        filt.append(struct.pack("<HHHH", 1, 0, 0, 1) + struct.pack("<I", 6) + b"\0" * 4)
    msgs = [(1, space), (3, dt), (8, layout)]
    if filt:
        msgs.append((0x0B, struct.pack("<BB6x", 1, len(filt)) + b"".join(filt)))
    body = bytearray()
    for t, d in msgs:
        d = d + b"\0" * (-len(d) % 8)
        body += struct.pack("<HHB3x", t, len(d), 0) + d
    ds_oh = alloc(struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + bytes(body))
    heap_data = b"\0" * 8 + b"data\0\0\0\0"
    hd = alloc(heap_data)
    heap = alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 2**64 - 1, hd))
    snod = bytearray(b"SNOD" + struct.pack("<BxH", 1, 1) + struct.pack("<QQI4x16x", 8, ds_oh, 0))
    snod += b"\0" * (8 + 40 * 8 - len(snod))
    sn = alloc(bytes(snod))
    tree = alloc(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, 2**64 - 1, 2**64 - 1) + struct.pack("<QQQ", 0, sn, 8) + b"\0" * 512)
    gbody = struct.pack("<HHB3x", 0x11, 16, 0) + struct.pack("<QQ", tree, heap)
    root = alloc(struct.pack("<BxHII4x", 1, 1, 1, len(gbody)) + gbody)
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, 2**64 - 1, len(buf), 2**64 - 1) + struct.pack("<QQI4xQQ", 0, root, 1, tree, heap)
    buf[: len(sb)] = sb
    return bytes(buf)


@pytest.mark.parametrize("deflate,shuffle", [(False, False), (True, False), (True, True)])
def test_chunked_filtered_dataset(deflate, shuffle):
    a = np.random.default_rng(3).standard_normal((10, 7)).astype(np.float32)
    f = H5File(_chunked_file(a, (4, 3), deflate, shuffle))
    np.testing.assert_array_equal(f.read("data"), a)


def test_keras_archive_and_legacy_h5(tmp_path, PSW):
    P, S, W = PSW
    write_keras_archive(tmp_path / "m.keras", W, P)
    got = load_keras_archive(tmp_path / "m.keras", P, S)
    assert set(got) == set(W)
    for k in W:
        np.testing.assert_array_equal(got[k], W[k])
    # auto-generated names that do not start at zero (a model built after others in the same session)
    write_keras_archive(tmp_path / "m7.keras", W, P, name_offset=7)
    got = load_keras_archive(tmp_path / "m7.keras", P, S)
    for k in W:
        np.testing.assert_array_equal(got[k], W[k])
    write_legacy_h5(tmp_path / "model_weights.h5", W, P)
    got = load_weights_h5(tmp_path / "model_weights.h5", P, S)
    for k in W:
        np.testing.assert_array_equal(got[k], W[k])
    # a file that holds another architecture is refused
    P2 = {**P, "model": {**P["model"], "filters": P["model"]["filters"][:-1]}}
    with pytest.raises(ValueError):
        load_keras_archive(tmp_path / "m.keras", P2, S)


def test_load_orcai_model_prefers_like_the_reference(tmp_path, PSW, monkeypatch):
    """io.load_orcai_model: <name>.keras, else model_weights.h5, else ValueError (reference io.py:386-410)."""
    import json

    from orcai_b200 import io as oio
    from orcai_b200 import model as omodel

    P, S, W = PSW
    monkeypatch.delenv("ORCAI_B200_SYNTHETIC_WEIGHTS", raising=False)
    seen = {}

    class FakeModel:
        def __init__(self, p, s, w, device=None):
            seen["W"] = w

    monkeypatch.setattr(omodel, "OrcaiModel", FakeModel)
    d = tmp_path / "orcai-V1"
    d.mkdir()
    (d / "orcai_parameter.json").write_text(json.dumps(P))
    (d / "model_shape.json").write_text(json.dumps(S))
    with pytest.raises(ValueError, match="Couldn't find model weights"):
        oio.load_orcai_model(d)
    write_legacy_h5(d / "model_weights.h5", W, P)
    oio.load_orcai_model(d)
    np.testing.assert_array_equal(seen["W"]["dense2/kernel"], W["dense2/kernel"])
    W2 = synthetic_weights(P, S, seed=9)
    write_keras_archive(d / (P["name"] + ".keras"), W2, P)
    oio.load_orcai_model(d)
    np.testing.assert_array_equal(seen["W"]["dense2/kernel"], W2["dense2/kernel"])


# ---------------------------------------------------------------------------------------------
# conformance with the published layouts (files written by the real producers cannot be fetched offline)
# ---------------------------------------------------------------------------------------------
def _keras3_arrays(W, P, h5_offset=0):
    """model.weights.h5 of Keras 3.10 for res_net_LSTM_arch: saving_lib names every layer of a container
    `<snake_case(class)>[_<k>]` with k counting that class INSIDE the container in `model.layers` order (not the layer's own,
    session-dependent name), variables under `layers/<entry>/vars/<i>` in `layer.weights` order, Bidirectional's cells under
    `forward_layer/cell/vars/<i>` / `backward_layer/cell/vars/<i>`, the optimizer's slots under `optimizer/vars/<i>`."""
    from orcai_b200.keras_weights import _LEGACY_ORDER, _SNAKE, _VARS, _layer_plan

    arrays, counts = {}, {}
    for cls, prefix in _layer_plan(P):
        k = counts.get(cls, 0)
        counts[cls] = k + 1
        entry = _SNAKE[cls] + (f"_{k + h5_offset}" if k + h5_offset else "")
        if cls == "Bidirectional":
            for d in ("forward", "backward"):
                for i, v in enumerate(_LEGACY_ORDER["lstm"]):
                    arrays[f"/layers/{entry}/{d}_layer/cell/vars/{i}"] = W[f"{prefix}/{d}/{v}"]
        else:
            for i, v in enumerate(_LEGACY_ORDER[_VARS[cls]]):
                arrays[f"/layers/{entry}/vars/{i}"] = W[f"{prefix}/{v}"]
    # Adam: iteration counter, learning rate, then two slots per trainable variable (train.py:223 saves with the optimizer)
    arrays["/optimizer/vars/0"] = np.zeros((), np.int64)
    arrays["/optimizer/vars/1"] = np.float32(1e-3) * np.ones((), np.float32)
    for j, (kname, a) in enumerate(sorted(W.items())[:6]):
        arrays[f"/optimizer/vars/{2 + j}"] = np.zeros_like(a)
    return arrays


def test_keras3_archive_layout_conformance(tmp_path, PSW):
    """A `.keras` archive laid out as Keras 3.10 writes it: config.json carries the layers' own names (session-dependent
    suffixes, plus every variable-less layer), model.weights.h5 the per-container class-counter names and the optimizer slots."""
    import zipfile

    from orcai_b200.hdf5_min import write_h5
    from orcai_b200.keras_weights import _SNAKE, _layer_plan, load_keras_archive

    P, S, W = PSW
    h5p = tmp_path / "model.weights.h5"
    write_h5(h5p, _keras3_arrays(W, P))
    # the layer list of the functional model: own names with a session offset, interleaved with layers that own no variables
    cfg_layers = [{"class_name": "InputLayer", "name": "input_layer_3", "config": {"name": "input_layer_3"}}]
    counts = {}
    for cls, _prefix in _layer_plan(P):
        k = counts.get(cls, 0)
        counts[cls] = k + 1
        cfg_layers.append({"class_name": cls, "name": f"{_SNAKE[cls]}_{k + 17}", "config": {"name": f"{_SNAKE[cls]}_{k + 17}"}})
        if cls == "BatchNormalization":
            cfg_layers.append({"class_name": "ReLU", "name": f"re_lu_{k + 40}", "config": {}})
        if cls == "Conv2D" and k > 0:
            cfg_layers += [{"class_name": "MaxPooling2D", "name": f"max_pooling2d_{k}", "config": {}}, {"class_name": "Add", "name": f"add_{k}", "config": {}}]
    arch = tmp_path / "orcai-v1.keras"
    with zipfile.ZipFile(arch, "w") as z:
        z.writestr("metadata.json", json.dumps({"keras_version": "3.10.0", "date_saved": "2025-06-01@12:00:00"}))
        z.writestr("config.json", json.dumps({"module": "keras", "class_name": "Functional", "config": {"name": "functional_2", "layers": cfg_layers}}))
        z.write(h5p, "model.weights.h5")
    got = load_keras_archive(arch, P, S)
    assert set(got) == set(W)
    for k in W:
        np.testing.assert_array_equal(got[k], W[k])


def test_ambiguous_layer_order_fails_loudly(tmp_path, PSW):
    """Legacy model_weights.h5 / bare weight files carry no layer list this reader parses; the numeric name suffix is the only
    order.  A renamed layer, a gap or a second name family must raise instead of silently swapping same-shaped layers."""
    from orcai_b200.hdf5_min import Hdf5Error, write_h5
    from orcai_b200.keras_weights import load_weights_h5

    P, S, W = PSW
    good = _keras3_arrays(W, P, h5_offset=5)                # an auto-generated family that starts at a session offset is fine
    write_h5(tmp_path / "ok.h5", good)
    got = load_weights_h5(tmp_path / "ok.h5", P, S)
    np.testing.assert_array_equal(got["block1/bn2/gamma"], W["block1/bn2/gamma"])

    def renamed(old, new):
        return {k.replace(f"/layers/{old}/", f"/layers/{new}/"): v for k, v in good.items()}

    for name, arrays in (("renamed", renamed("batch_normalization_7", "my_norm")),            # a custom layer name
                         ("gap", renamed("batch_normalization_7", "batch_normalization_99")),  # re-created layer: the suffix jumps
                         ("swapped-family", renamed("separable_conv2d_6", "sepconv_6"))):
        write_h5(tmp_path / f"{name}.h5", arrays)
        with pytest.raises(Hdf5Error, match="order of the"):
            load_weights_h5(tmp_path / f"{name}.h5", P, S)
