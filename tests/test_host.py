"""Host-side logic that needs no GPU: WAV reader, zarr writer, parameter block, writers, CLI surface."""

import gzip
import json
import re
from pathlib import Path

import numpy as np
import pandas as pd
import pytest
from click.testing import CliRunner

from oracle import postprocess_oracle as po
from orcai_b200 import _lib, cli, io, predict, runtime, wavio
from orcai_b200.synth import synth_pcm16

ROOT = Path(__file__).resolve().parent.parent


def test_abi_exports_every_declared_symbol():
    header = (ROOT / "include" / "orcai_b200.h").read_text()
    declared = set(re.findall(r"\b(orcai_[a-z_0-9]+)\s*\(", header))
    lib = _lib.load_library()
    assert declared == set(_lib.exported_symbols())
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.orcai_version() >= 100
    assert lib.orcai_num_frames(28800000, 256) == 112501 and lib.orcai_num_snippets(112501, 736) == 304
    assert lib.orcai_num_snippets(735, 736) == 0 and lib.orcai_num_snippets(736, 736) == 1


def test_params_block(params):
    P, S = params
    p = _lib.params_from_dicts(P, S)
    assert (p.n_fft, p.hop, p.band_lo, p.band_hi, p.n_freq, p.snippet_len, p.n_labels, p.n_blocks) == (512, 256, 0, 171, 171, 736, 7, 4)
    assert p.q_lo == 0.01 and p.q_hi == 0.9990000000000001
    assert list(p.filters)[:4] == [30, 40, 50, 60] and p.lstm_units == 128
    assert runtime.shape_for(P) == S


def test_no_cpu_fallback(params):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    P, S = params
    with pytest.raises(_lib.OrcaiError) as e:
        _lib.Context(P, S, 0)
    assert "no CPU fallback" in str(e.value)


@pytest.mark.parametrize("kind", ["pcm16", "pcm16_stereo", "pcm24", "pcm32", "float32", "u8"])
def test_wav_reader(tmp_path, kind):
    import struct

    rng = np.random.default_rng(1)
    n = 1000
    x = rng.integers(-32768, 32767, size=(n, 2)).astype(np.int16)
    path = tmp_path / f"{kind}.wav"

    def write(fmt_tag, ch, bits, payload):
        hdr = struct.pack("<HHIIHH", fmt_tag, ch, 48000, 48000 * ch * bits // 8, ch * bits // 8, bits)
        with open(path, "wb") as f:
            f.write(b"RIFF" + struct.pack("<I", 36 + len(payload)) + b"WAVE" + b"fmt " + struct.pack("<I", 16) + hdr)
            f.write(b"LIST" + struct.pack("<I", 4) + b"abcd")  # an unrelated chunk to skip
            f.write(b"data" + struct.pack("<I", len(payload)) + payload)

    if kind == "pcm16":
        wavio.write_wav_pcm16(path, x[:, 0])
        y, sr, ch = wavio.read_wav(path)
        assert y.dtype == np.int16 and np.array_equal(y, x[:, 0]) and (sr, ch) == (48000, 1)
    elif kind == "pcm16_stereo":
        wavio.write_wav_pcm16(path, x)
        y, _, ch = wavio.read_wav(path, channel=2)
        assert ch == 2 and np.array_equal(y, x[:, 1])
        with pytest.raises(IndexError):
            wavio.read_wav(path, channel=3)
    elif kind == "pcm24":
        v = x[:, 0].astype(np.int32) * 256 + 17
        b = np.stack([(v & 0xFF), (v >> 8) & 0xFF, (v >> 16) & 0xFF], axis=1).astype(np.uint8).tobytes()
        write(1, 1, 24, b)
        y, _, _ = wavio.read_wav(path)
        np.testing.assert_array_equal(y, (v / 8388608.0).astype(np.float32))
    elif kind == "pcm32":
        v = x[:, 0].astype(np.int32) * 65536
        write(1, 1, 32, v.astype("<i4").tobytes())
        y, _, _ = wavio.read_wav(path)
        np.testing.assert_array_equal(y, (v / 2147483648.0).astype(np.float32))
    elif kind == "float32":
        v = (x[:, 0] / 32768.0).astype("<f4")
        write(3, 1, 32, v.tobytes())
        y, _, _ = wavio.read_wav(path)
        np.testing.assert_array_equal(y, v)
    else:
        v = rng.integers(0, 255, n).astype(np.uint8)
        write(1, 1, 8, v.tobytes())
        y, _, _ = wavio.read_wav(path)
        np.testing.assert_array_equal(y, ((v.astype(np.float32) - 128) / 128).astype(np.float32))


def test_wav_reader_rejects_garbage(tmp_path):
    p = tmp_path / "x.wav"
    p.write_bytes(b"not a wave file at all")
    with pytest.raises(ValueError):
        wavio.read_wav(p)


def test_zarr_v3_store_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.random((4321, 171), dtype=np.float32)
    a[2000:4000] = 0.0  # an all-fill chunk is omitted from the store
    io.save_as_zarr(a, tmp_path / "spectrogram.zarr")
    meta = json.loads((tmp_path / "spectrogram.zarr" / "zarr.json").read_text())
    assert meta["zarr_format"] == 3 and meta["node_type"] == "array" and meta["shape"] == [4321, 171] and meta["data_type"] == "float32"
    assert meta["chunk_grid"]["configuration"]["chunk_shape"] == [2000, 171]
    assert [c["name"] for c in meta["codecs"]] == ["bytes", "gzip"]
    assert (tmp_path / "spectrogram.zarr" / "c" / "0" / "0").exists() and not (tmp_path / "spectrogram.zarr" / "c" / "1").exists()
    raw = gzip.decompress((tmp_path / "spectrogram.zarr" / "c" / "2" / "0").read_bytes())
    assert len(raw) == 2000 * 171 * 4  # edge chunk stored full size
    np.testing.assert_array_equal(io.read_zarr(tmp_path / "spectrogram.zarr"), a)
    t = np.arange(4321) * 256 / 48000.0
    io.write_vector_to_json(t, tmp_path / "times.json")
    assert json.loads((tmp_path / "times.json").read_text()) == {"min": 0.0, "max": float(t[-1]), "length": 4321}
    np.testing.assert_allclose(io.generate_times_from_spectrogram(tmp_path / "times.json"), t, rtol=0, atol=1e-9)


def test_label_writer_matches_oracle_writer(golden_dir, tmp_path):
    fx = json.loads((golden_dir / "postprocess_seed11.json").read_text())
    g = np.load(golden_dir / "postprocess_seed11.npz")
    labels = predict.compute_labels(list(g["starts"]), list(g["stops"]), fx["labels"], 16, "*")
    assert list(labels.columns) == ["start", "stop", "label"] and labels["start"].dtype == np.int64
    predict.save_predictions(labels, tmp_path / "out.txt", fx["delta_t"])
    assert (tmp_path / "out.txt").read_bytes() == fx["tsv"].encode()
    for case in fx["writer_cases"]:
        rows = case["rows"]
        df = predict.compute_labels([r[0] // 16 for r in rows], [r[1] // 16 for r in rows], [r[2][:-1] for r in rows], 16, "*")
        assert predict.labels_to_tsv(df, fx["delta_t"]) == case["tsv"]
    empty = predict.compute_labels([], [], [], 16, "*")
    assert predict.labels_to_tsv(empty, fx["delta_t"]) == "start\tstop\tlabel\n"


def test_probability_writer_and_filter(tmp_path):
    agg = np.random.default_rng(0).random((50, 7))
    P, _ = runtime.bundled_parameters()
    predict.save_prediction_probabilities(agg, P, 256 / 48000, tmp_path / "rec_predicted.txt")
    out = tmp_path / "rec_predicted_probabilities.csv.gz"
    assert gzip.decompress(out.read_bytes()).decode() == po.probabilities_csv(agg, P["calls"], 256 / 48000)
    df = pd.DataFrame({"start": [0, 0, 32], "stop": [160, 16, 6000], "label": ["BR*", "SS*", "X*"]})
    lim = {"default": [0.5, None], "SS": [0, 0.05], "BR": [None, 1.0]}
    from orcai_b200.auxiliary import Messenger

    kept = predict.filter_predictions(df, 256 / 48000, lim, msgr=Messenger(verbosity=0))
    assert [tuple(r) for r in kept[["start", "stop", "label"]].itertuples(index=False)] == po.filter_rows(
        [(0, 160, "BR*"), (0, 16, "SS*"), (32, 6000, "X*")], 256 / 48000, lim
    )


def test_cli_surface():
    """Option names, short flags and defaults of the two drop-in commands (reference cli.py:93-184, 359-416)."""
    cmds = cli.cli.commands
    assert {"predict", "create-spectrograms", "filter-predictions"} <= set(cmds)
    opts = {o.name: o for o in cmds["predict"].params}
    assert set(opts) == {"recording_path", "channel", "model", "model_dir", "output_path", "overwrite", "save_probabilities",
                         "base_dir_recording", "call_duration_limits", "label_suffix", "verbosity"}
    assert opts["channel"].default == 1 and opts["output_path"].default == "default" and opts["label_suffix"].default == "*"
    assert opts["verbosity"].default == 2 and opts["model"].default == "orcai-v1"
    assert set(opts["model_dir"].opts) == {"--model_dir", "-md"} and set(opts["save_probabilities"].opts) == {"--save_probabilities", "-sp"}
    assert set(opts["call_duration_limits"].opts) == {"--call_duration_limits", "-cdl"} and set(opts["base_dir_recording"].opts) == {"--base_dir_recording", "-bdr"}
    so = {o.name: o for o in cmds["create-spectrograms"].params}
    assert set(so) == {"recording_table_path", "output_dir", "base_dir_recording", "orcai_parameter", "include_not_annotated",
                       "include_no_possible_annotations", "overwrite", "verbosity"}
    assert set(so["include_not_annotated"].opts) == {"--include_not_annotated", "-en"}
    assert set(so["include_no_possible_annotations"].opts) == {"--include_no_possible_annotations", "-enp"}
    r = CliRunner().invoke(cli.cli, ["predict", "--help"])
    assert r.exit_code == 0 and "RECORDING_PATH" in r.output


def test_predict_rejects_unknown_suffix(tmp_path, monkeypatch):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("model loading needs the GPU context")
    monkeypatch.setenv("ORCAI_B200_SYNTHETIC_WEIGHTS", "1234")
    p = tmp_path / "rec.flac"
    p.write_bytes(b"x")
    with pytest.raises(ValueError, match="wav or csv"):
        predict.predict(p, verbosity=0)


def test_synthetic_audio_is_deterministic():
    a = synth_pcm16(1.0, seed=7)
    b = synth_pcm16(1.0, seed=7)
    assert a.dtype == np.int16 and len(a) == 48000 and np.array_equal(a, b) and not np.array_equal(a, synth_pcm16(1.0, seed=8))


def test_precision_selector(monkeypatch):
    from orcai_b200 import model

    monkeypatch.delenv("ORCAI_B200_PRECISION", raising=False)
    # the shipped default is the path inside the 1e-3 gate; the single-fp16 "fast" path is opt-in
    assert model.precision_from_env() == "precise" == model.DEFAULT_PRECISION and model.PRECISION_PATHS["precise"] == 4
    monkeypatch.setenv("ORCAI_B200_PRECISION", "fast")
    assert model.precision_from_env() == "fast" and model.PRECISION_PATHS["fast"] == 3
    monkeypatch.setenv("ORCAI_B200_PRECISION", "Reference ")
    assert model.precision_from_env() == "reference" and model.PRECISION_PATHS["reference"] == 0
    monkeypatch.setenv("ORCAI_B200_PRECISION", "accurate")
    assert model.precision_from_env() == "accurate" and model.PRECISION_PATHS["accurate"] == 3 and model.PRECISION_CONV0["accurate"] == 0
    assert model.PRECISION_CONV0["fast"] == 1
    monkeypatch.setenv("ORCAI_B200_PRECISION", "bf16")
    with pytest.raises(ValueError, match="ORCAI_B200_PRECISION"):
        model.precision_from_env()


def test_predict_stream_order_of_calls():
    """Context.predict_stream: prefetch(k+1) is issued after swap(k) and before predict(k); results come back in order."""
    from orcai_b200._lib import Context

    log = []

    class Fake:
        prefetch_pcm = lambda self, pcm: log.append(("prefetch", int(pcm[0])))  # noqa: E731
        swap_pcm = lambda self: log.append(("swap",))  # noqa: E731

        def predict_pcm(self, pcm, threshold=0.5, want_agg=True, resident=False):
            assert resident
            log.append(("predict", int(pcm[0])))
            return int(pcm[0])

    recs = [np.array([k], np.int16) for k in range(3)]
    out = list(Context.predict_stream(Fake(), iter(recs)))
    assert out == [0, 1, 2]
    assert log == [("prefetch", 0), ("swap",), ("prefetch", 1), ("predict", 0), ("swap",), ("prefetch", 2), ("predict", 1), ("swap",), ("predict", 2)]
    assert list(Context.predict_stream(Fake(), iter([]))) == []


def test_host_resampler_preserves_a_tone():
    from orcai_b200.spectrogram import resample

    sr, target = 44100, 48000
    t = np.arange(sr) / sr
    x = (0.5 * np.sin(2 * np.pi * 1000.0 * t) * 32767).astype(np.int16)
    y = resample(x, sr, target)
    assert y.dtype == np.float32 and abs(len(y) - target) <= 1
    ref = 0.5 * np.sin(2 * np.pi * 1000.0 * np.arange(len(y)) / target) * 32767 / 32768
    assert np.abs(y[2000:-2000] - ref[2000:-2000]).max() < 2e-3
    z = resample(x.astype(np.float32) / 32768.0, sr, target)
    np.testing.assert_allclose(z, y, atol=1e-6)


def test_fast_label_table_equals_compute_labels():
    """predict._labels_of (array operations on label indices, used by predict_wav / table mode) == the reference-shaped
    compute_labels (predict.py:320-340) on the same segments, including ties, every suffix and the empty result; the cached
    time formatting writes the same text as the uncached one."""
    import pandas as pd

    from orcai_b200 import predict as op, runtime
    from orcai_b200.auxiliary import Messenger

    P, _ = runtime.bundled_parameters()
    calls = P["calls"]
    rng = np.random.default_rng(1)
    m = Messenger(verbosity=0)
    for suf in ("*", "", None, "_x"):
        for n in (0, 1, 7, 3000):
            sta = np.sort(rng.integers(0, 42000, n))
            sto = sta + rng.integers(0, 30, n)
            lab = rng.integers(0, len(calls), n).astype(np.int32)
            if n > 20:
                sta[10:20] = sta[10]
                sto[10:20] = sto[10]
            a = op._labels_of(lab, sta, sto, P, suf, m)
            b = op.compute_labels(sta.tolist(), sto.tolist(), [calls[i] for i in lab.tolist()], 16, suf)
            pd.testing.assert_frame_equal(a, b)
            dt = 256 / 48000
            text = op.labels_to_tsv(a, dt)
            assert text == op.labels_to_tsv(b, dt)
            if n:
                want = "start\tstop\tlabel\n" + "".join(
                    f"{float(x)!r}\t{float(y)!r}\t{l}\n"
                    for x, y, l in zip(np.round(b["start"].to_numpy() * np.float64(dt), 4), np.round(b["stop"].to_numpy() * np.float64(dt), 4), b["label"])
                )
                assert text == want


def test_wav_reader_into_caller_memory(tmp_path):
    """read_wav(alloc=...) (table mode decodes into page-locked buffers): same samples as the plain reader for mono PCM16
    (read straight into the buffer) and for every format that is decoded first (stereo, 24-bit, float)."""
    import struct

    from orcai_b200 import wavio

    rng = np.random.default_rng(4)
    taken = []

    def alloc(nbytes):
        a = np.full(nbytes + 64, 0xAB, dtype=np.uint8)   # larger than asked for, like a recycled pool buffer
        taken.append(a)
        return a

    mono = rng.integers(-32768, 32767, 48000 + 7, dtype=np.int16)
    wavio.write_wav_pcm16(tmp_path / "mono.wav", mono)
    stereo = rng.integers(-32768, 32767, (5000, 2), dtype=np.int16)
    wavio.write_wav_pcm16(tmp_path / "stereo.wav", stereo)
    f32 = rng.standard_normal(3000).astype("<f4")
    with open(tmp_path / "f32.wav", "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + f32.nbytes) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 3, 1, 48000, 48000 * 4, 4, 32))
        f.write(b"data" + struct.pack("<I", f32.nbytes) + f32.tobytes())
    raw24 = rng.integers(0, 256, 3 * 2000, dtype=np.uint8)
    with open(tmp_path / "p24.wav", "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + raw24.size) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, 48000, 48000 * 3, 3, 24))
        f.write(b"data" + struct.pack("<I", raw24.size) + raw24.tobytes())
    for name, ch in (("mono.wav", 1), ("stereo.wav", 2), ("f32.wav", 1), ("p24.wav", 1)):
        plain, sr, n_ch = wavio.read_wav(tmp_path / name, ch)
        taken.clear()
        got, sr2, n_ch2 = wavio.read_wav(tmp_path / name, ch, alloc=alloc)
        assert (sr, n_ch) == (sr2, n_ch2) and got.dtype == plain.dtype and len(taken) == 1
        np.testing.assert_array_equal(got, plain)
        assert np.shares_memory(got, taken[0])            # the samples live in the caller's buffer
    np.testing.assert_array_equal(wavio.read_wav(tmp_path / "mono.wav", 1, alloc=alloc)[0], mono)
    # a truncated data chunk is reported, not silently zero-filled
    data = (tmp_path / "mono.wav").read_bytes()
    (tmp_path / "cut.wav").write_bytes(data[: len(data) - 1000])
    cut_plain = wavio.read_wav(tmp_path / "cut.wav", 1)[0]
    cut_alloc = wavio.read_wav(tmp_path / "cut.wav", 1, alloc=alloc)[0]
    np.testing.assert_array_equal(cut_plain, cut_alloc)
    # the first recording of a table is read by concurrent slices (any slice count, ragged tail): the same samples
    big = rng.integers(-32768, 32767, (1 << 22) + 12345, dtype=np.int16)
    wavio.write_wav_pcm16(tmp_path / "big.wav", big)
    for th in (2, 3, 4, 7):
        np.testing.assert_array_equal(wavio.read_wav(tmp_path / "big.wav", 1, alloc=alloc, read_threads=th)[0], big)


def test_model_rebinds_a_shared_context(monkeypatch):
    """Contexts are shared per (device, parameters); a model must re-bind its own weights and arithmetic before it predicts,
    so that loading a second model never makes the first one predict with the other's weights (model.bind())."""
    from orcai_b200 import model as model_mod, runtime
    from orcai_b200.weights import synthetic_weights

    calls = []

    class FakeCtx:
        owner = None

        def load_weights(self, W):
            calls.append(("load", float(W["dense2/bias"][0])))
            self.owner = None

        def set_option(self, k, v):
            calls.append((k, v))

        def calibrate(self, *a, **k):
            calls.append(("calibrate",))

    fake = FakeCtx()
    monkeypatch.setattr(model_mod, "get_context", lambda *a, **k: fake)
    P, S = runtime.bundled_parameters()
    W1 = synthetic_weights(P, S, seed=1)
    W2 = {k: v.copy() for k, v in W1.items()}
    W2["dense2/bias"] = W2["dense2/bias"] + 1.0
    m1 = model_mod.OrcaiModel(P, S, W1, precision="precise")
    m2 = model_mod.OrcaiModel(P, S, W2, precision="fast")
    assert ("net_path", 4) in calls and ("net_path", 3) in calls and ("calibrate",) in calls
    del calls[:]
    assert m2.bind() is fake and calls == []                     # still the owner: nothing to do
    m1.bind()                                                    # m1 takes the context back: its weights, its path, no calibration
    assert calls[0] == ("load", float(W1["dense2/bias"][0])) and ("net_path", 4) in calls and ("calibrate",) not in calls
    assert "precise" in m1.describe() and "calibration" in m2.describe()
    with pytest.raises(ValueError):
        model_mod.OrcaiModel(P, S, W1, precision="bf16")


def _spec_read_zarr_v3(root: Path) -> np.ndarray:
    """A reader written from the Zarr v3 core specification alone (no knowledge of orcai_b200.io): zarr.json -> array.
    Supports what zarr-python 3.0.8 emits for `zarr.open(mode="w", shape, chunks, dtype="float32")` with the gzip default
    compressor the reference configures (io.py:320-330): regular chunk grid, default chunk-key encoding, bytes + gzip codecs."""
    import zlib

    meta = json.loads((root / "zarr.json").read_text())
    required = {"zarr_format", "node_type", "shape", "data_type", "chunk_grid", "chunk_key_encoding", "fill_value", "codecs"}
    assert required <= set(meta), sorted(required - set(meta))
    assert meta["zarr_format"] == 3 and meta["node_type"] == "array"
    assert set(meta) <= required | {"attributes", "storage_transformers", "dimension_names"}   # nothing a v3 reader would not know
    assert meta.get("storage_transformers", []) == [] and isinstance(meta.get("attributes", {}), dict)
    dtype = {"float32": "f4", "float64": "f8", "int16": "i2", "int32": "i4"}[meta["data_type"]]
    grid = meta["chunk_grid"]
    assert grid["name"] == "regular"
    cshape = tuple(grid["configuration"]["chunk_shape"])
    enc = meta["chunk_key_encoding"]
    assert enc["name"] in ("default", "v2")
    sep = enc.get("configuration", {}).get("separator", "/" if enc["name"] == "default" else ".")
    codecs = meta["codecs"]
    kinds = [c["name"] for c in codecs]
    assert kinds.count("bytes") == 1 and kinds.index("bytes") == 0, "exactly one array->bytes codec, first"
    endian = codecs[0].get("configuration", {}).get("endian", "little")
    shape = tuple(meta["shape"])
    out = np.full(shape, meta["fill_value"], dtype=dtype)
    n_chunks = [-(-s // c) for s, c in zip(shape, cshape)]
    for idx in np.ndindex(*n_chunks):
        key = ("c" + sep + sep.join(map(str, idx))) if enc["name"] == "default" else sep.join(map(str, idx))
        f = root / key
        if not f.exists():
            continue                                          # missing chunk = fill value
        raw = f.read_bytes()
        for c in reversed(codecs[1:]):                        # bytes -> bytes codecs, undone last to first
            assert c["name"] == "gzip" and 0 <= c["configuration"].get("level", 5) <= 9
            raw = zlib.decompress(raw, 31)                    # gzip container (RFC 1952)
        blk = np.frombuffer(raw, dtype=("<" if endian == "little" else ">") + dtype).reshape(cshape)   # edge chunks are full-sized
        sl = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, cshape, shape))
        out[sl] = blk[tuple(slice(0, x.stop - x.start) for x in sl)]
    return out


def test_zarr_v3_store_conforms_to_the_specification(tmp_path):
    """The store `create-spectrograms` writes, read back by a reader derived from the Zarr v3 specification only, and the exact
    metadata zarr-python 3.0.8 produces for the reference's call (keys, codec chain, gzip level 5, chunk keys c/<i>/0)."""
    rng = np.random.default_rng(3)
    a = rng.random((4500, 171), dtype=np.float32)
    a[2000:4000] = 0.0
    root = tmp_path / "spectrogram.zarr"
    io.save_as_zarr(a, root)
    np.testing.assert_array_equal(_spec_read_zarr_v3(root), a)
    meta = json.loads((root / "zarr.json").read_text())
    assert meta == {
        "shape": [4500, 171], "data_type": "float32",
        "chunk_grid": {"name": "regular", "configuration": {"chunk_shape": [2000, 171]}},
        "chunk_key_encoding": {"name": "default", "configuration": {"separator": "/"}},
        "fill_value": 0.0,
        "codecs": [{"name": "bytes", "configuration": {"endian": "little"}}, {"name": "gzip", "configuration": {"level": 5}}],
        "attributes": {}, "zarr_format": 3, "node_type": "array", "storage_transformers": [],
    }
    files = sorted(str(f.relative_to(root)) for f in root.rglob("*") if f.is_file())
    assert files == ["c/0/0", "c/2/0", "zarr.json"]          # the all-zero chunk is not stored (write_empty_chunks=False)
    assert (root / "c" / "0" / "0").read_bytes()[:3] == b"\x1f\x8b\x08"   # gzip magic + deflate
    # and the other direction: a store with different but valid choices (big-endian bytes codec, another gzip level, an attribute,
    # dimension names) is read by orcai_b200's reader
    alt = tmp_path / "alt.zarr"
    (alt / "c" / "1").mkdir(parents=True)
    b = rng.random((2500, 171), dtype=np.float32)
    (alt / "zarr.json").write_text(json.dumps({
        "zarr_format": 3, "node_type": "array", "shape": [2500, 171], "data_type": "float32", "fill_value": 0.0,
        "chunk_grid": {"name": "regular", "configuration": {"chunk_shape": [2000, 171]}},
        "chunk_key_encoding": {"name": "default", "configuration": {"separator": "/"}},
        "codecs": [{"name": "bytes", "configuration": {"endian": "little"}}, {"name": "gzip", "configuration": {"level": 1}}],
        "attributes": {"note": "x"}, "dimension_names": ["time", "frequency"]}))
    pad = np.zeros((2000, 171), np.float32)
    pad[:500] = b[2000:]
    (alt / "c" / "1" / "0").write_bytes(gzip.compress(pad.tobytes(), 1))
    got = io.read_zarr(alt)
    assert (got[:2000] == 0).all()
    np.testing.assert_array_equal(got[2000:], b[2000:])


def test_model_directory_is_loaded_once_per_process(tmp_path, monkeypatch):
    """load_orcai_model keeps a per-process cache keyed by (directory, file sizes / mtimes, device, arithmetic): table after table
    reuses the bound model, an edited weight file is re-read."""
    from orcai_b200 import model as model_mod
    from orcai_b200.weights import save_npz, synthetic_weights

    P, S = runtime.bundled_parameters()
    d = tmp_path / "orcai-V1"
    d.mkdir()
    (d / "orcai_parameter.json").write_text(json.dumps(P))
    (d / "model_shape.json").write_text(json.dumps(S))
    save_npz(synthetic_weights(P, S, seed=1), d / "orcai-v1.weights.npz")
    made = []

    class FakeModel:
        def __init__(self, orcai_parameter, shape, W, device=None, precision=None):
            made.append((float(W["dense2/bias"][0]), device))

    monkeypatch.setattr(model_mod, "OrcaiModel", FakeModel)
    monkeypatch.delenv("ORCAI_B200_PRECISION", raising=False)
    io._MODEL_CACHE.clear()
    m1, p1, s1 = io.load_orcai_model(d, device=0)
    m2, _, _ = io.load_orcai_model(d, device=0)
    assert m1 is m2 and len(made) == 1 and p1 == P and s1 == S
    io.load_orcai_model(d, device=1)                                   # another device: its own model
    monkeypatch.setenv("ORCAI_B200_PRECISION", "reference")            # another arithmetic: its own model
    io.load_orcai_model(d, device=0)
    assert len(made) == 3
    W2 = synthetic_weights(P, S, seed=2)
    save_npz(W2, d / "orcai-v1.weights.npz")                           # edited weights are re-read
    import os as _os
    _os.utime(d / "orcai-v1.weights.npz", ns=(1, 1))
    io.load_orcai_model(d, device=0)
    assert len(made) == 4 and made[-1][0] == float(W2["dense2/bias"][0])
    io._MODEL_CACHE.clear()


def test_table_pipeline_keeps_two_recordings_in_flight_and_reports_failures_per_row(tmp_path, monkeypatch):
    """predict(TABLE.csv) on one device, with a stand-in context that logs the C-ABI calls: recording k+1 is handed over and enqueued
    (swap, predict_begin) BEFORE recording k is collected (predict_end), its upload starts only after recording k-1 - whose device
    buffer it takes - has been collected; a missing file, a recording that is too short and an existing label file are reported
    against their own rows and the rows after them still run (reference loop predict.py:752-755)."""
    import pandas as pd

    from orcai_b200 import _lib, predict as op, wavio
    from orcai_b200._lib import ORCAI_ERR_TOO_SHORT, OrcaiError, SpecStats

    P, S = runtime.bundled_parameters()
    log = []

    class Pool:
        def __init__(self, max_free=0):
            self.out = 0

        def take(self, nbytes):
            self.out += 1
            return np.zeros(nbytes, np.uint8)

        def give(self, a):
            self.out -= 1

    pools = []
    monkeypatch.setattr(_lib, "PinnedPool", lambda max_free=0: pools.append(Pool()) or pools[-1])

    class Ctx:
        device = 0

        class params:
            n_freq = S["input_shape"][1]

        def __init__(self):
            self.next = self.cur = None
            self.flight = []

        def prefetch_pcm(self, pcm):
            log.append(("prefetch", int(pcm[0])))
            self.next = int(pcm[0])

        def swap_pcm(self):
            assert self.next is not None
            log.append(("swap", self.next))
            self.cur, self.next = self.next, None

        def predict_begin(self, n_samples, threshold=0.5, want_agg=True):
            assert len(self.flight) < 2 and not want_agg
            if n_samples < 1000:
                raise OrcaiError(ORCAI_ERR_TOO_SHORT, "recording has 3 frames, shorter than one snippet of 736")
            log.append(("begin", self.cur))
            self.flight.append(self.cur)
            return self.cur

        def predict_end(self, token):
            assert self.flight.pop(0) == token
            log.append(("end", token))
            st = SpecStats()
            st.n_frames = 1000
            return st, None, None, np.array([token % 7], np.int32), np.array([token], np.int64), np.array([token + 3], np.int64)

    class Model:
        ctx = Ctx()

        def bind(self):
            return self.ctx

        def describe(self):
            return "stand-in"

    monkeypatch.setattr(op, "load_orcai_model", lambda model_dir, device=None: (Model(), P, S))
    monkeypatch.delenv("ORCAI_B200_DEVICES", raising=False)
    monkeypatch.delenv("ORCAI_B200_SHARD", raising=False)
    for v in ("RANK", "WORLD_SIZE"):
        monkeypatch.delenv(v, raising=False)
    monkeypatch.setattr(op, "_table_devices", lambda: [0])
    rows = []
    for k in range(7):
        n = 500 if k == 4 else 5000            # row 4 is too short
        pcm = np.full(n, k + 1, np.int16)      # the first sample names the recording in the log
        if k != 2:                             # row 2's file is missing
            wavio.write_wav_pcm16(tmp_path / f"w{k}.wav", pcm)
        rows.append({"recording": f"r{k}", "channel": 1, "base_dir_recording": str(tmp_path), "rel_recording_path": f"w{k}.wav"})
    csv = tmp_path / "t.csv"
    pd.DataFrame(rows).to_csv(csv, index=False)
    out = tmp_path / "out"
    out.mkdir()
    (out / "r5_orcai-V1_predicted.txt").write_text("old")    # row 5: label file exists, no overwrite
    errors = []

    class Msgr(op.Messenger):
        def error(self, text, *a, **k):
            errors.append(str(text))

    op.predict(csv, model_dir=tmp_path / "orcai-V1", output_path=str(out), msgr=Msgr(verbosity=0))
    ev = [e for e in log if e[0] in ("begin", "end")]
    # rows 0, 1 overlap; row 2 (missing) only collects row 1; row 3 runs; row 4 (too short) collects it; row 5 (file exists); row 6
    assert ev == [("begin", 1), ("begin", 2), ("end", 1), ("end", 2), ("begin", 4), ("end", 4), ("begin", 7), ("end", 7)]
    # the upload of the row after next starts only when the buffer it takes has been released by predict_end
    assert log.index(("prefetch", 4)) > log.index(("end", 1)) and log.index(("prefetch", 7)) > log.index(("end", 4))
    assert sorted(f.name for f in out.iterdir()) == [f"r{k}_orcai-V1_predicted.txt" for k in (0, 1, 3, 5, 6)]
    assert (out / "r5_orcai-V1_predicted.txt").read_text() == "old"
    assert (out / "r3_orcai-V1_predicted.txt").read_text().splitlines()[1].split("\t")[2] == P["calls"][4 % 7] + "*"
    assert len(errors) == 3 and "r2" in errors[0] and "r4" in errors[1] and "shorter than one snippet" in errors[1] and "r5" in errors[2] and "already exists" in errors[2]
    assert pools and pools[0].out == 0           # every page-locked buffer went back to the pool


def test_stft_kernel_dataflow_replayed_on_the_cpu(tmp_path):
    """tests/host_emul/stft_emul.cpp replays K1's per-thread building blocks (stft_core.cuh: 8 threads per frame,
    stft_core16.cuh: 16 threads per frame) thread by thread on the CPU and checks every bin against a naive float64 DFT."""
    import shutil
    import subprocess

    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    root = Path(__file__).resolve().parent.parent
    exe = tmp_path / "stft_emul"
    subprocess.run([gxx, "-O2", "-std=c++17", "-I", str(root / "orcai_b200" / "csrc"), str(root / "tests" / "host_emul" / "stft_emul.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
