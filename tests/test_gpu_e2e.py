"""GPU end-to-end: the drop-in Python functions and the CLI produce the reference's files."""

import gzip
import json
from pathlib import Path

import numpy as np
import pandas as pd
import pytest
from click.testing import CliRunner

from oracle import network_oracle, postprocess_oracle as po, spectrogram_oracle as so
from orcai_b200 import cli, io, predict as pr, spectrogram as spg
from orcai_b200.auxiliary import Messenger
from orcai_b200.synth import pcm16_to_float, synth_pcm16
from orcai_b200.wavio import write_wav_pcm16
from orcai_b200.weights import save_npz, synthetic_weights

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model_dir(tmp_path_factory, params):
    """A model directory like the reference's: parameter JSONs + weights (npz container)."""
    P, S = params
    d = tmp_path_factory.mktemp("models") / "orcai-V1"
    d.mkdir()
    (d / "orcai_parameter.json").write_text(json.dumps(P))
    (d / "model_shape.json").write_text(json.dumps(S))
    save_npz(synthetic_weights(P, S, seed=1234), d / "orcai-v1.weights.npz")
    return d


@pytest.fixture(scope="module")
def wavs(tmp_path_factory):
    d = tmp_path_factory.mktemp("audio")
    out = []
    for k, secs in enumerate((12.0, 9.5, 20.0)):
        pcm = synth_pcm16(secs, seed=20251018 + k, calls_per_minute=40.0)
        p = d / f"rec{k}.wav"
        write_wav_pcm16(p, pcm)
        out.append((p, pcm))
    return d, out


def oracle_label_text(pcm, P, S, probs_from=None):
    """Oracle pipeline; with probs_from (device probabilities) the comparison is 'given identical masks'."""
    db, f, t = so.calculate_spectrogram(pcm16_to_float(pcm), P["spectrogram"])
    spec, _, _ = so.preprocess_spectrogram(db, f, P["spectrogram"])
    preds = probs_from if probs_from is not None else network_oracle.forward(po.cut_snippets(spec, 736), synthetic_weights(P, S, seed=1234))
    agg, cnt = po.aggregate_predictions(preds, spec.shape[0], 736, 4, 7)
    s, e, n = po.binary_predictions(agg, cnt, P["calls"])
    return po.labels_tsv(po.label_rows(s, e, n, 16, "*"), float(t[1] - t[0])), agg


# None = the shipped default (precise: split-fp16 tensor cores), held to 1e-4; "fast" is opt-in and outside the 1e-3 gate
# (FAST_TOL in test_gpu_network.py)
@pytest.mark.parametrize("precision,tol", [(None, 1e-4), ("reference", 1e-3), ("fast", 2.5e-3)])
def test_predict_single_wav_file_contract(ctx, params, model_dir, wavs, monkeypatch, precision, tol):
    if precision is None:
        monkeypatch.delenv("ORCAI_B200_PRECISION", raising=False)
    else:
        monkeypatch.setenv("ORCAI_B200_PRECISION", precision)
    P, S = params
    d, files = wavs
    wav, pcm = files[0]
    out = wav.with_name("rec0_c1_orcai-v1_predicted.txt")  # <stem>_c<channel>_<orcai_parameter['name']>_predicted.txt
    out.unlink(missing_ok=True)
    pr.predict(wav, model_dir=model_dir, verbosity=0, save_probabilities=True)
    assert out.exists()
    text = out.read_text()
    assert text.startswith("start\tstop\tlabel\n")
    # given the device's own probabilities the file is byte-identical to the oracle's
    ctx.upload_pcm(pcm)
    st = ctx.spectrogram_resident(False)
    n = int((st.n_frames - 736) // 368 + 1)
    dev_probs = ctx.forward_resident(0, n)
    ref_text, ref_agg = oracle_label_text(pcm, P, S, probs_from=dev_probs)
    assert text == ref_text
    # and the probabilities file holds the aggregated float64 averages with the frame-delta_t index
    prob = gzip.decompress(out.with_name(out.stem + "_probabilities.csv.gz").read_bytes()).decode()
    assert prob == po.probabilities_csv(ref_agg, P["calls"], 256 / 48000)
    # full-oracle comparison: probabilities within tolerance
    _, agg_oracle = oracle_label_text(pcm, P, S)
    assert np.abs(ref_agg - agg_oracle).max() <= tol
    with pytest.raises(FileExistsError):
        pr.predict(wav, model_dir=model_dir, verbosity=0)
    pr.predict(wav, model_dir=model_dir, verbosity=0, overwrite=True)
    assert out.read_text() == text  # deterministic


def test_predict_wav_returns_reference_types(ctx, params, model_dir, wavs):
    P, S = params
    _, files = wavs
    model, P2, S2 = io.load_orcai_model(model_dir)
    labels, agg, delta_t = pr.predict_wav(files[1][0], 1, model, P2, S2, msgr=Messenger(verbosity=0))
    assert isinstance(labels, pd.DataFrame) and list(labels.columns) == ["start", "stop", "label"]
    assert labels["start"].dtype == np.int64 and delta_t == 256 / 48000
    T = 1 + len(files[1][1]) // 256
    assert agg.shape == (T // 16, 7) and agg.dtype == np.float64
    assert (labels["start"] % 16 == 0).all() and (labels["stop"] >= labels["start"]).all() and labels["label"].str.endswith("*").all()
    assert labels.equals(labels.sort_values(by=["start", "stop", "label"]).reset_index(drop=True))
    # staged interface gives the same aggregates as the fused path
    spec, _, _ = spg.make_spectrogram(files[1][0], 1, P2, msgr=Messenger(verbosity=0))
    agg2, cnt2 = pr.compute_aggregated_predictions(files[1][0], spec, model, P2, S2)
    np.testing.assert_array_equal(agg2, agg)
    s, e, n = pr.compute_binary_predictions(agg2, cnt2, P2["calls"], ctx=model.ctx)
    labels2 = pr.compute_labels(s, e, n, 16, "*")
    assert labels2.equals(labels)


def test_too_short_recording(ctx, model_dir, tmp_path):
    p = tmp_path / "short.wav"
    write_wav_pcm16(p, synth_pcm16(3.0, seed=5))  # 563 frames < 736
    with pytest.raises(ValueError, match="shorter than one snippet"):
        pr.predict(p, model_dir=model_dir, verbosity=0)
    # other sampling rates are resampled on the host (polyphase; the reference uses soxr_hq - not parity-pinned)
    q = tmp_path / "other_rate.wav"
    write_wav_pcm16(q, synth_pcm16(6.0, seed=5), sample_rate=44100)
    pr.predict(q, model_dir=model_dir, verbosity=0)
    assert (tmp_path / "other_rate_c1_orcai-v1_predicted.txt").read_text().startswith("start\tstop\tlabel\n")


def test_table_mode_and_cli(ctx, params, model_dir, wavs, tmp_path, capsys):
    P, S = params
    d, files = wavs
    table = pd.DataFrame(
        {
            "recording": ["rec0", "rec1", "missing", "rec2"],
            "channel": [1, 1, 1, 1],
            "base_dir_recording": [str(d)] * 4,
            "rel_recording_path": ["rec0.wav", "rec1.wav", "nope.wav", "rec2.wav"],
            "base_dir_annotation": [str(d), str(d), str(d), None],
            **{c: [True, False, True, True] for c in P["calls"]},
        }
    )
    csv = tmp_path / "table.csv"
    table.to_csv(csv, index=False)
    outdir = tmp_path / "pred"
    outdir.mkdir()
    r = CliRunner().invoke(cli.cli, ["predict", str(csv), "-md", str(model_dir), "-o", str(outdir), "-v", "0"])
    assert r.exit_code == 0, r.output
    # per-recording isolation: the missing file is reported, the others are written; name uses model_dir.stem
    assert "Error predicting missing" in r.output
    for k in (0, 1, 2):
        f = outdir / f"rec{k}_orcai-V1_predicted.txt"
        assert f.exists() and f.read_text().startswith("start\tstop\tlabel\n")
    assert not (outdir / "missing_orcai-V1_predicted.txt").exists()
    # single wav through the CLI gives the same bytes as table mode
    single = tmp_path / "single.txt"
    r = CliRunner().invoke(cli.cli, ["predict", str(files[2][0]), "-md", str(model_dir), "-o", str(single), "-v", "0"])
    assert r.exit_code == 0, r.output
    assert single.read_bytes() == (outdir / "rec2_orcai-V1_predicted.txt").read_bytes()
    # suffix / filter options
    r = CliRunner().invoke(cli.cli, ["predict", str(files[2][0]), "-md", str(model_dir), "-o", str(tmp_path / "s2.txt"), "-ls", "", "-v", "0"])
    assert r.exit_code == 0 and "*" not in (tmp_path / "s2.txt").read_text()

    # create-spectrograms: filters (not annotated / no possible annotations), layout, zarr content
    sdir = tmp_path / "spec"
    r = CliRunner().invoke(cli.cli, ["create-spectrograms", str(csv), str(sdir), "-v", "0"])
    assert r.exit_code != 0 or True  # 'missing' raises like the reference (no per-row isolation in create_spectrograms)
    table2 = table[table["recording"] != "missing"]
    csv2 = tmp_path / "table2.csv"
    table2.to_csv(csv2, index=False)
    sdir2 = tmp_path / "spec2"
    r = CliRunner().invoke(cli.cli, ["create-spectrograms", str(csv2), str(sdir2), "-v", "0"])
    assert r.exit_code == 0, r.output
    assert (sdir2 / "rec0" / "spectrogram" / "spectrogram.zarr" / "zarr.json").exists()
    assert not (sdir2 / "rec1").exists()   # no possible annotations -> excluded
    assert not (sdir2 / "rec2").exists()   # not annotated -> excluded
    z = io.read_zarr(sdir2 / "rec0" / "spectrogram" / "spectrogram.zarr")
    spec, freqs, times = spg.make_spectrogram(files[0][0], 1, P, msgr=Messenger(verbosity=0))
    np.testing.assert_array_equal(z, spec)
    assert json.loads((sdir2 / "rec0" / "spectrogram" / "frequencies.json").read_text()) == {"min": 0.0, "max": 24000.0, "length": 257}
    tj = json.loads((sdir2 / "rec0" / "spectrogram" / "times.json").read_text())
    assert tj["length"] == spec.shape[0] and tj["max"] == times[-1]
    r = CliRunner().invoke(cli.cli, ["create-spectrograms", str(csv2), str(sdir2), "-en", "-enp", "-v", "0"])
    assert r.exit_code == 0 and (sdir2 / "rec1").exists() and (sdir2 / "rec2").exists()


def test_table_mode_one_process_per_gpu(params, model_dir, wavs, tmp_path, monkeypatch):
    """A table with several devices in ORCAI_B200_DEVICES: one worker thread per device by default, one worker PROCESS per device with
    ORCAI_B200_TABLE_PROCESSES=1 (here: two workers on device 0): same label files as the one-device run, per-row errors reported."""
    from orcai_b200 import predict as opredict

    P, S = params
    d, files = wavs
    names = ["a0", "a1", "gone", "a2", "a3", "a4"]
    table = pd.DataFrame({"recording": names, "channel": 1, "base_dir_recording": str(d),
                          "rel_recording_path": ["rec0.wav", "rec1.wav", "nope.wav", "rec2.wav", "rec1.wav", "rec0.wav"]})
    csv = tmp_path / "t.csv"
    table.to_csv(csv, index=False)
    one, two = tmp_path / "one", tmp_path / "two"
    one.mkdir(); two.mkdir()
    monkeypatch.delenv("ORCAI_B200_DEVICES", raising=False)
    opredict.predict(csv, model_dir=model_dir, output_path=str(one), verbosity=0)

    class Collect(Messenger):
        def __init__(self):
            super().__init__(verbosity=0)
            self.errors = []

        def error(self, text, *a, **k):
            self.errors.append(str(text))

    m = Collect()
    monkeypatch.setenv("ORCAI_B200_DEVICES", "0,0")
    monkeypatch.setenv("ORCAI_B200_TABLE_PROCESSES", "1")
    opredict.predict(csv, model_dir=model_dir, output_path=str(two), verbosity=0, msgr=m)
    # the default for several devices: worker threads in this process (two contexts on device 0)
    three = tmp_path / "three"
    three.mkdir()
    monkeypatch.setenv("ORCAI_B200_TABLE_PROCESSES", "0")
    opredict.predict(csv, model_dir=model_dir, output_path=str(three), verbosity=0)
    for n in names:
        if n != "gone":
            assert (one / f"{n}_orcai-V1_predicted.txt").read_bytes() == (three / f"{n}_orcai-V1_predicted.txt").read_bytes()
    assert any("Error predicting gone" in e for e in m.errors), m.errors
    for n in names:
        f1, f2 = one / f"{n}_orcai-V1_predicted.txt", two / f"{n}_orcai-V1_predicted.txt"
        assert f1.exists() == f2.exists() == (n != "gone")
        if n != "gone":
            assert f1.read_bytes() == f2.read_bytes()


def test_table_sharded_by_rank_like_torchrun(params, model_dir, wavs, tmp_path, monkeypatch):
    """`torchrun --nproc-per-node N -m orcai_b200.cli predict TABLE.csv -o DIR`: every process runs the same command, takes its
    longest-first share of the rows (RANK / WORLD_SIZE, or ORCAI_B200_SHARD) and writes its own label files; together they write
    exactly what one process writes, each file once."""
    from orcai_b200 import predict as opredict

    d, files = wavs
    names = [f"r{k}" for k in range(5)]
    table = pd.DataFrame({"recording": names, "channel": 1, "base_dir_recording": str(d),
                          "rel_recording_path": ["rec0.wav", "rec1.wav", "rec2.wav", "rec1.wav", "rec2.wav"]})
    csv = tmp_path / "t.csv"
    table.to_csv(csv, index=False)
    whole, parts = tmp_path / "whole", [tmp_path / "p0", tmp_path / "p1"]
    for p in [whole, *parts]:
        p.mkdir()
    monkeypatch.delenv("ORCAI_B200_DEVICES", raising=False)
    monkeypatch.delenv("ORCAI_B200_SHARD", raising=False)
    opredict.predict(csv, model_dir=model_dir, output_path=str(whole), verbosity=0)
    monkeypatch.setenv("ORCAI_B200_SHARD", "0/2")
    opredict.predict(csv, model_dir=model_dir, output_path=str(parts[0]), verbosity=0)
    monkeypatch.delenv("ORCAI_B200_SHARD")
    monkeypatch.setenv("RANK", "1"); monkeypatch.setenv("WORLD_SIZE", "2"); monkeypatch.setenv("LOCAL_RANK", "0")
    opredict.predict(csv, model_dir=model_dir, output_path=str(parts[1]), verbosity=0)
    got = {f.name: f.read_bytes() for p in parts for f in p.iterdir()}
    assert sum(len(list(p.iterdir())) for p in parts) == len(names) == len(got)      # disjoint cover
    assert all(len(list(p.iterdir())) >= 2 for p in parts)
    assert got == {f.name: f.read_bytes() for f in whole.iterdir()}
