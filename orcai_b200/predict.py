"""Prediction workflow with the reference's function surface (``src/orcAI/predict.py``).

Signatures, output naming, error behaviour and file contents follow predict.py:235-757 of the
reference; the arithmetic runs in liborcai_b200:

* ``predict_wav`` uses the fused device-resident path (spectrogram -> strided snippets -> forward ->
  overlap-average -> threshold -> run lengths); only segments and aggregates come back.
* ``compute_aggregated_predictions`` / ``compute_binary_predictions`` keep the reference's staged
  interfaces (host spectrogram in, numpy out) on top of the same kernels.
* The label file is written without pandas' ``.loc`` float-into-int assignment, which raises on
  pandas >= 3 (predict.py:362-363); the bytes are what pandas 2.2.3 produced.
"""

from __future__ import annotations

import gzip
import os
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from importlib.resources import files
from pathlib import Path

import numpy as np
import pandas as pd
from tqdm import tqdm

from orcai_b200._lib import ORCAI_ERR_TOO_SHORT, OrcaiError
from orcai_b200.auxiliary import Messenger
from orcai_b200.io import load_orcai_model, read_json
from orcai_b200.runtime import get_context
from orcai_b200.spectrogram import frames_to_time, load_recording, make_spectrogram


# ---------------------------------------------------------------------------------------------
# duration filter (predict.py:14-159)
# ---------------------------------------------------------------------------------------------
def _check_duration(calls, call_duration_limits: dict, delta_t: float, label_suffix: str = "*") -> str:
    label = calls["label"].replace(f"{label_suffix}", "")
    if label in call_duration_limits:
        mn, mx = call_duration_limits[label]
    elif "default" in call_duration_limits:
        mn, mx = call_duration_limits["default"]
    else:
        mn, mx = 0, np.inf
    mn = 0 if mn is None else mn
    mx = np.inf if mx is None else mx
    if calls["duration"] * delta_t < mn:
        return "too short"
    if calls["duration"] * delta_t > mx:
        return "too long"
    return "keep"


def filter_predictions(
    predicted_labels: pd.DataFrame,
    delta_t: float,
    call_duration_limits: (Path | str) | dict = files("orcai_b200.defaults").joinpath("default_call_duration_limits.json"),
    label_suffix: str = "*",
    verbosity: int = 2,
    msgr: Messenger | None = None,
) -> pd.DataFrame:
    """Drop predicted calls whose duration is outside the per-label limits."""
    if msgr is None:
        msgr = Messenger(verbosity=verbosity, title="Filtering predictions")
    msgr.part("Filtering predictions")
    predicted_labels = predicted_labels.copy()
    predicted_labels["duration"] = predicted_labels["stop"] - predicted_labels["start"]
    if not isinstance(call_duration_limits, dict):
        call_duration_limits = read_json(call_duration_limits)
    msgr.debug("Call duration limits:")
    msgr.debug(call_duration_limits)
    msgr.part("Filtering calls based on duration")
    if len(predicted_labels):
        verdict = predicted_labels.apply(lambda x: _check_duration(x, call_duration_limits, delta_t, label_suffix), axis=1)
    else:
        verdict = pd.Series([], dtype=object)
    predicted_labels["duration_ok"] = verdict
    n_long = int((verdict == "too long").sum())
    n_short = int((verdict == "too short").sum())
    msgr.info(f"Discarding {n_long + n_short} calls based on duration (too short: {n_short}, too long: {n_long})")
    out = predicted_labels[predicted_labels["duration_ok"] == "keep"]
    msgr.success("Filtering predictions finished.")
    return out


def filter_predictions_file(
    predicted_labels: Path | str,
    output_file: Path | str = "default",
    overwrite: bool = False,
    call_duration_limits: (Path | str) | dict = files("orcai_b200.defaults").joinpath("default_call_duration_limits.json"),
    label_suffix: str = "*",
    verbosity: int = 2,
    msgr: Messenger | None = None,
):
    """Filter a saved label file by duration -> ``<stem>_filtered.txt`` (predict.py:162-232)."""
    if msgr is None:
        msgr = Messenger(verbosity=verbosity, title="Filtering predictions")
    if output_file == "default":
        output_file = Path(predicted_labels).with_name(Path(predicted_labels).stem + "_filtered.txt")
    else:
        output_file = Path(output_file)
    msgr.info(f"Output file: {output_file}")
    if output_file.exists() and not overwrite:
        raise FileExistsError(f"Annotation file already exists: {output_file}")
    table = pd.read_csv(predicted_labels, sep="\t", encoding="utf-8")
    kept = filter_predictions(table, delta_t=1, call_duration_limits=call_duration_limits, label_suffix=label_suffix, verbosity=verbosity, msgr=msgr)
    save_predictions(kept, output_file, delta_t=1, msgr=msgr)


# ---------------------------------------------------------------------------------------------
# staged interfaces (predict.py:235-340)
# ---------------------------------------------------------------------------------------------
def _context_of(model, orcai_parameter: dict, shape: dict):
    """The model's context with THIS model's weights and arithmetic bound (contexts are shared per device, model.bind())."""
    if hasattr(model, "bind"):
        return model.bind()
    return getattr(model, "ctx", None) or get_context(orcai_parameter, shape)


def compute_aggregated_predictions(
    recording_path: Path,
    spectrogram: np.ndarray,
    model,
    orcai_parameter: dict,
    shape: dict,
    msgr: Messenger = Messenger(verbosity=0),
    progressbar: tqdm = None,
) -> tuple[np.ndarray, np.ndarray]:
    """Snippets -> model.predict -> float64 overlap-average; returns (aggregated (T//16, L), overlap_count)."""
    snippet_length = shape["input_shape"][0]
    shift = snippet_length // 2
    num_snippets = (spectrogram.shape[0] - snippet_length) // shift + 1
    msgr.info(f"slicing into {num_snippets} snippets for prediction")
    if num_snippets <= 0:
        raise ValueError(
            f"Spectrogram of {Path(recording_path).stem} has {spectrogram.shape[0]} frames, shorter than one snippet ({snippet_length})"
        )
    spec = np.ascontiguousarray(spectrogram, dtype=np.float32)
    # strided window view; model.predict stages it in bounded batches, nothing is materialised on the host
    windows = np.lib.stride_tricks.sliding_window_view(spec, (snippet_length, spec.shape[1]))[::shift, 0][:num_snippets]
    msgr.info("Prediction of snippets")
    predictions = model.predict(windows[..., np.newaxis], verbose=0 if msgr.verbosity < 2 else 1)
    msgr.info("Aggregating predictions")
    if progressbar:
        progressbar.set_description(f"{Path(recording_path).stem} - Aggregating predictions")
        progressbar.refresh()
    ctx = _context_of(model, orcai_parameter, shape)
    agg, cnt, _, _, _ = ctx.postprocess(predictions, spectrogram.shape[0], threshold=0.5, want_agg=True)
    return agg, cnt


def compute_binary_predictions(
    aggregated_predictions: np.ndarray,
    overlap_count: np.ndarray,
    calls: list[str],
    threshold: float = 0.5,
    ctx=None,
) -> tuple[list[int], list[int], list[str]]:
    """Threshold at threshold/max(overlap) (strict >) and extract per-label runs, label-major order."""
    if ctx is None:
        from orcai_b200.runtime import default_context

        ctx = default_context()
    lab, sta, sto = ctx.threshold_segments(aggregated_predictions, overlap_count, threshold)
    return [int(v) for v in sta], [int(v) for v in sto], [calls[int(i)] for i in lab]


def compute_labels(
    row_starts: list[int],
    row_stops: list[int],
    label_names: list[str],
    time_steps_per_output_step: int,
    label_suffix: str | None,
) -> pd.DataFrame:
    if (label_suffix is not None) & (label_suffix != ""):
        label_names = [label + label_suffix for label in label_names]
    start = np.asarray(row_starts) * time_steps_per_output_step
    stop = np.asarray(row_stops) * time_steps_per_output_step
    if len(label_names) and start.dtype.kind in "iu":
        # sort_values(by=[start, stop, label]) of the reference (predict.py:329-339) as one lexsort: labels compare like their
        # rank among the distinct names (a table run builds thousands of these tables under the interpreter lock)
        names = sorted(set(label_names))
        rank = {n: r for r, n in enumerate(names)}
        lab = np.fromiter((rank[n] for n in label_names), dtype=np.int64, count=len(label_names))
        order = np.lexsort((lab, stop, start))
        name_arr = np.empty(len(names), dtype=object)
        name_arr[:] = names
        return pd.DataFrame({"start": start[order], "stop": stop[order], "label": name_arr[lab[order]]})
    return (
        pd.DataFrame({"start": start, "stop": stop, "label": label_names})
        .sort_values(by=["start", "stop", "label"])
        .reset_index(drop=True)
    )


# ---------------------------------------------------------------------------------------------
# predict_wav (predict.py:367-471)
# ---------------------------------------------------------------------------------------------
def _time_split_contexts(model, ctx):
    """Contexts (one per device of ORCAI_B200_DEVICES) that annotate ONE recording together, or None for the one-device path.

    Replicas of the model on the other devices are created once and kept on the model object.
    """
    devs = _visible_devices()
    if len(devs) < 2 or not hasattr(model, "weights"):
        return None
    from orcai_b200.model import OrcaiModel

    cache = model.__dict__.setdefault("_device_replicas", {ctx.device: model})
    for d in devs:
        if d not in cache:
            cache[d] = OrcaiModel(model.orcai_parameter, model.shape, model.weights, device=d, precision=model.precision)
    ordered = [ctx.device] + [d for d in devs if d != ctx.device]
    return [cache[d].ctx for d in ordered]


def _device_predict(recording_path, channel, model, orcai_parameter, shape, msgr, progressbar, _resident_samples=None, _time_split=False):
    """The device part of predict_wav: samples -> (stats, aggregated probabilities, label indices, start steps, stop steps, delta_t)."""
    recording_path = Path(recording_path)
    if progressbar:
        progressbar.set_description(f"{recording_path.stem}: Generating spectrogram")
        progressbar.refresh()
    sp = orcai_parameter["spectrogram"]
    msgr.part("Calculating power spectrogram by stft")
    msgr.info(f"Loading & resampling (to {sp['sampling_rate'] / 1000:.2f} kHz) wav file: {recording_path.stem}")
    samples = _resident_samples if _resident_samples is not None else load_recording(recording_path, channel, sp, msgr)
    ctx = _context_of(model, orcai_parameter, shape)
    if ctx.params.n_freq != shape["input_shape"][1]:
        raise ValueError(f"Spectrogram shape ({ctx.params.n_freq}) for {recording_path.stem} not equal to input shape ({shape['input_shape'][1]})")
    times01 = frames_to_time(2, sp)
    delta_t = times01[1] - times01[0]

    msgr.part(f"Prediction of annotations for wav_file: {recording_path.stem}")
    if progressbar:
        progressbar.set_description(f"{recording_path.stem} - Predicting annotations")
        progressbar.refresh()
    try:
        replicas = _time_split_contexts(model, ctx) if (_time_split and _resident_samples is None) else None
        if replicas:
            # ONE recording on several GPUs: time chunks with halos, one host-side exchange of statistics (timesplit.py)
            from orcai_b200.timesplit import predict_pcm_timesplit

            msgr.info(f"splitting the recording by time across {len(replicas)} devices")
            stats, agg, _cnt, lab, sta, sto = predict_pcm_timesplit(replicas, samples, threshold=0.5, want_agg=True)
        else:
            stats, agg, _cnt, lab, sta, sto = ctx.predict_pcm(samples, threshold=0.5, want_agg=True, resident=_resident_samples is not None)
    except OrcaiError as e:
        if e.code == ORCAI_ERR_TOO_SHORT:
            raise ValueError(f"{recording_path.stem}: {e.message}") from e
        raise
    msgr.info(f"Duration of wav file: {(int(stats.n_frames) - 1) * delta_t:.2f} seconds")
    return stats, agg, lab, sta, sto, delta_t


def _labels_of(lab, sta, sto, orcai_parameter: dict, label_suffix: str, msgr) -> pd.DataFrame:
    """Segments of the device scan -> the reference's label table (predict.py:320-340): what ``compute_labels`` returns for
    the same segments, built from the label INDICES with array operations only (no per-segment Python work)."""
    msgr.info("converting binary predictions into start and stop frames")
    step = 2 ** len(orcai_parameter["model"]["filters"])
    lab = np.asarray(lab, dtype=np.int64)
    if lab.size == 0:
        return compute_labels([], [], [], step, label_suffix)
    suffix = label_suffix if (label_suffix is not None) and (label_suffix != "") else ""
    names = np.empty(len(orcai_parameter["calls"]), dtype=object)
    names[:] = [c + suffix for c in orcai_parameter["calls"]]
    rank = np.empty(len(names), dtype=np.int64)
    rank[sorted(range(len(names)), key=lambda j: names[j])] = np.arange(len(names))   # labels compare like their names
    start = np.asarray(sta, dtype=np.int64) * step
    stop = np.asarray(sto, dtype=np.int64) * step
    order = np.lexsort((rank[lab], stop, start))
    predicted_labels = pd.DataFrame({"start": start[order], "stop": stop[order], "label": names[lab[order]]})
    msgr.info(f"found {len(predicted_labels)} acoustic signals")
    return predicted_labels


def predict_wav(
    recording_path: Path | str,
    channel: int,
    model,
    orcai_parameter: dict,
    shape: dict,
    label_suffix: str = "*",
    msgr: Messenger = Messenger(verbosity=0),
    progressbar: tqdm = None,
    _resident_samples=None,
    _time_split: bool = False,
):
    """Predicts calls in a single wav file -> (predicted_labels DataFrame, aggregated_predictions, delta_t).

    ``_resident_samples`` (table mode): the recording has already been read and uploaded by the prefetcher.
    ``_time_split`` (single-file mode of ``predict``): with several devices in ORCAI_B200_DEVICES the recording is cut into
    time chunks, one per device (``orcai_b200/timesplit.py``); the result is bit-identical to the one-device path.
    """
    _stats, agg, lab, sta, sto, delta_t = _device_predict(recording_path, channel, model, orcai_parameter, shape, msgr, progressbar,
                                                           _resident_samples, _time_split)
    predicted_labels = _labels_of(lab, sta, sto, orcai_parameter, label_suffix, msgr)
    msgr.success("Prediction finished.")
    return predicted_labels, agg, delta_t


# ---------------------------------------------------------------------------------------------
# writers (predict.py:343-364, 474-531)
# ---------------------------------------------------------------------------------------------
TABLE_STATS: dict = {}   # host-side accounting of this process' last table run (bench.py reports it): where its main thread waited
_SECONDS_TEXT: dict[float, dict[int, str]] = {}   # delta_t -> {frame index: text of its time in seconds}
_SECONDS_LOCK = threading.Lock()                  # writer threads of several GPUs share the cache


def _seconds_column(values, delta_t: float) -> list[str]:
    """Text of one time column as pandas 2.2.3 wrote it after ``df.loc[:, c] = df.loc[:, c] * delta_t``.

    An int64 column stays int64 when every product is integer-valued (lossless in-place set), otherwise it
    becomes float64, rounded half-to-even to 4 decimals and printed with the shortest round-trip repr.
    """
    v = np.asarray(values)
    prod = v * np.float64(delta_t)
    if prod.size and v.dtype.kind in "iu" and np.all(prod == np.trunc(prod)):
        return [str(int(x)) for x in prod]
    if v.dtype.kind in "iu" and v.size > 64:
        # frame indices repeat from recording to recording: format each distinct (delta_t, frame) once per process
        keys = v.tolist()
        with _SECONDS_LOCK:
            cache = _SECONDS_TEXT.setdefault(float(delta_t), {})
            if len(cache) > 4_000_000:
                cache = _SECONDS_TEXT[float(delta_t)] = {}      # readers holding the old dict keep a consistent view
            missing = [k for k in set(keys) if k not in cache]
            if missing:
                m = np.asarray(missing, dtype=v.dtype)
                for k, x in zip(missing, np.round((m * np.float64(delta_t)).astype(np.float64), 4)):
                    cache[k] = repr(float(x))
            return [cache[k] for k in keys]
    return [repr(float(x)) for x in np.round(prod.astype(np.float64), 4)]


def labels_to_tsv(predicted_labels: pd.DataFrame, delta_t: float) -> str:
    start = _seconds_column(predicted_labels["start"].to_numpy(), delta_t)
    stop = _seconds_column(predicted_labels["stop"].to_numpy(), delta_t)
    lines = ["start\tstop\tlabel"]
    lines += [f"{a}\t{b}\t{c}" for a, b, c in zip(start, stop, predicted_labels["label"].tolist())]
    return "\n".join(lines) + "\n"


def save_predictions(predicted_labels: pd.DataFrame, output_path: Path | str, delta_t: float, msgr: Messenger = Messenger(verbosity=0)) -> None:
    """Audacity-compatible label file: header ``start\\tstop\\tlabel``, times in seconds rounded to 4 decimals."""
    with open(output_path, "w", encoding="utf-8", newline="") as f:
        f.write(labels_to_tsv(predicted_labels, delta_t))
    msgr.info(f"Predictions saved to {output_path}")


def probabilities_to_csv(aggregated_predictions: np.ndarray, calls: list[str], delta_t: float) -> str:
    idx = np.float64(delta_t) * np.arange(len(aggregated_predictions))  # frame delta_t, like the reference (quirk)
    lines = ["time," + ",".join(calls)]
    for t, row in zip(idx, np.asarray(aggregated_predictions, dtype=np.float64)):
        lines.append(repr(float(t)) + "," + ",".join(repr(float(v)) for v in row))
    return "\n".join(lines) + "\n"


def save_prediction_probabilities(
    aggregated_predictions: np.ndarray,
    orcai_parameter: dict,
    delta_t: float,
    output_path: Path | str,
    msgr: Messenger = Messenger(verbosity=0),
) -> None:
    output_path = Path(output_path)
    predictions_path = output_path.with_name(f"{output_path.stem}_probabilities.csv.gz")
    with gzip.open(predictions_path, "wt", encoding="utf-8", newline="") as f:
        f.write(probabilities_to_csv(aggregated_predictions, orcai_parameter["calls"], delta_t))
    msgr.info(f"Prediction probabilities saved to {predictions_path}")


# ---------------------------------------------------------------------------------------------
# drivers (predict.py:534-757)
# ---------------------------------------------------------------------------------------------
def _resolved_output_path(recording_path: Path, channel: int, orcai_parameter: dict, output_path, overwrite: bool, msgr):
    """Where the label file goes (None: nowhere); an existing file is an error unless ``overwrite`` (predict.py:569-583)."""
    if output_path is None:
        return None
    if output_path == "default":
        output_path = recording_path.with_name(f"{recording_path.stem}_c{channel}_{orcai_parameter['name']}_predicted.txt")
    else:
        output_path = Path(output_path)
    msgr.info(f"Output file: {output_path}")
    if output_path.exists():
        if overwrite:
            msgr.warning(f"Output file {output_path} already exists. Overwriting.")
        else:
            raise FileExistsError(f"Annotation file already exists: {output_path}")
    return output_path


def _host_tail(aggregated_predictions, lab, sta, sto, delta_t, orcai_parameter, output_path, save_probabilities, call_duration_limits,
               label_suffix, msgr):
    """Segments of the device scan -> label table -> duration filter -> files (predict.py:585-611)."""
    predicted_labels = _labels_of(lab, sta, sto, orcai_parameter, label_suffix, msgr)
    msgr.success("Prediction finished.")
    if call_duration_limits is not None:
        predicted_labels = filter_predictions(predicted_labels, delta_t=delta_t, call_duration_limits=call_duration_limits, label_suffix=label_suffix, msgr=msgr)
    save_predictions(predicted_labels=predicted_labels, output_path=output_path, delta_t=delta_t, msgr=msgr)
    if save_probabilities:
        save_prediction_probabilities(aggregated_predictions, orcai_parameter, delta_t, output_path, msgr=msgr)


def _predict_and_save(
    recording_path: Path | str,
    channel: int,
    model,
    orcai_parameter: dict,
    shape: dict,
    output_path: Path | str = "default",
    overwrite: bool = False,
    save_probabilities: bool = False,
    call_duration_limits: (Path | str) | dict = None,
    label_suffix: str = "*",
    msgr: Messenger = Messenger(verbosity=0),
    progressbar: tqdm = None,
    _time_split: bool = False,
):
    recording_path = Path(recording_path)
    output_path = _resolved_output_path(recording_path, channel, orcai_parameter, output_path, overwrite, msgr)
    _stats, aggregated_predictions, lab, sta, sto, delta_t = _device_predict(
        recording_path, channel, model, orcai_parameter, shape, msgr, progressbar, None, _time_split
    )
    _host_tail(aggregated_predictions, lab, sta, sto, delta_t, orcai_parameter, output_path, save_probabilities, call_duration_limits,
               label_suffix, msgr)
    return None


def _visible_devices() -> list[int]:
    """Devices named by ORCAI_B200_DEVICES="0,1,.." | "all" (ONE recording is split by time over them); default: the process' one device."""
    spec = os.environ.get("ORCAI_B200_DEVICES", "").strip()
    if not spec:
        return []
    if spec == "all":
        import torch

        return list(range(torch.cuda.device_count()))
    return [int(x) for x in spec.split(",") if x.strip() != ""]


def _table_devices() -> list[int]:
    """Devices a recording TABLE shards over: ORCAI_B200_DEVICES if set, else every GPU the process can see."""
    devs = _visible_devices()
    if devs or os.environ.get("ORCAI_B200_DEVICES", "").strip():
        return devs
    try:
        import torch

        return list(range(torch.cuda.device_count()))
    except Exception:
        return []


def _rank_shard() -> tuple[int, int] | None:
    """(rank, world) when this process is ONE WORKER of a job that runs the same `orcai predict TABLE.csv` in every process:
    `torchrun --nproc-per-node N -m orcai_b200.cli predict TABLE.csv -o DIR` (RANK / WORLD_SIZE / LOCAL_RANK from the launcher)
    or ORCAI_B200_SHARD="rank/world".  Every worker computes the same longest-first plan from the file sizes and annotates
    its own share on its own GPU; nothing is exchanged (shard by recording, no collective)."""
    spec = os.environ.get("ORCAI_B200_SHARD", "").strip()
    if spec:
        r, w = spec.split("/")
        return int(r), int(w)
    w = os.environ.get("WORLD_SIZE", "")
    r = os.environ.get("RANK", "")
    if w.isdigit() and r.isdigit() and int(w) > 1:
        return int(r), int(w)
    return None


def _predict_worker(kwargs: dict) -> None:
    """Entry point of a table worker process (spawned): one GPU, its share of the rows."""
    q = kwargs["_worker"]["queue"]
    try:
        os.environ["ORCAI_B200_DEVICES"] = str(kwargs["_worker"]["device"])
        predict(**kwargs)
    except BaseException as e:  # noqa: BLE001 - everything is reported to the parent
        q.put(("fatal", f"worker on device {kwargs['_worker']['device']}: {type(e).__name__}: {e}"))
    finally:
        q.put(("done", None))


def _predict_table_multiprocess(recording_table, rows, devices, msgr, kwargs) -> None:
    """Shard a recording table by recording over one process per GPU; host-side gather of ticks and error messages only."""
    import multiprocessing as mp
    import queue as queue_mod

    from orcai_b200.sharding import assign_rows, recording_costs

    paths = [Path(recording_table.loc[i, "base_dir_recording"]).joinpath(recording_table.loc[i, "rel_recording_path"]) for i in rows]
    plan = assign_rows(recording_costs(paths), len(devices))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = []
    for d, share in zip(devices, plan):
        kw = dict(kwargs, _worker={"device": int(d), "rows": [rows[j] for j in share], "queue": q})
        p = ctx.Process(target=_predict_worker, args=(kw,), daemon=True)
        p.start()
        procs.append(p)
    progressbar = tqdm(total=len(rows), desc=f"{len(devices)} GPUs", unit="file")
    done = 0
    while done < len(procs):
        try:
            kind, text = q.get(timeout=1.0)
        except queue_mod.Empty:
            if not any(p.is_alive() for p in procs) and q.empty():
                break
            continue
        if kind == "tick":
            progressbar.update(1)
        elif kind == "error":
            msgr.error(text)
        elif kind == "fatal":
            msgr.error(f"Error in table worker: {text}")
        elif kind == "done":
            done += 1
    for p in procs:
        p.join(timeout=30)
    progressbar.close()
    msgr.success("Predictions finished.")


def predict(
    recording_path: str | Path,
    channel: int = 1,
    model_dir: str | Path = files("orcai_b200.models").joinpath("orcai-V1"),
    output_path: str | Path = "default",
    overwrite: bool = False,
    save_probabilities: bool = False,
    base_dir_recording: str | Path | None = None,
    call_duration_limits: str | Path | None = None,
    label_suffix: str = "*",
    verbosity: int = 2,
    msgr: Messenger | None = None,
    _worker: dict | None = None,
) -> None:
    """Predicts calls in a wav file or in every recording of a recording table (.csv).

    A table with several devices in ORCAI_B200_DEVICES is sharded by recording over ONE PROCESS PER GPU: the
    parent plans the shares (longest files first) and spawns the workers, which call this function again with ``_worker``
    = {"device", "rows", "queue"}; the only traffic between them is progress ticks and per-row error messages.
    """
    if msgr is None:
        msgr = Messenger(verbosity=verbosity, title="Predicting calls")
    model_dir = Path(str(model_dir))
    recording_path = Path(recording_path)
    is_table = recording_path.suffix == ".csv"
    shard = _rank_shard() if (is_table and not _worker) else None
    if _worker:
        devices = [_worker["device"]]
    elif shard:
        devices = []                       # one worker of a multi-process job: its own device (LOCAL_RANK / ORCAI_B200_DEVICE)
    elif is_table:
        devices = _table_devices()         # every visible GPU unless ORCAI_B200_DEVICES narrows it
    else:
        devices = _visible_devices()
    # A table on several devices: one worker PROCESS per GPU (own interpreter: the per-recording host work - WAV decode, label
    # table, label file - does not serialise on one interpreter lock; measured round 1: threads reached 2.1x on 8 GPUs).  The
    # workers live for the whole table, so their start-up (CUDA context, model, page-locked buffers) is paid once; tables too
    # short to amortise it and ORCAI_B200_TABLE_PROCESSES=0 use worker THREADS instead (ORCAI_B200_TABLE_PROCESSES=1 forces processes).
    # processes pay ~5 s of start-up each (in parallel): worth it from ~12 s of single-GPU work on, i.e. ~100 GB of WAV files
    procs_default = "0"
    if is_table and len(devices) > 1 and not _worker:
        try:
            t = pd.read_csv(recording_path, usecols=["base_dir_recording", "rel_recording_path"])
            base = str(base_dir_recording) if base_dir_recording is not None else None
            total = 0
            for b, r in zip(t["base_dir_recording"], t["rel_recording_path"]):
                try:
                    total += os.path.getsize(Path(base if base is not None else b).joinpath(r))
                except OSError:
                    pass
            procs_default = "1" if total >= 100e9 else "0"
        except Exception:  # noqa: BLE001 - an unreadable table fails below, in the normal path
            procs_default = "0"
    multi = is_table and len(devices) > 1 and not _worker and os.environ.get("ORCAI_B200_TABLE_PROCESSES", procs_default) == "1"
    if not multi:
        devices = list(dict.fromkeys(devices))   # one context per device in this process: a device listed twice counts once
        msgr.part(f"Loading model: {model_dir.stem}")
        model, orcai_parameter, shape = load_orcai_model(model_dir, device=devices[0] if devices else None)
        if hasattr(model, "describe"):
            msgr.info(model.describe())

    if recording_path.suffix == ".wav":
        return _predict_and_save(
            recording_path=recording_path,
            channel=channel,
            model=model,
            orcai_parameter=orcai_parameter,
            shape=shape,
            output_path=output_path,
            overwrite=overwrite,
            save_probabilities=save_probabilities,
            call_duration_limits=call_duration_limits,
            label_suffix=label_suffix,
            msgr=msgr,
            progressbar=None,
            _time_split=True,
        )
    elif recording_path.suffix == ".csv":
        recording_table = pd.read_csv(recording_path)
    else:
        raise ValueError("Recording file must be a wav or csv file")

    if base_dir_recording is not None:
        recording_table["base_dir_recording"] = str(base_dir_recording)
    if (output_path is not None) & (output_path != "default"):
        out_paths = [Path(output_path).joinpath(rec + "_" + model_dir.stem + "_predicted.txt") for rec in recording_table["recording"]]
    else:
        out_paths = [output_path] * len(recording_table)

    msgr.part(f"Predicting annotations for {len(recording_table)} wav files")
    rows = list(recording_table.index)
    if shard:
        from orcai_b200.sharding import assign_rows, recording_costs

        rank, world = shard
        plan = assign_rows(recording_costs([Path(recording_table.loc[i, "base_dir_recording"]).joinpath(recording_table.loc[i, "rel_recording_path"]) for i in rows]), world)
        rows = [rows[j] for j in plan[rank]]
        msgr.info(f"worker {rank} of {world}: {len(rows)} of {len(recording_table)} recordings")
    if multi:
        return _predict_table_multiprocess(
            recording_table, rows, devices, msgr,
            dict(recording_path=recording_path, channel=channel, model_dir=model_dir, output_path=output_path, overwrite=overwrite,
                 save_probabilities=save_probabilities, base_dir_recording=base_dir_recording, call_duration_limits=call_duration_limits,
                 label_suffix=label_suffix, verbosity=0),
        )
    if len(devices) > max(1, len(rows)):
        devices = devices[: max(1, len(rows))]      # no more worker contexts than recordings
    progressbar = None if _worker else tqdm(total=len(rows), desc="Starting ...", unit="file")

    out_path_of = dict(zip(recording_table.index, out_paths))

    def row_path(i):
        return Path(recording_table.loc[i, "base_dir_recording"]).joinpath(recording_table.loc[i, "rel_recording_path"])

    def report(i, e):
        text = f"Error predicting {recording_table.loc[i, 'recording']}: {e.args[0] if e.args else e}"
        if _worker:
            _worker["queue"].put(("error", text))
        else:
            msgr.error(text)

    def pipelined(mdl, my_rows, pb, tick):
        """One GPU's share of the table as a pipeline around the device that never lets it idle:
        loader threads decode WAV files straight into page-locked buffers (a 1-h file takes ~70 ms to read, twice the device time,
        hence several loaders) -> this thread uploads recording k+1 (DMA on the copy stream: orcai_prefetch_pcm / orcai_swap_pcm)
        and ENQUEUES its whole annotation (orcai_predict_resident_begin) while the device still works on recording k -> it then
        sleeps until recording k is done (orcai_predict_resident_end), hands its segments to a writer thread (label table, duration
        filter, files) and goes on.  Every failure of row k - load, hand-over, annotation, files - is reported against row k, like
        the reference loop (predict.py:752-755); the rows after it still run."""
        from collections import deque

        from orcai_b200._lib import PinnedPool

        ctx = _context_of(mdl, orcai_parameter, shape)
        sp = orcai_parameter["spectrogram"]
        if ctx.params.n_freq != shape["input_shape"][1]:
            raise ValueError(f"Spectrogram shape ({ctx.params.n_freq}) not equal to input shape ({shape['input_shape'][1]})")
        times01 = frames_to_time(2, sp)
        delta_t = times01[1] - times01[0]
        quiet = Messenger(verbosity=0)
        n_load = 3
        depth = n_load + 1
        # the pool lives on the context: page-locking 346 MB takes ~0.1 s, a table run must not pay it per call
        pool = ctx.__dict__.get("_pinned_pool")
        if pool is None:
            pool = ctx.__dict__["_pinned_pool"] = PinnedPool(max_free=depth + 3)

        def load(i, read_threads=1):
            taken = []

            def alloc(nbytes):
                a = pool.take(nbytes)
                taken.append(a)
                return a

            t0 = time.perf_counter()
            try:
                return load_recording(row_path(i), int(recording_table.loc[i, "channel"]), sp, quiet, alloc=alloc, read_threads=read_threads), taken
            except Exception as e:  # surfaced in the main loop so that the row is reported like any other failure
                for a in taken:
                    pool.give(a)
                return e, []
            finally:
                with stats_lock:
                    stats["load_s"] += time.perf_counter() - t0

        def give(taken):
            for a in taken:
                pool.give(a)

        stats = {"rows": 0, "load_s": 0.0, "wait_loader_s": 0.0, "wait_device_s": 0.0, "wait_writer_s": 0.0, "device": getattr(ctx, "device", None)}
        stats_lock = threading.Lock()
        pending = []      # (row, future of its host-side tail)
        in_flight = deque()   # (row, output path, token of predict_begin, page-locked buffers of its samples), oldest first
        with ThreadPoolExecutor(max_workers=n_load) as loaders, ThreadPoolExecutor(max_workers=1) as writer:
            # nothing overlaps the read of the share's first recording: it is read by four concurrent slices
            futs = {k: loaders.submit(load, my_rows[k], 4 if k == 0 else 1) for k in range(min(depth, len(my_rows)))}

            def fetch(k):
                """wait for recording k on the host and start its upload"""
                t0 = time.perf_counter()
                res, taken = futs.pop(k).result()
                stats["wait_loader_s"] += time.perf_counter() - t0
                if not isinstance(res, BaseException):
                    try:
                        ctx.prefetch_pcm(res)
                    except Exception as e:
                        give(taken)
                        return e, []
                return res, taken

            def begin(i, res):
                """hand-over of row i's samples and its whole annotation, enqueued behind whatever the device is doing"""
                if isinstance(res, BaseException):
                    raise res
                rp = row_path(i)
                out = _resolved_output_path(rp, int(recording_table.loc[i, "channel"]), orcai_parameter, out_path_of[i], overwrite, quiet)
                ctx.swap_pcm()
                try:
                    return out, ctx.predict_begin(res.size, threshold=0.5, want_agg=bool(save_probabilities))
                except OrcaiError as e:
                    if e.code == ORCAI_ERR_TOO_SHORT:
                        raise ValueError(f"{rp.stem}: {e.message}") from e
                    raise

            def collect():
                """sleep until the oldest row in flight is done; its host-side tail goes to the writer thread"""
                i, out, token, taken = in_flight.popleft()
                try:
                    t0 = time.perf_counter()
                    _stats, agg, _cnt, lab, sta, sto = ctx.predict_end(token)
                    stats["wait_device_s"] += time.perf_counter() - t0
                    stats["rows"] += 1
                    pending.append((i, writer.submit(_host_tail, agg, lab, sta, sto, delta_t, orcai_parameter, out, save_probabilities,
                                                     call_duration_limits, label_suffix, quiet)))
                except Exception as e:
                    report(i, e)
                give(taken)
                tick()

            cur = fetch(0) if my_rows else None
            for k, i in enumerate(my_rows):
                res, taken = cur
                begun = None
                try:
                    begun = begin(i, res)                   # queued behind row k-1, which is still on the device
                except Exception as e:
                    report(i, e)
                if in_flight:
                    collect()                               # row k-1: done -> its device PCM buffer is free for row k+1's upload
                if begun is not None:
                    in_flight.append((i, begun[0], begun[1], taken))
                else:
                    give(taken)
                    tick()
                if k + depth < len(my_rows):
                    futs[k + depth] = loaders.submit(load, my_rows[k + depth])
                try:
                    cur = fetch(k + 1) if k + 1 < len(my_rows) else None   # its upload overlaps the annotation of row k
                except Exception as e:
                    cur = (e, [])
            while in_flight:
                collect()
            t0 = time.perf_counter()
            for i, fut in pending:
                try:
                    fut.result()
                except Exception as e:
                    report(i, e)
            stats["wait_writer_s"] = time.perf_counter() - t0
        TABLE_STATS.clear()
        TABLE_STATS.update(stats)

    if _worker:
        q = _worker["queue"]
        mine = set(_worker["rows"])
        pipelined(model, [i for i in rows if i in mine], None, lambda: q.put(("tick", None)))
        return None
    if len(devices) <= 1:
        pipelined(model, rows, progressbar, lambda: progressbar.update(1))
    else:
        # shard by recording: one worker thread (context + streams) per GPU; rows are assigned
        # longest-first (LPT on the file sizes) and every worker pipelines its own share; host-side gather only
        from orcai_b200.model import OrcaiModel
        from orcai_b200.sharding import assign_rows, recording_costs

        models = [model] + [OrcaiModel(orcai_parameter, shape, model.weights, device=d) for d in devices[1:]]
        plan = assign_rows(recording_costs([row_path(i) for i in rows]), len(models))
        lock = threading.Lock()

        def tick():
            with lock:
                progressbar.update(1)

        failures = []

        def guarded(m, share):
            try:
                pipelined(m, share, None, tick)
            except BaseException as e:  # noqa: BLE001 - a dead worker must not look like a finished table
                failures.append((m.ctx.device if hasattr(m.ctx, "device") else "?", e, share))

        threads = [threading.Thread(target=guarded, args=(m, [rows[j] for j in share]), daemon=True) for m, share in zip(models, plan)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for dev, e, share in failures:
            msgr.error(f"Error in table worker on device {dev}: {type(e).__name__}: {e} ({len(share)} recordings assigned, some not processed)")
    progressbar.close()
    msgr.success("Predictions finished.")
