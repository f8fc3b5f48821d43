"""Spectrogram stage with the reference's function surface (``src/orcAI/spectrogram.py``).

``make_spectrogram`` / ``save_spectrogram`` / ``create_spectrograms`` keep the signatures, return
types, file layout and filtering rules of the reference (spectrogram.py:90-321); the arithmetic
(STFT, dB, crop, exact percentiles, clip, normalise) runs in liborcai_b200 on the GPU.
"""

from __future__ import annotations

from importlib.resources import files
from pathlib import Path

import numpy as np
import pandas as pd
from tqdm import tqdm

from orcai_b200.auxiliary import Messenger
from orcai_b200.io import read_json, save_as_zarr, write_vector_to_json
from orcai_b200.runtime import get_context, shape_for
from orcai_b200.wavio import read_wav


def load_recording(wav_file_path: Path | str, channel: int, spectrogram_parameter: dict, msgr: Messenger | None = None, alloc=None,
                   read_threads: int = 1) -> np.ndarray:
    """Mono samples of the requested channel at the model's sampling rate (int16 or float32).

    Stands in for ``librosa.load(path, sr=..., mono=False)`` + channel selection (spectrogram.py:23-31).
    ``alloc``: see ``wavio.read_wav`` (table mode decodes into page-locked buffers).
    """
    samples, sr, n_ch = read_wav(wav_file_path, channel, alloc=alloc, read_threads=read_threads)
    if n_ch > 1 and msgr is not None:
        msgr.warning(f"Multiple channels found, using channel {channel}")
    target = int(spectrogram_parameter["sampling_rate"])
    if sr != target:
        samples = resample(samples, sr, target)
        if msgr is not None:
            msgr.warning(
                f"{Path(wav_file_path).name}: resampled {sr} -> {target} Hz with a polyphase Kaiser filter on the host "
                "(the reference uses soxr_hq; results agree only to resampler tolerance)"
            )
    return samples


def resample(samples: np.ndarray, sr: int, target: int) -> np.ndarray:
    """Host-side rational resampling to the model's rate -> float32 in [-1, 1] scale.

    The reference's ``librosa.load(sr=...)`` resamples with soxr_hq (spectrogram.py:23-27), a third-party resampler that is
    not available here and has no in-tree specification, so this step is NOT parity-pinned:
    ``scipy.signal.resample_poly`` (polyphase FIR, Kaiser beta 5) in float64.
    """
    from math import gcd

    from scipy.signal import resample_poly

    x = np.asarray(samples)
    if x.dtype == np.int16:
        x = x.astype(np.float64) / 32768.0
    g = gcd(int(sr), int(target))
    y = resample_poly(x.astype(np.float64), int(target) // g, int(sr) // g)
    return np.ascontiguousarray(y, dtype=np.float32)


def fft_frequencies(spectrogram_parameter: dict) -> np.ndarray:
    return np.fft.rfftfreq(n=spectrogram_parameter["nfft"], d=1.0 / spectrogram_parameter["sampling_rate"])


def frames_to_time(n_frames: int, spectrogram_parameter: dict) -> np.ndarray:
    return (np.arange(n_frames) * spectrogram_parameter["n_overlap"]) / float(spectrogram_parameter["sampling_rate"])


def make_spectrogram(
    wav_file_path: Path | str,
    channel: int = 1,
    orcai_parameter: (Path | str) | dict = files("orcai_b200.defaults").joinpath("default_orcai_parameter.json"),
    verbosity: int = 2,
    msgr: Messenger | None = None,
    device: int | None = None,
) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Makes the normalised (T, n_band) float32 spectrogram of a .wav file.

    Returns (spectrogram, frequencies (n_fft/2+1,), times (T,)) like the reference.
    """
    if msgr is None:
        msgr = Messenger(verbosity=verbosity, title="Making spectrogram")
    if not isinstance(orcai_parameter, dict):
        orcai_parameter = read_json(orcai_parameter)
    sp = orcai_parameter["spectrogram"]
    wav_file_path = Path(wav_file_path)

    msgr.part("Calculating power spectrogram by stft")
    msgr.info(f"Loading & resampling (to {sp['sampling_rate'] / 1000:.2f} kHz) wav file: {wav_file_path.stem}")
    samples = load_recording(wav_file_path, channel, sp, msgr)
    ctx = get_context(orcai_parameter, shape_for(orcai_parameter), device)
    spectrogram, stats = ctx.spectrogram(samples)
    frequencies = fft_frequencies(sp)
    times = frames_to_time(int(stats.n_frames), sp)
    msgr.info(f"Duration of wav file: {times[-1]:.2f} seconds")
    msgr.info("Extracting frequency range and clipping spectrogram")
    return spectrogram, frequencies, times


def save_spectrogram(
    spectrogram: np.ndarray,
    frequencies: np.ndarray,
    times: np.ndarray,
    output_dir: Path | str,
    verbosity: int = 2,
    msgr: Messenger | None = None,
) -> None:
    """Saves the spectrogram as zarr store plus frequencies.json / times.json in output_dir."""
    if msgr is None:
        msgr = Messenger(verbosity=verbosity, title="Saving spectrogram")
    msgr.part("Saving spectrogram")
    Path(output_dir).mkdir(parents=True, exist_ok=True)
    save_as_zarr(spectrogram, filename=Path(output_dir, "spectrogram.zarr"))
    write_vector_to_json(frequencies, Path(output_dir, "frequencies.json"))
    write_vector_to_json(times, Path(output_dir, "times.json"))


def _make_and_save_spectrogram(recording_info, orcai_parameter, output_dir, device=None):
    silent = Messenger(verbosity=0)
    spectrogram, frequencies, times = make_spectrogram(
        Path(recording_info.base_dir_recording).joinpath(recording_info.rel_recording_path),
        recording_info.channel,
        orcai_parameter,
        msgr=silent,
        device=device,
    )
    save_spectrogram(spectrogram, frequencies, times, Path(output_dir).joinpath(recording_info.recording, "spectrogram"), msgr=silent)
    return recording_info.recording


def create_spectrograms(
    recording_table_path: Path | str,
    output_dir: Path | str,
    base_dir_recording: Path | str | None = None,
    orcai_parameter: Path | str | dict | None = files("orcai_b200.defaults").joinpath("default_orcai_parameter.json"),
    include_not_annotated: bool = False,
    include_no_possible_annotations: bool = False,
    overwrite: bool = False,
    verbosity: int = 2,
    msgr: Messenger | None = None,
) -> None:
    """Creates spectrograms for all files of a recording table (same filters and layout as the reference)."""
    if msgr is None:
        msgr = Messenger(verbosity=verbosity, title="Creating spectrograms")
    msgr.part("Reading recordings table")
    recording_table = pd.read_csv(recording_table_path)
    output_dir = Path(output_dir)
    if not isinstance(orcai_parameter, dict):
        orcai_parameter = read_json(orcai_parameter)

    if not include_not_annotated:
        not_annotated = recording_table["base_dir_annotation"].isna()
        if len(not_annotated) > 0:
            msgr.info(f"Excluded {not_annotated.sum()} recordings because they are not annotated.")
            recording_table = recording_table[~not_annotated]
    if not include_no_possible_annotations:
        is_included = recording_table[orcai_parameter["calls"]].apply(lambda x: x.any(), axis=1)
        if sum(~is_included) > 0:
            msgr.info("Excluded recordings because they lack any possible annotations:", indent=1)
            msgr.info(str(recording_table[~is_included]["recording"].values), indent=-1)
            recording_table = recording_table[is_included]
    if not overwrite:
        existing = recording_table["recording"].apply(lambda x: output_dir.joinpath(x, "spectrogram").exists())
        if sum(existing) > 0:
            msgr.info(f"Skipping {sum(existing)} recordings because they already have spectrograms.")
            recording_table = recording_table[~existing]
    if base_dir_recording is not None:
        recording_table = recording_table.assign(base_dir_recording=str(base_dir_recording))

    # Sharding by recording (the reference loop, spectrogram.py:313-318, is sequential; every recording is independent):
    #   * one worker of a multi-process job (torchrun / ORCAI_B200_SHARD, see predict._rank_shard) takes its longest-first share;
    #   * otherwise one worker thread per visible GPU (ORCAI_B200_DEVICES narrows the set) - the host side of a recording is the
    #     WAV read, the 462 MB/h read-back and the gzip of the zarr chunks, all of which release the interpreter lock.
    import threading

    from orcai_b200.predict import _rank_shard, _table_devices
    from orcai_b200.sharding import assign_rows, recording_costs

    recordings = list(recording_table.itertuples(index=False))
    costs = recording_costs([Path(r.base_dir_recording).joinpath(r.rel_recording_path) for r in recordings])
    shard = _rank_shard()
    if shard:
        rank, world = shard
        recordings = [recordings[j] for j in assign_rows(costs, world)[rank]]
        costs = recording_costs([Path(r.base_dir_recording).joinpath(r.rel_recording_path) for r in recordings])
        devices = [None]
        msgr.info(f"worker {rank} of {world}: {len(recordings)} recordings")
    else:
        devices = _table_devices()[: max(1, len(recordings))] or [None]
    msgr.part(f"Creating {len(recordings)} spectrograms")
    progress = tqdm(desc="Making spectrograms", total=len(recordings))
    if len(devices) <= 1:
        for recording in recordings:
            _make_and_save_spectrogram(recording, orcai_parameter, output_dir, device=devices[0])
            progress.update(1)
    else:
        lock, failures = threading.Lock(), []

        def work(dev, share):
            try:
                for j in share:
                    _make_and_save_spectrogram(recordings[j], orcai_parameter, output_dir, device=dev)
                    with lock:
                        progress.update(1)
            except BaseException as e:  # noqa: BLE001 - re-raised below, like the sequential loop would
                failures.append(e)

        threads = [threading.Thread(target=work, args=(d, sh), daemon=True) for d, sh in zip(devices, assign_rows(costs, len(devices)))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if failures:
            raise failures[0]
    progress.close()
    msgr.success("Spectrograms created.")
