"""In-house RIFF/WAVE reader (the reference reads audio through librosa.load -> soundfile,
``src/orcAI/spectrogram.py:23-31``; neither is available here).

Supported: RIFF and RF64 containers; PCM 8/16/24/32 bit, IEEE float 32/64, WAVE_FORMAT_EXTENSIBLE;
any channel count.  Conversion to float follows libsndfile: PCM16 / 2^15, PCM24 / 2^23,
PCM32 / 2^31, unsigned 8 bit (x-128) / 2^7, floats unchanged.  PCM16 mono is handed to the
GPU as raw int16 (half the upload); the kernel applies the same exact power-of-two scale.

This module never resamples.  The reference resamples through soxr_hq when the file rate differs
from orcai_parameter["spectrogram"]["sampling_rate"]; that resampler has no in-tree specification
(SURVEY.md section 8f rank 3), so ``orcai_b200.spectrogram.load_recording`` stands in with a
polyphase resampler (scipy.signal.resample_poly) - the one step of the path whose parity is unpinned.
"""

from __future__ import annotations

import struct
from dataclasses import dataclass
from pathlib import Path

import numpy as np

WAVE_FORMAT_PCM = 0x0001
WAVE_FORMAT_IEEE_FLOAT = 0x0003
WAVE_FORMAT_EXTENSIBLE = 0xFFFE


@dataclass
class WavInfo:
    sample_rate: int
    channels: int
    bits: int
    is_float: bool
    n_frames: int
    data_offset: int


def read_wav_info(path: Path | str) -> WavInfo:
    with open(path, "rb") as f:
        head = f.read(12)
        if len(head) < 12 or head[8:12] != b"WAVE" or head[0:4] not in (b"RIFF", b"RF64"):
            raise ValueError(f"{path}: not a RIFF/RF64 WAVE file")
        rf64 = head[0:4] == b"RF64"
        fmt = None
        data_size64 = None
        while True:
            hdr = f.read(8)
            if len(hdr) < 8:
                raise ValueError(f"{path}: no data chunk")
            cid, size = hdr[0:4], struct.unpack("<I", hdr[4:8])[0]
            if cid == b"ds64":
                body = f.read(size)
                data_size64 = struct.unpack("<Q", body[8:16])[0]
            elif cid == b"fmt ":
                body = f.read(size)
                tag, ch, sr, _, _, bits = struct.unpack("<HHIIHH", body[:16])
                if tag == WAVE_FORMAT_EXTENSIBLE and size >= 26:
                    tag = struct.unpack("<H", body[24:26])[0]
                fmt = (tag, ch, sr, bits)
            elif cid == b"data":
                if fmt is None:
                    raise ValueError(f"{path}: data chunk before fmt chunk")
                if rf64 and size == 0xFFFFFFFF and data_size64 is not None:
                    size = data_size64
                tag, ch, sr, bits = fmt
                if tag not in (WAVE_FORMAT_PCM, WAVE_FORMAT_IEEE_FLOAT):
                    raise ValueError(f"{path}: unsupported WAVE format tag {tag:#x}")
                offset = f.tell()
                end = f.seek(0, 2)
                size = min(size, end - offset)
                bpf = ch * bits // 8
                return WavInfo(sr, ch, bits, tag == WAVE_FORMAT_IEEE_FLOAT, size // bpf, offset)
            else:
                f.seek(size, 1)
            if size & 1:
                f.seek(1, 1)


def _read_slices(path, offset: int, out: memoryview, threads: int) -> int:
    """Fill ``out`` from ``path`` at ``offset`` with ``threads`` concurrent positional reads (the copy out of the page cache runs at
    ~7 GB/s per core; os.preadv releases the interpreter lock) -> bytes read."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    n = len(out)
    step = -(-n // threads + 4095) // 4096 * 4096 if n else 0
    fd = os.open(path, os.O_RDONLY)
    try:
        def part(k):
            lo, hi = k * step, min(n, (k + 1) * step)
            got = 0
            while lo + got < hi:
                r = os.preadv(fd, [out[lo + got:hi]], offset + lo + got)
                if r <= 0:
                    break
                got += r
            return got

        with ThreadPoolExecutor(max_workers=threads) as ex:
            return sum(ex.map(part, range(-(-n // step) if step else 0)))
    finally:
        os.close(fd)


def read_wav(path: Path | str, channel: int = 1, alloc=None, read_threads: int = 1) -> tuple[np.ndarray, int, int]:
    """Return (mono samples, sample_rate, n_channels).

    ``channel`` is 1-indexed and only used for multi-channel files (reference
    spectrogram.py:29-31).  The result is int16 for PCM16 input, else float32.
    ``alloc(nbytes) -> uint8 ndarray``: where to put the samples (e.g. page-locked memory, ``_lib.PinnedPool.take``); a mono
    PCM16 file is read straight into it (by ``read_threads`` concurrent positional reads when > 1: the first recording of a table,
    whose read nothing overlaps), everything else is decoded first and then copied.
    """
    info = read_wav_info(path)
    if alloc is not None:
        if info.channels == 1 and info.bits == 16 and not info.is_float:
            n = info.n_frames
            buf = alloc(2 * n)
            if read_threads > 1 and n > (1 << 22):
                got = _read_slices(path, info.data_offset, memoryview(buf)[: 2 * n], read_threads)
            else:
                with open(path, "rb") as f:
                    f.seek(info.data_offset)
                    got = f.readinto(memoryview(buf)[: 2 * n])
            if got != 2 * n:
                raise ValueError(f"{path}: truncated data chunk ({got} of {2 * n} bytes)")
            return buf[: 2 * n].view("<i2"), info.sample_rate, 1
        samples, sr, ch = read_wav(path, channel)
        buf = alloc(samples.nbytes)
        out = buf[: samples.nbytes].view(samples.dtype)
        out[:] = samples
        return out, sr, ch
    ch = info.channels
    if ch > 1 and not (1 <= channel <= ch):
        raise IndexError(f"channel {channel} out of range for a {ch}-channel file")
    col = channel - 1 if ch > 1 else 0
    n = info.n_frames
    with open(path, "rb") as f:
        f.seek(info.data_offset)
        if info.is_float:
            dt = {32: "<f4", 64: "<f8"}.get(info.bits)
            if dt is None:
                raise ValueError(f"{path}: {info.bits}-bit float WAVE not supported")
            a = np.fromfile(f, dtype=dt, count=n * ch).reshape(-1, ch)[:, col]
            return np.ascontiguousarray(a, dtype=np.float32), info.sample_rate, ch
        if info.bits == 16:
            a = np.fromfile(f, dtype="<i2", count=n * ch).reshape(-1, ch)[:, col]
            return np.ascontiguousarray(a), info.sample_rate, ch
        if info.bits == 8:
            a = np.fromfile(f, dtype=np.uint8, count=n * ch).reshape(-1, ch)[:, col]
            return ((a.astype(np.float32) - 128.0) / 128.0).astype(np.float32), info.sample_rate, ch
        if info.bits == 32:
            a = np.fromfile(f, dtype="<i4", count=n * ch).reshape(-1, ch)[:, col]
            return (a.astype(np.float64) / 2147483648.0).astype(np.float32), info.sample_rate, ch
        if info.bits == 24:
            raw = np.fromfile(f, dtype=np.uint8, count=n * ch * 3).reshape(-1, ch, 3)[:, col, :].astype(np.int32)
            v = raw[:, 0] | (raw[:, 1] << 8) | (raw[:, 2] << 16)
            v = np.where(v & 0x800000, v - (1 << 24), v)
            return (v.astype(np.float32) / np.float32(8388608.0)).astype(np.float32), info.sample_rate, ch
        raise ValueError(f"{path}: {info.bits}-bit PCM WAVE not supported")


def write_wav_pcm16(path: Path | str, pcm: np.ndarray, sample_rate: int = 48000) -> None:
    """Write int16 samples, shape (n,) or (n, channels); RF64 when the data exceeds 4 GiB."""
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    ch = 1 if pcm.ndim == 1 else pcm.shape[1]
    nbytes = pcm.size * 2
    fmt = struct.pack("<HHIIHH", WAVE_FORMAT_PCM, ch, sample_rate, sample_rate * ch * 2, ch * 2, 16)
    with open(path, "wb") as f:
        if nbytes + 36 < 0xFFFFFFFF:
            f.write(b"RIFF" + struct.pack("<I", 36 + nbytes) + b"WAVE")
            f.write(b"fmt " + struct.pack("<I", 16) + fmt)
            f.write(b"data" + struct.pack("<I", nbytes))
        else:
            f.write(b"RF64" + struct.pack("<I", 0xFFFFFFFF) + b"WAVE")
            f.write(b"ds64" + struct.pack("<I", 28) + struct.pack("<QQQI", nbytes + 72, nbytes, pcm.size // ch, 0))
            f.write(b"fmt " + struct.pack("<I", 16) + fmt)
            f.write(b"data" + struct.pack("<I", 0xFFFFFFFF))
        pcm.tofile(f)
