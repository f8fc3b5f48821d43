"""ctypes binding of liborcai_b200.so (C ABI declared in include/orcai_b200.h).

There is no CPU fallback: if the shared library is missing, or no sm_100 device is
present, every compute entry point raises.  ``load_library()`` alone needs no GPU
(it only dlopens and resolves symbols), which is what the CPU test-suite checks.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "liborcai_b200.so"

ORCAI_OK = 0
ORCAI_ERR_ARG = -1
ORCAI_ERR_CUDA = -2
ORCAI_ERR_STATE = -3
ORCAI_ERR_CAPACITY = -4
ORCAI_ERR_TOO_SHORT = -5
PCM_I16 = 0
PCM_F32 = 1


class OrcaiError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"liborcai_b200 error {code}: {message}")
        self.code = code
        self.message = message


class Params(C.Structure):
    _fields_ = [
        ("sampling_rate", C.c_int32),
        ("n_fft", C.c_int32),
        ("hop", C.c_int32),
        ("band_lo", C.c_int32),
        ("band_hi", C.c_int32),
        ("q_lo", C.c_double),
        ("q_hi", C.c_double),
        ("snippet_len", C.c_int32),
        ("n_freq", C.c_int32),
        ("n_labels", C.c_int32),
        ("n_blocks", C.c_int32),
        ("filters", C.c_int32 * 8),
        ("kernel_size", C.c_int32),
        ("lstm_units", C.c_int32),
        ("reserved", C.c_int32 * 8),
    ]


class SpecStats(C.Structure):
    _fields_ = [
        ("n_frames", C.c_int64),
        ("ref_power", C.c_float),
        ("db_ref", C.c_float),
        ("lo", C.c_float),
        ("hi", C.c_float),
        ("rank_lo", C.c_int64),
        ("rank_hi", C.c_int64),
    ]


class Timings(C.Structure):
    _fields_ = [
        ("h2d_ms", C.c_float),
        ("stft_ms", C.c_float),
        ("select_ms", C.c_float),
        ("normalise_ms", C.c_float),
        ("network_ms", C.c_float),
        ("post_ms", C.c_float),
        ("d2h_ms", C.c_float),
        ("total_ms", C.c_float),
        ("kernel_launches", C.c_uint64),
        ("net_stage_ms", C.c_float * 16),
    ]

    def as_dict(self) -> dict:
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "net_stage_ms"}
        d["net_stage_ms"] = list(self.net_stage_ms)
        return d


_P = C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)  -- must list every symbol declared in include/orcai_b200.h
    "orcai_create": (C.c_int, [C.c_int, C.POINTER(Params), C.POINTER(_P)]),
    "orcai_destroy": (None, [_P]),
    "orcai_last_error": (C.c_char_p, [_P]),
    "orcai_get_timings": (C.c_int, [_P, C.POINTER(Timings)]),
    "orcai_version": (C.c_int, []),
    "orcai_load_weights": (C.c_int, [_P, C.POINTER(C.c_char_p), C.POINTER(_P), C.POINTER(C.c_int64), C.c_int32]),
    "orcai_num_frames": (C.c_int64, [C.c_int64, C.c_int32]),
    "orcai_upload_pcm": (C.c_int, [_P, _P, C.c_int32, C.c_int64]),
    "orcai_spectrogram_resident": (C.c_int, [_P, C.c_int32, C.POINTER(SpecStats)]),
    "orcai_spectrogram": (C.c_int, [_P, _P, C.c_int32, C.c_int64, _P, C.POINTER(SpecStats)]),
    "orcai_read_spectrogram": (C.c_int, [_P, C.c_int64, C.c_int64, _P]),
    "orcai_read_db": (C.c_int, [_P, C.c_int64, C.c_int64, _P]),
    "orcai_num_snippets": (C.c_int64, [C.c_int64, C.c_int32]),
    "orcai_forward_host": (C.c_int, [_P, _P, C.c_int64, _P]),
    "orcai_forward_resident": (C.c_int, [_P, C.c_int64, C.c_int64, _P]),
    "orcai_postprocess": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_double, _P, _P, _P, _P, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "orcai_threshold_segments": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_double, _P, _P, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "orcai_predict_resident": (C.c_int, [_P, C.c_double, C.POINTER(SpecStats), _P, _P, _P, _P, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "orcai_predict_resident_begin": (C.c_int, [_P, C.c_double, C.c_int32, C.c_int64]),
    "orcai_predict_resident_end": (C.c_int, [_P, C.POINTER(SpecStats), _P, _P, _P, _P, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "orcai_predict_in_flight": (C.c_int, [_P]),
    "orcai_predict_pcm": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_double, C.POINTER(SpecStats), _P, _P, _P, _P, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "orcai_set_option": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "orcai_calibrate": (C.c_int, [_P, C.c_int64]),
    "orcai_prefetch_pcm": (C.c_int, [_P, _P, C.c_int32, C.c_int64]),
    "orcai_swap_pcm": (C.c_int, [_P]),
    "orcai_debug_read": (C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "orcai_host_alloc": (_P, [C.c_size_t]),
    "orcai_host_free": (None, [_P]),
    "orcai_chunk_spectrogram": (C.c_int, [_P, C.c_int64, C.c_int64, C.POINTER(C.c_float)]),
    "orcai_chunk_select_begin": (C.c_int, [_P, C.c_float]),
    "orcai_chunk_histogram": (C.c_int, [_P, C.c_int32, C.c_int64, C.c_int64, _P, _P]),
    "orcai_chunk_select_end": (C.c_int, [_P, _P, C.POINTER(SpecStats)]),
}

_lib = None


def load_library() -> C.CDLL:
    """dlopen liborcai_b200.so and bind every exported entry point (no GPU needed)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise OrcaiError(
            ORCAI_ERR_STATE,
            f"{LIB_PATH} is missing - build it with `python -m orcai_b200.build` (sm_100a, nvcc). "
            "orcai_b200 has no CPU fallback.",
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols() -> list[str]:
    return list(_SIGNATURES)


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(_P)


def params_from_dicts(orcai_parameter: dict, shape: dict) -> Params:
    """Fill the C parameter block from orcai_parameter.json / model_shape.json contents."""
    sp = orcai_parameter["spectrogram"]
    sr, n_fft = int(sp["sampling_rate"]), int(sp["nfft"])
    freqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    # band selection rule of preprocess_spectrogram (reference spectrogram.py:62-68)
    band_lo = int(np.argwhere(freqs <= sp["freq_range"][0])[0][0])
    band_hi = int(np.argwhere(freqs >= sp["freq_range"][1])[0][0])
    p = Params()
    p.sampling_rate = sr
    p.n_fft = n_fft
    p.hop = int(sp["n_overlap"])
    p.band_lo, p.band_hi = band_lo, band_hi
    # np.percentile(x, 100*q) divides by 100 again before computing the nearest rank
    p.q_lo = float(np.true_divide(100 * sp["quantiles"][0], 100))
    p.q_hi = float(np.true_divide(100 * sp["quantiles"][1], 100))
    p.snippet_len = int(shape["input_shape"][0])
    p.n_freq = int(shape["input_shape"][1])
    p.n_labels = int(shape["num_labels"])
    filters = list(orcai_parameter["model"]["filters"])
    p.n_blocks = len(filters)
    for i, f in enumerate(filters[:8]):
        p.filters[i] = int(f)
    p.kernel_size = int(orcai_parameter["model"]["kernel_size"])
    p.lstm_units = int(orcai_parameter["model"]["lstm_units"])
    return p


class PinnedPool:
    """A few page-locked host buffers that recordings are decoded into (orcai_host_alloc): uploads from them are plain DMA.

    ``take(nbytes)`` -> uint8 ndarray of at least nbytes (keep it until ``give`` returns it); thread-safe.  Pinning memory is
    slow (~0.3 ms/MB), so buffers are recycled; a buffer that is too small is replaced by a larger one.
    """

    def __init__(self, max_free: int = 6):
        import threading

        self.lib = load_library()
        self._free: list[np.ndarray] = []
        self._ptr: dict[int, int] = {}       # id(buffer) -> host pointer
        self._keep: dict[int, np.ndarray] = {}   # keeps every live buffer's array object (and thereby its id) alive
        self._lock = threading.Lock()
        self._max_free = max_free

    def _alloc(self, nbytes: int) -> np.ndarray:
        nbytes = max(int(nbytes), 1)
        nbytes = (nbytes + (1 << 20) - 1) & ~((1 << 20) - 1)
        p = self.lib.orcai_host_alloc(nbytes)
        if not p:
            raise MemoryError(f"orcai_host_alloc({nbytes}) failed")
        a = np.ctypeslib.as_array((C.c_ubyte * nbytes).from_address(p))
        self._ptr[id(a)] = p
        self._keep[id(a)] = a
        return a

    def take(self, nbytes: int) -> np.ndarray:
        with self._lock:
            best = None
            for i, a in enumerate(self._free):
                if a.size >= nbytes and (best is None or a.size < self._free[best].size):
                    best = i
            if best is not None:
                return self._free.pop(best)
            if len(self._free) >= self._max_free:      # every free buffer is too small: drop the smallest
                self._release(self._free.pop(min(range(len(self._free)), key=lambda j: self._free[j].size)))
        return self._alloc(nbytes)

    def give(self, a: np.ndarray) -> None:
        with self._lock:
            if id(a) in self._ptr:
                self._free.append(a)

    def _release(self, a: np.ndarray) -> None:
        p = self._ptr.pop(id(a), None)
        self._keep.pop(id(a), None)
        if p:
            self.lib.orcai_host_free(p)

    def close(self) -> None:
        with self._lock:
            for a in self._free:
                self._release(a)
            self._free.clear()


class Context:
    """One liborcai_b200 context = one device, one stream, one resident recording."""

    def __init__(self, orcai_parameter: dict, shape: dict, device: int = 0):
        self.lib = load_library()
        self.params = params_from_dicts(orcai_parameter, shape)
        self.orcai_parameter = orcai_parameter
        self.shape = shape
        h = _P()
        rc = self.lib.orcai_create(int(device), C.byref(self.params), C.byref(h))
        if rc != ORCAI_OK:
            raise OrcaiError(rc, (self.lib.orcai_last_error(None) or b"").decode())
        self._h = h
        self.device = device
        self.pred_len = self.params.snippet_len >> self.params.n_blocks
        self.weights_loaded = False

    # -- plumbing ------------------------------------------------------------------------------
    def close(self):
        pool = self.__dict__.pop("_pinned_pool", None)
        if pool is not None:
            pool.close()
        if getattr(self, "_h", None):
            self.lib.orcai_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != ORCAI_OK:
            raise OrcaiError(rc, (self.lib.orcai_last_error(self._h) or b"").decode())

    def timings(self) -> dict:
        t = Timings()
        self._check(self.lib.orcai_get_timings(self._h, C.byref(t)))
        return t.as_dict()

    def set_option(self, key: str, value: int):
        self._check(self.lib.orcai_set_option(self._h, key.encode(), int(value)))

    def debug_stage(self, snippets: np.ndarray, stage: int) -> np.ndarray:
        """Test hook: run the forward up to `stage` and return that stage's activations (n, h, w, c) float32."""
        x = np.ascontiguousarray(snippets, dtype=np.float32)
        self.set_option("debug_stop", stage)
        try:
            dummy = np.empty((x.shape[0], self.pred_len, self.params.n_labels), dtype=np.float32)
            self._check(self.lib.orcai_forward_host(self._h, _ptr(x), x.shape[0], _ptr(dummy)))
            cap = x.shape[0] * x.shape[1] * x.shape[2] * 64
            out = np.empty(cap, dtype=np.float32)
            dims = (C.c_int64 * 4)()
            self._check(self.lib.orcai_debug_read(self._h, _ptr(out), cap, dims))
            n, h, w, c = (int(v) for v in dims)
            return out[: n * h * w * c].reshape(n, h, w, c).copy()
        finally:
            self.set_option("debug_stop", -1)

    # -- weights -------------------------------------------------------------------------------
    def load_weights(self, W: dict):
        names = list(W)
        arrs = [np.ascontiguousarray(W[k], dtype=np.float32) for k in names]
        n = len(names)
        c_names = (C.c_char_p * n)(*[k.encode() for k in names])
        c_data = (_P * n)(*[a.ctypes.data for a in arrs])
        c_sizes = (C.c_int64 * n)(*[a.size for a in arrs])
        self._check(self.lib.orcai_load_weights(self._h, c_names, c_data, c_sizes, n))
        self.weights_loaded = True
        self.owner = None   # the OrcaiModel whose weights / options are bound (model.bind()); raw loads belong to nobody

    def calibrate(self, pcm: np.ndarray | None = None, max_snippets: int = 8):
        """Fold the mean effect of fp16 weight rounding into the biases of the tensor-core operands (orcai_calibrate).

        `pcm`: a representative recording (>= one snippet); default: a built-in deterministic synthetic recording.
        Replaces the context's resident recording.
        """
        if pcm is None:
            from orcai_b200.synth import synth_pcm16

            pcm = synth_pcm16(24.0, seed=918273645, calls_per_minute=30.0)   # a seed no test, smoke or bench recording uses
        self.upload_pcm(pcm)
        self.spectrogram_resident(False)
        self._check(self.lib.orcai_calibrate(self._h, int(max_snippets)))

    # -- spectrogram ---------------------------------------------------------------------------
    @staticmethod
    def _pcm(pcm: np.ndarray):
        pcm = np.ascontiguousarray(pcm)
        if pcm.ndim != 1:
            raise ValueError("pcm must be mono (1-D)")
        if pcm.dtype == np.int16:
            return pcm, PCM_I16
        if pcm.dtype == np.float32:
            return pcm, PCM_F32
        raise ValueError(f"pcm dtype {pcm.dtype} not supported (int16 or float32)")

    def upload_pcm(self, pcm: np.ndarray):
        pcm, dt = self._pcm(pcm)
        self._check(self.lib.orcai_upload_pcm(self._h, _ptr(pcm), dt, pcm.size))

    def prefetch_pcm(self, pcm: np.ndarray):
        """Start the host->device copy of the NEXT recording on the copy stream (returns at once; keep `pcm` alive)."""
        pcm, dt = self._pcm(pcm)
        self._prefetched = pcm
        self._check(self.lib.orcai_prefetch_pcm(self._h, _ptr(pcm), dt, pcm.size))

    def swap_pcm(self):
        """Make the prefetched recording the resident one (the compute stream waits for the copy, the host does not)."""
        self._check(self.lib.orcai_swap_pcm(self._h))

    def predict_stream(self, recordings, threshold: float = 0.5, want_agg: bool = True):
        """Annotate a sequence of PCM arrays; the upload of recording k+1 overlaps the annotation of recording k.

        Yields what predict_pcm returns, in order.  (Recording tables: predict.py:733-755.)
        """
        it = iter(recordings)
        try:
            cur = next(it)
        except StopIteration:
            return
        self.prefetch_pcm(cur)
        while cur is not None:
            self.swap_pcm()
            nxt = next(it, None)
            if nxt is not None:
                self.prefetch_pcm(nxt)
            yield self.predict_pcm(cur, threshold=threshold, want_agg=want_agg, resident=True)
            cur = nxt

    def spectrogram_resident(self, normalise: bool = True) -> SpecStats:
        st = SpecStats()
        self._check(self.lib.orcai_spectrogram_resident(self._h, int(normalise), C.byref(st)))
        return st

    def spectrogram(self, pcm: np.ndarray):
        """(normalised (T, n_freq) float32, stats) through the one-call entry point."""
        pcm, dt = self._pcm(pcm)
        T = self.lib.orcai_num_frames(pcm.size, self.params.hop)
        out = np.empty((T, self.params.n_freq), dtype=np.float32)
        st = SpecStats()
        self._check(self.lib.orcai_spectrogram(self._h, _ptr(pcm), dt, pcm.size, _ptr(out), C.byref(st)))
        return out, st

    def read_spectrogram(self, row0: int, nrows: int) -> np.ndarray:
        out = np.empty((nrows, self.params.n_freq), dtype=np.float32)
        self._check(self.lib.orcai_read_spectrogram(self._h, row0, nrows, _ptr(out)))
        return out

    def read_db(self, row0: int, nrows: int) -> np.ndarray:
        out = np.empty((nrows, self.params.n_freq), dtype=np.float32)
        self._check(self.lib.orcai_read_db(self._h, row0, nrows, _ptr(out)))
        return out

    # -- time chunks of one recording (orcai_b200/timesplit.py) -----------------------------------
    def chunk_spectrogram(self, stat_row0: int, stat_row1: int) -> float:
        """STFT -> dB of the uploaded chunk; max |S|^2 over the rows [stat_row0, stat_row1) the chunk owns."""
        m = C.c_float(0.0)
        self._check(self.lib.orcai_chunk_spectrogram(self._h, int(stat_row0), int(stat_row1), C.byref(m)))
        return float(m.value)

    def chunk_select_begin(self, max_power: float):
        self._check(self.lib.orcai_chunk_select_begin(self._h, float(max_power)))

    def chunk_histogram(self, pass_: int, row0: int, row1: int, prefix) -> np.ndarray:
        pre = np.asarray(prefix, dtype=np.uint32)
        hist = np.zeros((2, 2048), dtype=np.uint64)
        self._check(self.lib.orcai_chunk_histogram(self._h, int(pass_), int(row0), int(row1), _ptr(pre), _ptr(hist)))
        return hist

    def chunk_select_end(self, keys) -> SpecStats:
        k = np.asarray(keys, dtype=np.uint32)
        st = SpecStats()
        self._check(self.lib.orcai_chunk_select_end(self._h, _ptr(k), C.byref(st)))
        return st

    # -- network -------------------------------------------------------------------------------
    def forward_host(self, snippets: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(snippets, dtype=np.float32)
        if x.ndim == 4 and x.shape[-1] == 1:
            x = x[..., 0]
        if x.ndim != 3 or x.shape[1] != self.params.snippet_len or x.shape[2] != self.params.n_freq:
            raise ValueError(
                f"expected snippets of shape (N, {self.params.snippet_len}, {self.params.n_freq}[, 1]), got {snippets.shape}"
            )
        x = np.ascontiguousarray(x)
        out = np.empty((x.shape[0], self.pred_len, self.params.n_labels), dtype=np.float32)
        self._check(self.lib.orcai_forward_host(self._h, _ptr(x), x.shape[0], _ptr(out)))
        return out

    def forward_resident(self, first: int, n: int) -> np.ndarray:
        out = np.empty((n, self.pred_len, self.params.n_labels), dtype=np.float32)
        self._check(self.lib.orcai_forward_resident(self._h, first, n, _ptr(out)))
        return out

    # -- post-processing -----------------------------------------------------------------------
    def _segments_call(self, fn, S: int, L: int):
        cap = max(1024, min(L * ((S + 1) // 2), 1 << 18))
        while True:
            lab = np.empty(cap, np.int32)
            sta = np.empty(cap, np.int64)
            sto = np.empty(cap, np.int64)
            n = C.c_int64(0)
            rc = fn(_ptr(lab), _ptr(sta), _ptr(sto), cap, C.byref(n))
            if rc == ORCAI_ERR_CAPACITY and n.value > cap:
                cap = int(n.value)
                continue
            self._check(rc)
            k = int(n.value)
            return lab[:k], sta[:k], sto[:k]

    def predict_begin(self, n_samples: int, threshold: float = 0.5, want_agg: bool = True):
        """Enqueue the whole annotation of the resident recording (``swap_pcm`` / ``upload_pcm``) and return at once -> a token for
        ``predict_end``.  Up to two calls may be in flight; ``predict_end`` collects them oldest first (orcai_predict_resident_begin)."""
        T = self.lib.orcai_num_frames(int(n_samples), self.params.hop)
        ds = 1 << self.params.n_blocks
        S, L = T // ds, self.params.n_labels
        cap = max(1024, min(L * ((S + 1) // 2), 1 << 22))
        self._check(self.lib.orcai_predict_resident_begin(self._h, float(threshold), 1 if want_agg else 0, cap))
        return (S, L, cap, bool(want_agg))

    def predict_end(self, token):
        """Wait (sleeping) for the oldest begun call -> what ``predict_pcm`` returns."""
        S, L, cap, want_agg = token
        agg = np.empty((S, L), np.float64) if want_agg else None
        cnt = np.empty(S, np.float64) if want_agg else None
        st = SpecStats()
        lab = np.empty(cap, np.int32)
        sta = np.empty(cap, np.int64)
        sto = np.empty(cap, np.int64)
        n = C.c_int64(0)
        self._check(self.lib.orcai_predict_resident_end(self._h, C.byref(st), _ptr(agg), _ptr(cnt), _ptr(lab), _ptr(sta), _ptr(sto), cap, C.byref(n)))
        k = int(n.value)
        return st, agg, cnt, lab[:k].copy(), sta[:k].copy(), sto[:k].copy()

    def predict_in_flight(self) -> int:
        return int(self.lib.orcai_predict_in_flight(self._h))

    def postprocess(self, preds: np.ndarray, n_frames: int, threshold: float = 0.5, want_agg: bool = True):
        p = np.ascontiguousarray(preds, dtype=np.float32)
        N = p.shape[0]
        ds = 1 << self.params.n_blocks
        S, L = n_frames // ds, self.params.n_labels
        agg = np.zeros((S, L), np.float64) if want_agg else None
        cnt = np.zeros(S, np.float64) if want_agg else None
        lab, sta, sto = self._segments_call(
            lambda a, b, c, cap, n: self.lib.orcai_postprocess(
                self._h, _ptr(p), N, n_frames, float(threshold), _ptr(agg), _ptr(cnt), a, b, c, cap, n
            ),
            S,
            L,
        )
        return agg, cnt, lab, sta, sto

    def threshold_segments(self, agg: np.ndarray, cnt: np.ndarray, threshold: float = 0.5):
        a = np.ascontiguousarray(agg, dtype=np.float64)
        k = np.ascontiguousarray(cnt, dtype=np.float64)
        S, L = a.shape
        return self._segments_call(
            lambda x, y, z, cap, n: self.lib.orcai_threshold_segments(
                self._h, _ptr(a), _ptr(k), S, L, float(threshold), x, y, z, cap, n
            ),
            S,
            L,
        )

    # -- fused ---------------------------------------------------------------------------------
    def predict_pcm(self, pcm: np.ndarray, threshold: float = 0.5, want_agg: bool = True, resident: bool = False):
        """WAV samples -> (stats, agg, cnt, label_idx, start_step, stop_step) without leaving the device.

        resident=True reuses the recording already uploaded with upload_pcm (pcm is only used for its length).
        """
        pcm, dt = self._pcm(pcm)
        T = self.lib.orcai_num_frames(pcm.size, self.params.hop)
        ds = 1 << self.params.n_blocks
        S, L = T // ds, self.params.n_labels
        agg = np.zeros((S, L), np.float64) if want_agg else None
        cnt = np.zeros(S, np.float64) if want_agg else None
        st = SpecStats()
        if resident:
            call = lambda a, b, c, cap, n: self.lib.orcai_predict_resident(  # noqa: E731
                self._h, float(threshold), C.byref(st), _ptr(agg), _ptr(cnt), a, b, c, cap, n
            )
        else:
            call = lambda a, b, c, cap, n: self.lib.orcai_predict_pcm(  # noqa: E731
                self._h, _ptr(pcm), dt, pcm.size, float(threshold), C.byref(st), _ptr(agg), _ptr(cnt), a, b, c, cap, n
            )
        lab, sta, sto = self._segments_call(call, S, L)
        return st, agg, cnt, lab, sta, sto
