"""Process-wide liborcai_b200 contexts: one per (device, parameter set)."""

from __future__ import annotations

import json
import os
from importlib.resources import files

from orcai_b200._lib import Context

_contexts: dict = {}


def default_device() -> int:
    for key in ("ORCAI_B200_DEVICE", "LOCAL_RANK"):
        v = os.environ.get(key)
        if v is not None and v.strip().isdigit():
            return int(v)
    return 0


def bundled_parameters() -> tuple[dict, dict]:
    """Parameter dicts of the bundled orcai-V1 model directory."""
    d = files("orcai_b200.models").joinpath("orcai-V1")
    return json.loads(d.joinpath("orcai_parameter.json").read_text()), json.loads(d.joinpath("model_shape.json").read_text())


def shape_for(orcai_parameter: dict) -> dict:
    """model_shape implied by a parameter dict when only spectrograms are wanted (create-spectrograms)."""
    import numpy as np

    sp = orcai_parameter["spectrogram"]
    freqs = np.fft.rfftfreq(n=int(sp["nfft"]), d=1.0 / int(sp["sampling_rate"]))
    lo = int(np.argwhere(freqs <= sp["freq_range"][0])[0][0])
    hi = int(np.argwhere(freqs >= sp["freq_range"][1])[0][0])
    dt = int(sp["n_overlap"]) / int(sp["sampling_rate"])
    # snippet length rule of the reference (snippets.py:103-108): 16 * ((duration / dt) // 16)
    n = 2 ** len(orcai_parameter["model"]["filters"])
    snippet = int(n * ((sp["duration"] / dt) // n))
    return {"input_shape": [snippet, hi - lo, 1], "num_labels": len(orcai_parameter["calls"])}


def get_context(orcai_parameter: dict, shape: dict, device: int | None = None) -> Context:
    dev = default_device() if device is None else int(device)
    key = (dev, json.dumps(orcai_parameter["spectrogram"], sort_keys=True), json.dumps(orcai_parameter["model"], sort_keys=True),
           json.dumps(shape, sort_keys=True))
    ctx = _contexts.get(key)
    if ctx is None:
        ctx = Context(orcai_parameter, shape, device=dev)
        _contexts[key] = ctx
    return ctx


def default_context(device: int | None = None) -> Context:
    p, s = bundled_parameters()
    return get_context(p, s, device)


def close_all() -> None:
    for c in _contexts.values():
        c.close()
    _contexts.clear()
