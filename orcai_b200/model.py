"""The model object boundary: ``model.predict(x (N,736,171,1) f32) -> (N,46,7) f32``.

Stands in for the ``keras.Model`` that the reference's ``load_orcai_model`` returns
(``src/orcAI/io.py:357-410``) and that ``compute_aggregated_predictions`` calls
(``src/orcAI/predict.py:266-268``).  The forward pass runs in liborcai_b200.
"""

from __future__ import annotations

import os

import numpy as np

from orcai_b200.runtime import get_context
from orcai_b200.weights import check_weights


# network arithmetic (ORCAI_B200_PRECISION):
#   "precise"   (default) every GEMM on the fp16 tensor cores as the three-term split A_hi*W_hi + A_lo*W_hi + A_hi*W_lo, fp32
#               accumulation, fp32 CUDA cores for the entry convolution / depthwise filters / recurrence (net_path 4): probabilities
#               within 2e-5 of the fp32 graph (gate of the reference comparison: 1e-3), resident recordings evaluated with the shared
#               interior of overlapping snippets.
#   "reference" fp32 CUDA-core path (net_path 0, 1e-6): the in-library yardstick the others are held to.
#   "fast"      single fp16 operands with calibrated biases (net_path 3): 2.7x the speed of "precise" but up to 2.7e-3 from the fp32
#               graph over an hour of audio - OUTSIDE the 1e-3 gate, opt-in only; "accurate" = "fast" + fp32 entry convolution.
PRECISION_PATHS = {"precise": 4, "fast": 3, "accurate": 3, "reference": 0}
PRECISION_CONV0 = {"precise": 1, "fast": 1, "accurate": 0, "reference": 1}     # "conv0_path" option of the fp16 fused path
DEFAULT_PRECISION = "precise"


def precision_from_env() -> str:
    p = os.environ.get("ORCAI_B200_PRECISION", DEFAULT_PRECISION).strip().lower()
    if p not in PRECISION_PATHS:
        raise ValueError(f"ORCAI_B200_PRECISION must be one of {sorted(PRECISION_PATHS)}, got {p!r}")
    return p


class OrcaiModel:
    def __init__(self, orcai_parameter: dict, shape: dict, weights: dict, device: int | None = None, precision: str | None = None):
        if orcai_parameter.get("architecture", "ResNetLSTM") != "ResNetLSTM":
            raise ValueError(f"Unknown model architecture: {orcai_parameter.get('architecture')}")
        check_weights(weights, orcai_parameter, shape)
        self.orcai_parameter = orcai_parameter
        self.shape = shape
        self.weights = weights
        self.precision = precision or precision_from_env()
        if self.precision not in PRECISION_PATHS:
            raise ValueError(f"precision must be one of {sorted(PRECISION_PATHS)}, got {self.precision!r}")
        # contexts are shared per (device, parameters): the device weights / options / calibration belong to whichever model
        # bound them last, and every entry point re-binds first (bind()), so two models never predict with each other's weights
        self.ctx = get_context(orcai_parameter, shape, device)
        self._token = object()
        self.bind()
        n_blocks = len(orcai_parameter["model"]["filters"])
        self.input_shape = (None, *shape["input_shape"])
        self.output_shape = (None, shape["input_shape"][0] // 2**n_blocks, shape["num_labels"])

    def bind(self):
        """Make this model the owner of its context: weights, arithmetic options and (fast path) bias calibration."""
        ctx = self.ctx
        if getattr(ctx, "owner", None) is self._token:
            return ctx
        ctx.load_weights(self.weights)
        ctx.set_option("net_path", PRECISION_PATHS[self.precision])
        ctx.set_option("conv0_path", PRECISION_CONV0[self.precision])
        self.calibration = None
        if PRECISION_PATHS[self.precision] == 3:
            # bias correction for fp16 weight rounding: channel means from a calibration recording - the built-in synthetic one,
            # or a representative recording of the deployment named by ORCAI_B200_CALIBRATION=<wav file>
            cal = os.environ.get("ORCAI_B200_CALIBRATION", "").strip()
            if cal:
                from orcai_b200.spectrogram import load_recording

                ctx.calibrate(load_recording(cal, 1, self.orcai_parameter["spectrogram"]))
            else:
                ctx.calibrate()
            self.calibration = cal or "built-in synthetic recording"
        ctx.owner = self._token
        return ctx

    def describe(self) -> str:
        """One line for the log: which arithmetic produces the label files."""
        text = f"network arithmetic: {self.precision} (net_path {PRECISION_PATHS[self.precision]})"
        if self.calibration:
            text += f", bias calibration on {self.calibration}"
        return text

    def predict(self, x, batch_size: int | None = None, verbose: int = 0, **_unused) -> np.ndarray:
        """Keras-style predict; ``batch_size`` bounds the snippets staged on the device per call."""
        self.bind()
        x = np.asarray(x, dtype=np.float32)
        if x.shape[0] == 0:
            raise ValueError("Expected input data to be non-empty.")
        step = int(batch_size) if batch_size else 1024
        step = max(step, 256)
        outs = [self.ctx.forward_host(x[i : i + step]) for i in range(0, x.shape[0], step)]
        return outs[0] if len(outs) == 1 else np.concatenate(outs, axis=0)

    __call__ = predict
