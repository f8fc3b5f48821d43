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


# network arithmetic: "fast" = fp16 tcgen05 fused residual blocks + tensor-core LSTM tail, biases calibrated against fp16
# weight rounding (probabilities: mean deviation 1e-4 from the fp32 graph, max ~2.7e-3 over an hour of audio; operand precision
# of TensorFlow's default TF32 execution on GPUs); "reference" = fp32 CUDA-core path (1e-6), 11x slower.
# "accurate" = "fast" with the fp32 CUDA-core entry convolution (fp32 spectrogram, weights and accumulation in the first layer):
# +5 % network time, mean deviation -29 % with BatchNorm-matched weights.
PRECISION_PATHS = {"fast": 3, "accurate": 3, "reference": 0}
PRECISION_CONV0 = {"fast": 1, "accurate": 0, "reference": 1}     # "conv0_path" option of the fused path


def precision_from_env() -> str:
    p = os.environ.get("ORCAI_B200_PRECISION", "fast").strip().lower()
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
        self.ctx = get_context(orcai_parameter, shape, device)
        self.ctx.load_weights(weights)
        self.precision = precision or precision_from_env()
        self.ctx.set_option("net_path", PRECISION_PATHS[self.precision])
        self.ctx.set_option("conv0_path", PRECISION_CONV0[self.precision])
        if PRECISION_PATHS[self.precision] == 3:
            # bias correction for fp16 weight rounding: channel means from a calibration recording - the built-in synthetic one,
            # or a representative recording of the deployment named by ORCAI_B200_CALIBRATION=<wav file>
            cal = os.environ.get("ORCAI_B200_CALIBRATION", "").strip()
            if cal:
                from orcai_b200.spectrogram import load_recording

                self.ctx.calibrate(load_recording(cal, 1, orcai_parameter["spectrogram"]))
            else:
                self.ctx.calibrate()
        n_blocks = len(orcai_parameter["model"]["filters"])
        self.input_shape = (None, *shape["input_shape"])
        self.output_shape = (None, shape["input_shape"][0] // 2**n_blocks, shape["num_labels"])

    def predict(self, x, batch_size: int | None = None, verbose: int = 0, **_unused) -> np.ndarray:
        """Keras-style predict; ``batch_size`` bounds the snippets staged on the device per call."""
        x = np.asarray(x, dtype=np.float32)
        if x.shape[0] == 0:
            raise ValueError("Expected input data to be non-empty.")
        step = int(batch_size) if batch_size else 1024
        step = max(step, 256)
        outs = [self.ctx.forward_host(x[i : i + step]) for i in range(0, x.shape[0], step)]
        return outs[0] if len(outs) == 1 else np.concatenate(outs, axis=0)

    __call__ = predict
