"""orcai_b200 - B200-native implementation of the orcAI prediction hot path.

WAV -> STFT/dB/crop/normalise -> sliding-window snippets -> orcai-V1 forward -> overlap-average ->
threshold -> labelled segments, as hand-written sm_100a CUDA kernels behind a C ABI
(include/orcai_b200.h), with the reference's Python function surface on top
(``orcai_b200.spectrogram``, ``orcai_b200.predict``, ``orcai_b200.cli``).
"""

__version__ = "0.1.0"
