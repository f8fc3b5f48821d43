"""CPU oracle for the orcAI prediction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker (or as the
timed CPU arm), never as a fallback for the CUDA path.

PARITY PARTLY PINNED.  The reference (ethz-tb/orcAI v1.0.3) ships no tests, no golden vectors and no sample audio
(SURVEY.md section 4).  Its numpy / pandas code, however, runs here once the uninstallable third-party imports are stubbed:
``tools/make_reference_golden.py`` executes the reference's own ``preprocess_spectrogram``,
``compute_aggregated_predictions`` (snippet batcher + overlap-average, fake model), ``compute_binary_predictions``,
``find_consecutive_ones``, ``compute_labels``, ``filter_predictions`` and ``save_prediction_probabilities`` on seeded inputs
and freezes the outputs under ``tests/golden/reference_*``; ``tests/test_reference_golden.py`` holds this oracle (and the CUDA
post-processing) to them bit for bit.
UNPINNED remain the rows whose numerics live in third-party wheels that cannot be installed here: the STFT / dB conversion
(librosa 0.11.0: rows a2-a4), the network (keras 3.10.0 / tensorflow 2.19.0: row a7) and the label writer's float
formatting (pandas 2.2.3, raises under the installed pandas 3: row a12).  For those the oracle restates the *published
semantics* at the reference's call sites and is triangulated against independent implementations available in this image
(``torch.stft`` in float64, ``scipy.signal.stft``, ``torch.nn.LSTM``, closed-form signals); see ``tests/test_oracle_*.py``
and ``tools/make_golden.py``.
"""
