"""CPU oracle for the orcAI prediction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker (or as the
timed CPU arm), never as a fallback for the CUDA path.

PARITY UNPINNED: the reference (ethz-tb/orcAI v1.0.3) ships no tests, no golden
vectors and no sample audio, and its numerics live in
third-party wheels that are not installable here (librosa 0.11.0, keras 3.10.0 /
tensorflow 2.19.0).  This oracle therefore restates the *published semantics* of
those libraries at the reference's call sites and is triangulated against
independent implementations available in this image (``torch.stft`` in float64,
``scipy.signal.stft``, ``torch.nn.LSTM``, closed-form signals); see
``tests/test_oracle_*.py`` and ``tools/make_golden.py``.
"""
