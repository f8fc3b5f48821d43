"""GPU parity: snippet batcher + orcai-V1 forward through the C ABI vs the torch-CPU oracle."""

import numpy as np
import pytest

from oracle import network_oracle, postprocess_oracle as po, spectrogram_oracle as so
from orcai_b200.synth import pcm16_to_float, synth_pcm16
from orcai_b200.weights import synthetic_weights

pytestmark = pytest.mark.gpu

PROB_TOL = 1e-3  # north_star: per-frame probabilities within 1e-3 absolute


def test_golden_probabilities(ctx, golden_dir):
    g = np.load(golden_dir / "network_seed1234.npz")
    x = np.random.default_rng(5).random((2, 736, 171), dtype=np.float32)
    out = ctx.forward_host(x)
    assert out.shape == (2, 46, 7) and out.dtype == np.float32
    assert np.abs(out - g["probs"]).max() <= PROB_TOL


def test_model_predict_boundary(ctx, params):
    """model.predict((N,736,171,1)) -> (N,46,7), Keras-style, incl. batches that do not divide the chunk."""
    from orcai_b200.model import OrcaiModel

    P, S = params
    W = synthetic_weights(P, S, seed=1234)
    model = OrcaiModel(P, S, W, device=0)
    x = np.random.default_rng(6).random((5, 736, 171, 1), dtype=np.float32)
    ref = network_oracle.forward(x, W)
    model.ctx.set_option("chunk", 2)  # 5 snippets in chunks of 2 -> ragged last chunk
    out = model.predict(x, verbose=0)
    model.ctx.set_option("chunk", 128)
    assert out.shape == (5, 46, 7) and np.abs(out - ref).max() <= PROB_TOL
    one = model.predict(x[:1])
    np.testing.assert_array_equal(one, out[:1])  # a snippet's result does not depend on its batch
    with pytest.raises(ValueError):
        model.predict(np.zeros((0, 736, 171, 1), np.float32))
    with pytest.raises(ValueError):
        model.ctx.forward_host(np.zeros((1, 700, 171), np.float32))


def test_strided_snippet_batcher_matches_copies(ctx, params):
    """Snippets cut as strided windows of the resident recording == the reference's materialised copies."""
    P, S = params
    pcm = synth_pcm16(14.0, seed=31)
    spec, st = ctx.spectrogram(pcm)
    n = int((st.n_frames - 736) // 368 + 1)
    assert n == 6
    resident = ctx.forward_resident(0, n)           # reads the raw dB buffer, normalises on load
    tail = ctx.forward_resident(2, 3)
    copies = ctx.forward_host(po.cut_snippets(spec, 736))  # (N,736,171) materialised from the device's own spectrogram
    np.testing.assert_array_equal(resident, copies)
    np.testing.assert_array_equal(tail, resident[2:5])
    # against the oracle end to end (oracle spectrogram -> oracle network)
    db, f, _ = so.calculate_spectrogram(pcm16_to_float(pcm), P["spectrogram"])
    spec_ref, _, _ = so.preprocess_spectrogram(db, f, P["spectrogram"])
    ref = network_oracle.forward(po.cut_snippets(spec_ref, 736), synthetic_weights(P, S, seed=1234))
    assert np.abs(resident - ref).max() <= PROB_TOL
    with pytest.raises(Exception):
        ctx.forward_resident(4, 3)  # past the last snippet


def test_other_weights_seed(ctx, params):
    P, S = params
    W = synthetic_weights(P, S, seed=99)
    ctx.load_weights(W)
    try:
        x = np.random.default_rng(8).random((3, 736, 171), dtype=np.float32)
        assert np.abs(ctx.forward_host(x) - network_oracle.forward(x, W)).max() <= PROB_TOL
        bad = dict(W)
        bad.pop("dense2/bias")
        with pytest.raises(Exception, match="dense2/bias"):
            ctx.load_weights(bad)
    finally:
        ctx.load_weights(synthetic_weights(P, S, seed=1234))
