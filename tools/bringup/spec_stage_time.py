import sys
sys.path.insert(0, '.')
from orcai_b200 import runtime
from orcai_b200.synth import synth_pcm16
P, S = runtime.bundled_parameters()
ctx = runtime.get_context(P, S, 0)
pcm = synth_pcm16(3600.0, seed=20251018)
ctx.upload_pcm(pcm)
acc = {}
for i in range(8):
    ctx.spectrogram_resident(True)
    if i >= 3:
        tm = ctx.timings()
        for k in ("stft_ms", "select_ms", "normalise_ms"):
            acc[k] = acc.get(k, 0.0) + tm[k] / 5
print(acc)
