import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from orcai_b200 import runtime
from orcai_b200.synth import synth_pcm16
from orcai_b200.weights import synthetic_weights
P, S = runtime.bundled_parameters()
ctx = runtime.get_context(P, S, 0)
ctx.load_weights(synthetic_weights(P, S, seed=1234))
import os
ctx.set_option("net_path", int(os.environ.get("ORCAI_TRACE_NET_PATH", "3")))
pcm = synth_pcm16(600.0, seed=20251018)
for _ in range(3):
    ctx.predict_pcm(pcm)
t = ctx.timings()
print([round(v, 3) for v in t["net_stage_ms"][:6]])
