import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from orcai_b200 import runtime
from orcai_b200.synth import synth_pcm16
from orcai_b200.weights import synthetic_weights
P, S = runtime.bundled_parameters()
ctx = runtime.get_context(P, S, 0)
ctx.load_weights(synthetic_weights(P, S, seed=1234))
ctx.set_option("net_path", 3); ctx.calibrate()
import bench
pcm = bench.make_recording(1.0, 20251018)
pinned = torch.from_numpy(pcm).pin_memory().numpy()
ctx.upload_pcm(pinned)
def timeit(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("resident want_agg=False", timeit(lambda: ctx.predict_pcm(pinned, want_agg=False, resident=True)))
print("resident want_agg=True ", timeit(lambda: ctx.predict_pcm(pinned, want_agg=True, resident=True)))
print("device total_ms", ctx.timings()["total_ms"])
def stream(n, agg):
    for _ in ctx.predict_stream((pinned for _ in range(n)), want_agg=agg): pass
for agg in (False, True):
    stream(2, agg); torch.cuda.synchronize(); t0 = time.perf_counter(); stream(6, agg); torch.cuda.synchronize()
    print("stream want_agg=%s" % agg, (time.perf_counter() - t0) / 6 * 1e3)
# pieces
t0 = time.perf_counter(); ctx.prefetch_pcm(pinned); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("prefetch call %.3f ms, copy done after %.3f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
t0 = time.perf_counter(); ctx.swap_pcm(); t1 = time.perf_counter(); print("swap call %.3f ms" % ((t1 - t0) * 1e3))
# network time while a copy is in flight
ctx.prefetch_pcm(pinned); r = ctx.predict_pcm(pinned, want_agg=False, resident=True); print("device total_ms with concurrent H2D", ctx.timings()["total_ms"], ctx.timings()["network_ms"])
torch.cuda.synchronize(); r = ctx.predict_pcm(pinned, want_agg=False, resident=True); print("device total_ms alone", ctx.timings()["total_ms"], ctx.timings()["network_ms"])
# device time per recording inside the streaming loop
tot = []
t0 = time.perf_counter()
for out in ctx.predict_stream((pinned for _ in range(8)), want_agg=True):
    tot.append((round(ctx.timings()["total_ms"], 3), round((time.perf_counter() - t0) * 1e3, 3)))
    t0 = time.perf_counter()
print("stream: (device total_ms, host ms per iteration)", tot)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for out in ctx.predict_stream((pinned for _ in range(8)), want_agg=True): pass
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
