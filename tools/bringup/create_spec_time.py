import sys, time, json, tempfile
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from orcai_b200 import runtime, io as oio
from orcai_b200.synth import synth_pcm16
P, S = runtime.bundled_parameters()
ctx = runtime.get_context(P, S, 0)
pcm = synth_pcm16(3600.0, seed=1)
t0 = time.perf_counter(); spec, st = ctx.spectrogram(pcm); t1 = time.perf_counter()
spec, st = ctx.spectrogram(pcm); t2 = time.perf_counter()
with tempfile.TemporaryDirectory() as d:
    t3 = time.perf_counter(); oio.save_as_zarr(spec, Path(d) / "s.zarr"); t4 = time.perf_counter()
    size = sum(f.stat().st_size for f in (Path(d) / "s.zarr").rglob("*") if f.is_file())
    t5 = time.perf_counter(); back = oio.read_zarr(Path(d) / "s.zarr"); t6 = time.perf_counter()
print(json.dumps({"hours": 1.0, "spectrogram_call_s": t2 - t1, "first_call_s": t1 - t0, "zarr_write_s": t4 - t3, "zarr_bytes": size, "raw_bytes": spec.nbytes,
                  "zarr_read_s": t6 - t5, "roundtrip_equal": bool(np.array_equal(back, spec)), "device_ms": ctx.timings()["total_ms"]}))
