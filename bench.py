#!/usr/bin/env python
"""Benchmark of the orcAI prediction hot path on B200 (contract: see README / DESIGN.md section 7).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1      # CPU arm (restated reference path)

A "step" = `orcai predict` of one synthetic 1-hour 48 kHz mono recording per GPU (weak scaling: every rank
annotates its own recording; recordings shard naturally, there is no data-path collective).  Metric: hours of
audio annotated per second, whole job.

* `value`  : PCM already resident in HBM when the timed region starts (device-resident predict).
* `e2e`    : the public call with HOST buffers: pinned int16 PCM -> H2D -> predict -> segments/aggregates D2H.
* `roofline` : dominant stage (the orcai-V1 network) in TFLOP/s against the measured bf16 peak;
  `roofline_stft` : K1 (fused STFT->dB->crop) in algorithmic GB/s against the measured HBM copy peak.
* `cpu_baseline` : the CPU oracle (restated librosa + Keras path) timed on this box's host cores on a bounded sample.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "hours_of_audio_annotated_per_second"
UNIT = "h_audio/s"
STFT_BYTES_PER_FRAME_F32 = 1024 + 684  # 256 new float32 samples in + 171 float32 out
STFT_BYTES_PER_FRAME_I16 = 512 + 684
FLOP_PER_SNIPPET = 0.972e9  # 485.8 M MAC
# residual block 1 (the dominant kernel): SepConv 16->30 and 30->30 at 736x171 plus the 1x1/2 residual convolution
BLOCK1_MAC_PER_SNIPPET = 736 * 171 * (9 * 16 + 16 * 30) + 736 * 171 * (9 * 30 + 30 * 30) + 368 * 86 * 16 * 30
# what the folded implicit GEMM executes for it: 123 steps x 3 strips per snippet, 88 tcgen05.mma of 128 x 32 x 16 per step
BLOCK1_EXECUTED_MAC_PER_SNIPPET = 123 * 3 * 88 * 128 * 32 * 16
# dram__bytes_read + dram__bytes_write of the dominant kernel per snippet, from the LATEST committed ncu --set full capture of that
# kernel (profiles/; the file is named with the number): net_path -> {bytes_per_snippet, source}
BLOCK1_TRAFFIC = {
    3: {"bytes_per_snippet": (734.0e6 + 425.0e6) / 182, "source": "profiles/r01g_ncu_full_summary.csv (182-snippet launch)"},
    4: {"bytes_per_snippet": (7.654334e9 + 3.691655e9 + 0.321043e9 + 0.137265e9) / 1833,
        "source": "profiles/r02zv_ncu_full_summary.csv (tall launch over the 1 833 snippets of the 1-h recording + its border-image launch)"},
}
SELECT_PASSES = 3   # times the select streams the 704 B/frame dB buffer


def _peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.rows: list[list[str]] = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def make_recording(hours: float, seed: int) -> np.ndarray:
    from orcai_b200.synth import synth_pcm16

    return synth_pcm16(hours * 3600.0, seed=seed)


def cpu_oracle_hours_per_second(seconds: float, seed: int, threads: int) -> tuple[float, float]:
    """Restated reference CPU path (numpy STFT stage single-threaded like pocketfft, torch-CPU network, batch 32)."""
    import torch

    from oracle import network_oracle, postprocess_oracle as po, spectrogram_oracle as so
    from orcai_b200 import runtime
    from orcai_b200.synth import pcm16_to_float, synth_pcm16
    from orcai_b200.weights import synthetic_weights

    torch.set_num_threads(threads)
    P, S = runtime.bundled_parameters()
    W = synthetic_weights(P, S, seed=1234)
    pcm = synth_pcm16(seconds, seed=seed)
    t0 = time.perf_counter()
    y = pcm16_to_float(pcm)
    spec, _, times = so.make_spectrogram(y, P["spectrogram"])
    snips = po.cut_snippets(spec, 736)
    preds = np.concatenate([network_oracle.forward(snips[i : i + 32], W) for i in range(0, len(snips), 32)])
    agg, cnt = po.aggregate_predictions(preds, spec.shape[0], 736, 4, 7)
    s, e, n = po.binary_predictions(agg, cnt, P["calls"])
    po.labels_tsv(po.label_rows(s, e, n, 16, "*"), float(times[1] - times[0]))
    dt = time.perf_counter() - t0
    return (seconds / 3600.0) / dt, dt


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = len(os.sched_getaffinity(0))
    sample_s = args.cpu_sample_seconds
    for _ in range(args.warmup):
        cpu_oracle_hours_per_second(min(sample_s, 8.0), 20251018, threads)
    vals, times_ = [], []
    for k in range(args.steps):
        v, dt = cpu_oracle_hours_per_second(sample_s, 20251018 + k, threads)
        vals.append(v); times_.append(dt)
    value = float(np.mean(vals))
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(times_)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME, "note": "reference arm = CPU oracle (restated librosa+Keras path; the real reference is not installable offline)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample_s:.0f} s of the same synthetic audio per step (STFT stage numpy single-thread, network torch-CPU fp32 batch 32, synthetic weights seed 1234)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


WORKLOAD_NAME = "orcai predict (orcai-V1, synthetic weights) on one synthetic 1-hour 48 kHz mono PCM16 recording per GPU per step"
NET_PATH_NAMES = {
    0: "fp32 cuda-core", 1: "fp16 tcgen05 implicit-GEMM trunk + fp32 LSTM/dense", 2: "bf16 tcgen05 implicit-GEMM trunk + fp32 LSTM/dense",
    3: "fp16 tcgen05: pixel-group entry conv, fused residual-block kernels, tcgen05 LSTM projections + TMEM-resident recurrence",
    4: "split-fp16 tcgen05 (A_hi*W_hi + A_lo*W_hi + A_hi*W_lo, fp32 accumulate = fp32 grade): fused block 1, un-folded blocks 2-4 with the depthwise "
       "filter fused into the split pointwise GEMM, split-GEMM LSTM projections + fp32 recurrence; shared interior of overlapping snippets",
}
NET_DTYPE = {0: "f32", 1: "f16", 2: "bf16", 3: "f16", 4: "f16x3 (split fp16 operands, fp32 accumulate)"}
STAGES = ["conv0", "block1", "block2", "block3", "block4", "final_sep", "lstm1_proj", "lstm1_rec", "lstm2_proj", "lstm2_rec", "dense"]


def bench_dir() -> Path:
    d = Path(os.environ.get("ORCAI_BENCH_DIR", f"/tmp/orcai_b200_bench_{os.getuid()}"))
    d.mkdir(parents=True, exist_ok=True)
    return d


def prepare_table(workdir: Path, rank: int, world: int, hours: float, rows_per_rank: int, barrier) -> tuple[Path, Path, int]:
    """The recording table of the e2e arm: rows_per_rank * world rows over K <= 8 distinct seeded 1-h WAV files ON DISK,
    a model directory with the seeded synthetic weights; rank r writes the files r, r + world, .."""
    import pandas as pd

    from orcai_b200 import runtime
    from orcai_b200.wavio import write_wav_pcm16
    from orcai_b200.weights import save_npz, synthetic_weights

    n_files = min(8, max(2, world))
    for k in range(rank, n_files, world):
        f = workdir / f"bench_{hours:g}h_{k}.wav"
        if not f.exists() or f.stat().st_size < hours * 3600 * 96000:
            write_wav_pcm16(f, make_recording(hours, 20251018 + k))
    model_dir = workdir / "orcai-V1"
    if rank == 0:
        from importlib.resources import files as pkg_files

        model_dir.mkdir(exist_ok=True)
        src = pkg_files("orcai_b200.models").joinpath("orcai-V1")
        for name in ("orcai_parameter.json", "model_shape.json"):
            (model_dir / name).write_text(src.joinpath(name).read_text())
        P, S = runtime.bundled_parameters()
        if not (model_dir / "orcai-v1.weights.npz").exists():
            save_npz(synthetic_weights(P, S, seed=1234), model_dir / "orcai-v1.weights.npz")
        n_rows = rows_per_rank * world
        pd.DataFrame({"recording": [f"rec{j:04d}" for j in range(n_rows)], "channel": 1, "base_dir_recording": str(workdir),
                      "rel_recording_path": [f"bench_{hours:g}h_{j % n_files}.wav" for j in range(n_rows)]}).to_csv(workdir / f"table_{world}.csv", index=False)
    barrier()
    return workdir / f"table_{world}.csv", model_dir, n_files


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--hours", type=float, default=1.0, help="length of the synthetic recording each GPU annotates per step")
    ap.add_argument("--rows-per-gpu", type=int, default=32, help="e2e arm: recordings per GPU in the table one step annotates")
    ap.add_argument("--cpu-sample-seconds", type=float, default=600.0, help="bounded CPU-oracle sample (BASELINE config 0: one 10-min recording)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs 1, 2, 4 (create-spectrograms, 24-h recording, batch sweep)")
    ap.add_argument("--no-parity", action="store_true", help="skip the fp32-path comparison of the measured recording")
    ap.add_argument("--chunk", type=int, default=0, help="snippets per network chunk (0 = library default)")
    ap.add_argument("--net-path", type=int, default=4, choices=[0, 1, 2, 3, 4],
                    help="4 split-fp16 tcgen05, fp32 grade (the shipped default, inside the 1e-3 gate); 3 single-fp16 fused blocks (opt-in 'fast', outside the gate); "
                         "0 fp32 CUDA cores; 1 / 2 fp16 / bf16 tcgen05 layer-wise")
    ap.add_argument("--stft-f64", type=int, default=1, choices=[0, 1], help="1 float64 FFT (parity grade, default), 0 float32 FFT")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.warmup < 3:
        print(f"note: --warmup {args.warmup} < 3; timing hygiene asks for at least 3", file=sys.stderr)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: orcai_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # this process = ONE GPU's worker, whatever else is visible (a table run would otherwise shard over every visible GPU)
    os.environ["ORCAI_B200_DEVICE"] = str(local_rank)
    os.environ["ORCAI_B200_DEVICES"] = str(local_rank)
    precision = {0: "reference", 3: "fast", 4: "precise"}.get(args.net_path)
    if precision:
        os.environ["ORCAI_B200_PRECISION"] = precision

    from orcai_b200 import predict as opredict, runtime
    from orcai_b200.weights import synthetic_weights

    P, S = runtime.bundled_parameters()
    ctx = runtime.get_context(P, S, local_rank)

    def bind_bench_weights():
        ctx.load_weights(synthetic_weights(P, S, seed=1234))
        if args.chunk:
            ctx.set_option("chunk", args.chunk)
        ctx.set_option("net_path", args.net_path)
        ctx.set_option("stft_f64", args.stft_f64)
        if args.net_path in (1, 2, 3):
            ctx.calibrate()   # what OrcaiModel does for the fp16 paths (built-in synthetic calibration recording, not the bench file)

    bind_bench_weights()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    workdir = bench_dir()
    table_csv, model_dir, n_files = prepare_table(workdir, rank, world, args.hours, args.rows_per_gpu, barrier)

    # K <= 8 distinct seeded files; rank r keeps file r % K resident for the device arm
    from orcai_b200.wavio import read_wav

    pcm, _sr, _ch = read_wav(workdir / f"bench_{args.hours:g}h_{rank % n_files}.wav")
    pinned = torch.from_numpy(np.ascontiguousarray(pcm)).pin_memory()
    pcm_pinned = pinned.numpy()
    T = 1 + pcm.size // 256
    n_snip = (T - 736) // 368 + 1

    # ---------------- device-resident arm (`value`) ----------------
    ctx.upload_pcm(pcm_pinned)
    for _ in range(args.warmup):
        ctx.predict_pcm(pcm_pinned, want_agg=False, resident=True)
    launches0 = ctx.timings()["kernel_launches"]
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    stage = {k: 0.0 for k in ("stft_ms", "select_ms", "network_ms", "post_ms", "total_ms")}
    n_segments = 0
    net_stage = None
    for _ in range(args.steps):
        out = ctx.predict_pcm(pcm_pinned, want_agg=False, resident=True)  # synchronous: returns after the segments are on the host
        n_segments = len(out[3])
        tm = ctx.timings()
        for k in stage:
            stage[k] += tm[k]
        net_stage = tm["net_stage_ms"]
    barrier()
    t_res = max_over_ranks(time.perf_counter() - t0)
    launches = ctx.timings()["kernel_launches"] - launches0
    dev_ms = max_over_ranks(stage["total_ms"] / args.steps)

    # ---------------- end-to-end arm (`e2e`): the product's own table run, files -> label files ----------------
    # Every rank runs the SAME public call, `orcai_b200.predict.predict(TABLE.csv, output_path=DIR)` (what `orcai predict TABLE.csv -o DIR`
    # and `torchrun ... -m orcai_b200.cli predict ...` run): it takes its longest-first share of the rows (RANK / WORLD_SIZE), reads the
    # WAV files from disk into page-locked buffers on loader threads, uploads recording k+1 while the device annotates recording k,
    # builds the label tables and writes the label files.  Everything of a step is inside the timed region.
    outdir = workdir / f"out_{world}_{rank}"
    outdir.mkdir(exist_ok=True)
    silent = opredict.Messenger(verbosity=0)

    def table_step():
        opredict.predict(table_csv, model_dir=model_dir, output_path=str(outdir), overwrite=True, verbosity=0, msgr=silent)

    for _ in range(max(1, args.warmup // 2)):
        table_step()               # first call: model directory -> weights on the device; later calls reuse the cached model
    barrier()
    le0 = ctx.timings()["kernel_launches"]
    cpu0 = time.process_time()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        table_step()
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    host_cpu_s = max_over_ranks(time.process_time() - cpu0)   # user + system CPU time of this rank's process (all its threads)
    launches_e2e = ctx.timings()["kernel_launches"] - le0
    clocks = sampler.stop()   # sampled every 50 ms across both timed regions (and the short warm-up between them)
    rows_total = args.rows_per_gpu * world
    label_files = len(list(outdir.glob("*_predicted.txt")))
    bind_bench_weights()      # the table run bound the model directory's weights (the same seeded set) and options

    # ---------------- parity of the measured path (outside every timed region) ----------------
    parity = None
    if rank == 0 and args.net_path != 0 and not args.no_parity:
        ctx.upload_pcm(pcm_pinned)
        st_f, agg_f, cnt_f, lab_f, sta_f, sto_f = ctx.predict_pcm(pcm_pinned, want_agg=True, resident=True)
        raw_f = ctx.forward_resident(0, n_snip)
        ctx.set_option("net_path", 0)
        st_r, agg_r, cnt_r, lab_r, sta_r, sto_r = ctx.predict_pcm(pcm_pinned, want_agg=True, resident=True)
        raw_r = ctx.forward_resident(0, n_snip)
        ctx.set_option("net_path", args.net_path)
        seg_f = set(zip(lab_f.tolist(), sta_f.tolist(), sto_f.tolist()))
        seg_r = set(zip(lab_r.tolist(), sta_r.tolist(), sto_r.tolist()))
        mask_f, mask_r = agg_f > 0.25, agg_r > 0.25
        parity = {"against": "the library's fp32 CUDA-core network path on the same recording (that path is held to the CPU oracle at 1e-6 in tests/; "
                             "the oracle itself cannot annotate an hour within the bench)",
                  "gate": 1e-3,
                  "probability_max_abs_dev": float(np.abs(agg_f - agg_r).max()), "probability_mean_abs_dev": float(np.abs(agg_f - agg_r).mean()),
                  "per_snippet_probability_max_abs_dev": float(np.abs(raw_f - raw_r).max()),
                  "frames_with_different_mask_frac": float((mask_f != mask_r).mean()),
                  "segments": len(seg_f), "segments_reference_path": len(seg_r), "segments_identical": len(seg_f & seg_r),
                  "spectrogram_stats_equal": bool(st_f.lo == st_r.lo and st_f.hi == st_r.hi and st_f.db_ref == st_r.db_ref)}

    hours_total = args.hours * world * args.steps
    value = hours_total / t_res
    e2e_value = args.hours * rows_total * args.steps / t_e2e

    configs = None
    if rank == 0 and world == 1 and not args.no_configs:
        configs = run_configs(ctx, args, pcm_pinned, T, workdir, bind_bench_weights)

    if rank == 0:
        peaks = _peaks()
        stft_ms = stage["stft_ms"] / args.steps
        net_ms = stage["network_ms"] / args.steps
        stft_gbs = T * STFT_BYTES_PER_FRAME_I16 / (stft_ms * 1e-3) / 1e9
        net_tflops = n_snip * FLOP_PER_SNIPPET / (net_ms * 1e-3) / 1e12
        roofline_net = {"kernel": "orcai-V1 forward (all layer kernels)", "bound": "tensor", "achieved": net_tflops, "peak": peaks["bf16_tflops_sustained"],
                        "unit": "TFLOP/s", "frac": net_tflops / peaks["bf16_tflops_sustained"], "traffic": None,
                        "peak_source": peaks["source"] + " bf16 sustained", "flop_per_snippet": FLOP_PER_SNIPPET}
        if args.net_path in (3, 4) and net_stage[15] > 0:
            # dominant kernel = fused residual block 1; duration = CUDA events around its launch(es) (first chunk) on the compute stream
            b1_ms, b1_snips = float(net_stage[1]), float(net_stage[15])
            b1_tflops = 2.0 * BLOCK1_MAC_PER_SNIPPET * b1_snips / (b1_ms * 1e-3) / 1e12
            traffic = BLOCK1_TRAFFIC.get(args.net_path)
            roofline_main = {"kernel": "fused_block_kernel<block 1: sepconv 16->30, sepconv 30->30, maxpool(3,2)/2 + residual 1x1/2>"
                                       + (" PREC (split fp16), tall image + border images" if args.net_path == 4 else ""),
                             "bound": "tensor", "achieved": b1_tflops, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                             "frac": b1_tflops / peaks["bf16_tflops_sustained"],
                             "traffic": traffic["bytes_per_snippet"] * b1_snips if traffic else None, "traffic_source": traffic["source"] if traffic else None,
                             "peak_source": peaks["source"] + " bf16 sustained", "ms": b1_ms, "snippets": int(b1_snips),
                             "algorithmic_mac_per_snippet": BLOCK1_MAC_PER_SNIPPET,
                             "note": "achieved counts the reference graph's MACs of block 1 for every snippet of the chunk (the reference evaluates each "
                                     "snippet separately); the kernel folds the depthwise filter into the GEMM weights (9 taps x pointwise MACs)"
                                     + (", issues three MMAs per product (split fp16) and evaluates the shared interior of overlapping snippets once"
                                        if args.net_path == 4 else "") + "; its bound is the shared-memory operand bandwidth of the tensor pipe, see DESIGN.md"}
        else:
            roofline_main = roofline_net
        # read back per recording of the table run (no probabilities file asked for -> no aggregates): the segment count, the
        # statistics and the speculative copy of the first min(capacity, 32768) segments (orcai_predict_resident_begin)
        seg_cap = max(1024, min(7 * ((T // 16 + 1) // 2), 1 << 22))
        d2h_per_recording = 8 + 48 + 20 * min(seg_cap, 32768)
        ts = dict(opredict.TABLE_STATS)
        per = max(1, ts.get("rows", 0))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": NET_DTYPE[args.net_path], "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "hours_per_gpu_per_step": args.hours, "frames": T, "snippets": n_snip,
                       "segments_found": n_segments, "parallelism": f"shard-by-recording x{world}", "l2": "inputs larger than L2 (346 MB PCM, 475 MB dB per step)",
                       "network_path": NET_PATH_NAMES[args.net_path], "stft": "float64 FFT" if args.stft_f64 else "float32 FFT",
                       "cpu_arm_note": "the CPU arms (cpu_baseline, --impl reference) time a bounded 600-s sample of the same synthetic audio and report the same rate"},
            "device_ms_per_step": dev_ms,
            "stage_ms": {k: v / args.steps for k, v in stage.items()},
            "net_stage_ms_first_chunk": dict(zip(STAGES, [round(v, 4) for v in net_stage[:11]])) | {"snippets": int(net_stage[15])},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(pcm.nbytes) * rows_total,
                    "d2h_bytes_per_step": int(d2h_per_recording) * rows_total, "ms_per_step": 1e3 * t_e2e / args.steps,
                    "what": f"orcai_b200.predict.predict(TABLE.csv, output_path=DIR): {rows_total} rows of 1-h WAV files on disk ({n_files} distinct files) -> "
                            f"{rows_total} label files per step; every rank = one GPU's worker process taking its longest-first share (RANK / WORLD_SIZE)",
                    "recordings_per_step": rows_total, "label_files_written_rank0": label_files,
                    "host_cpu_ms_per_recording": 1e3 * host_cpu_s / (args.rows_per_gpu * args.steps), "host_cores": len(os.sched_getaffinity(0)),
                    "main_thread_ms_per_recording_rank0": {"waiting_for_the_device": 1e3 * ts.get("wait_device_s", 0.0) / per,
                                                           "waiting_for_a_loader_thread": 1e3 * ts.get("wait_loader_s", 0.0) / per,
                                                           "file_read_on_a_loader_thread": 1e3 * ts.get("load_s", 0.0) / per,
                                                           "waiting_for_the_writer_at_table_end_ms": 1e3 * ts.get("wait_writer_s", 0.0)},
                    "host_note": "CPU time of one rank's process per recording it annotates (max over ranks), mostly the read of the 346 MB file into the "
                                 "page-locked buffer; N ranks need N x that per device-time-per-recording of cores - beyond the host's core count the "
                                 "e2e arm is host-bound, the device-resident `value` is not",
                    "limiter": "per recording: WAV read + decode on 3 loader threads, one H2D copy of 346 MB overlapped with the previous recording's kernels, "
                               "device time (recording k+1 is enqueued before recording k is collected: orcai_predict_resident_begin / _end), label table + "
                               "label file on a writer thread; the slowest of these per recording bounds the rate"},
            "gpu_launches": int(launches), "gpu_launches_e2e": int(launches_e2e),
            "parity": parity,
            "clocks": clocks,
            "roofline": roofline_main,
            "roofline_network": roofline_net,
            "roofline_stft": {"kernel": "stft_db_kernel<int16>", "bound": "hbm", "achieved": stft_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                              "frac": stft_gbs / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"] + " copy",
                              "bytes_per_frame": STFT_BYTES_PER_FRAME_I16, "ms": stft_ms},
        }
        if configs:
            line["create_spectrograms"] = configs.pop("create_spectrograms")
            line["configs"] = configs
        if world == 1 and not args.no_cpu_baseline:
            threads = len(os.sched_getaffinity(0))
            v, dt = cpu_oracle_hours_per_second(args.cpu_sample_seconds, 20251018, threads)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{args.cpu_sample_seconds:.0f} s of the same synthetic audio ({dt:.1f} s of CPU work; numpy STFT stage + torch-CPU fp32 network batch 32)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_configs(ctx, args, pcm_pinned, T, workdir, rebind) -> dict:
    """BASELINE configs beyond the headline, one GPU, outside the headline's timed regions:
    configs[1] create-spectrograms (device stage + files -> zarr wall time), configs[2] one 24-h recording, configs[4] batch sweep."""
    import torch

    from orcai_b200 import spectrogram as ospec
    from orcai_b200.auxiliary import Messenger

    peaks = _peaks()
    hbm = peaks["hbm_gbs"]
    out = {}
    # ---- configs[1]: STFT -> dB -> crop (+ global max), exact percentile select, clip + normalise into the compact (T, 171) array ----
    cs = {k: 0.0 for k in ("stft_ms", "select_ms", "normalise_ms", "total_ms")}
    ctx.upload_pcm(pcm_pinned)
    reps = max(3, args.steps)
    for i in range(2 + reps):
        ctx.spectrogram_resident(normalise=True)
        if i >= 2:
            tm = ctx.timings()
            for k in cs:
                cs[k] += tm[k] / reps
    ctx.set_option("stft_f64", 0)
    for _ in range(3):
        ctx.spectrogram_resident(normalise=True)
    stft_f32_ms = ctx.timings()["stft_ms"]
    ctx.set_option("stft_f64", args.stft_f64)
    ctx.set_option("stft_threads", 8)          # the 8-threads-per-frame float64 kernel of round 1, for comparison
    for _ in range(3):
        ctx.spectrogram_resident(normalise=True)
    stft_t8_ms = ctx.timings()["stft_ms"]
    ctx.set_option("stft_threads", 16)
    # the application: one-row recording table -> OUTDIR/<recording>/spectrogram/{spectrogram.zarr, times.json, frequencies.json}
    import pandas as pd
    import shutil

    tab = workdir / "spec_table.csv"
    P, _S = __import__("orcai_b200.runtime", fromlist=["x"]).bundled_parameters()
    row = {"recording": "spec0", "channel": 1, "base_dir_recording": str(workdir), "rel_recording_path": f"bench_{args.hours:g}h_0.wav", "base_dir_annotation": "x"}
    row.update({c: True for c in P["calls"]})
    pd.DataFrame([row]).to_csv(tab, index=False)
    walls = []
    for _ in range(3):
        shutil.rmtree(workdir / "spec_out", ignore_errors=True)
        t0 = time.perf_counter()
        ospec.create_spectrograms(tab, workdir / "spec_out", orcai_parameter=P, verbosity=0, msgr=Messenger(verbosity=0))
        walls.append(time.perf_counter() - t0)
    zarr_bytes = sum(f.stat().st_size for f in (workdir / "spec_out").rglob("*") if f.is_file())
    shutil.rmtree(workdir / "spec_out", ignore_errors=True)
    from orcai_b200 import io as oio

    out["create_spectrograms"] = {
        "workload": "BASELINE configs[1]: create-spectrograms on the same 1-h recording: device stage (STFT/dB/crop + exact percentiles + clip/normalise) and the application (WAV file -> zarr store)",
        "stage_ms": {k: round(v, 4) for k, v in cs.items()},
        "hours_per_second_device": args.hours / (cs["total_ms"] * 1e-3),
        "compulsory_bytes_per_frame": STFT_BYTES_PER_FRAME_I16,
        "whole_stage": {"achieved": T * STFT_BYTES_PER_FRAME_I16 / (cs["total_ms"] * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                        "frac": T * STFT_BYTES_PER_FRAME_I16 / (cs["total_ms"] * 1e-3) / 1e9 / hbm,
                        "note": "compulsory traffic only (int16 samples in, 171 float32 out); the stage also streams the 704 B/frame dB buffer for the exact select and once more for the normalise"},
        "normalise_kernel": {"bytes_per_frame": 1368, "achieved": T * 1368 / (cs["normalise_ms"] * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                             "frac": T * 1368 / (cs["normalise_ms"] * 1e-3) / 1e9 / hbm},
        "select_passes": {"bytes_per_frame": SELECT_PASSES * 704, "achieved": T * SELECT_PASSES * 704 / (cs["select_ms"] * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                          "frac": T * SELECT_PASSES * 704 / (cs["select_ms"] * 1e-3) / 1e9 / hbm},
        "stft_float32_fft_variant": {"ms": stft_f32_ms, "achieved": T * STFT_BYTES_PER_FRAME_I16 / (stft_f32_ms * 1e-3) / 1e9, "unit": "GB/s",
                                     "frac": T * STFT_BYTES_PER_FRAME_I16 / (stft_f32_ms * 1e-3) / 1e9 / hbm,
                                     "note": "max error 1e-3 dB against the float64 oracle: at the gate, hence not the default"},
        "stft_float64_8_threads_per_frame": {"ms": stft_t8_ms, "note": "the round-1 decomposition (32 x 8, 255 registers, 8 warps per SM); the default is 16 x 16 at 128 registers"},
        "application": {"what": "orcai_b200.spectrogram.create_spectrograms(TABLE.csv, OUTDIR): 1-h WAV on disk -> spectrogram.zarr (zarr v3, chunks (2000, 171), gzip) + times.json + frequencies.json",
                        "wall_s_best_of_3": min(walls), "hours_per_second": args.hours / min(walls), "gzip_threads": oio.gzip_workers(),
                        "store_bytes": zarr_bytes, "limiter": "host: gzip of 462 MB per hour of audio on the stated threads, then file writes"},
    }
    # ---- configs[2]: ONE 24-h recording (the 1-h recording repeated 24 times: 4.15 G samples, 8.3 GB of PCM16), long sliding-window path ----
    try:
        long_pcm = np.tile(np.asarray(pcm_pinned), 24)
        pin24 = torch.from_numpy(long_pcm).pin_memory()
        p24 = pin24.numpy()
        del long_pcm
        ctx.predict_pcm(p24, want_agg=True)                      # warm-up: buffers grow to the 24-h size
        t0 = time.perf_counter()
        r = ctx.predict_pcm(p24, want_agg=True)                  # host buffer -> H2D -> predict -> aggregates + segments on the host
        wall = time.perf_counter() - t0
        tm = ctx.timings()
        out["one_24h_recording"] = {"workload": "BASELINE configs[2]: orcai predict on ONE 24-hour recording, 1 B200 (the 1-h recording repeated 24 times)",
                                    "frames": int(r[0].n_frames), "snippets": int((r[0].n_frames - 736) // 368 + 1), "segments": int(len(r[3])),
                                    "device_ms": tm["total_ms"], "network_ms": tm["network_ms"], "e2e_ms": 1e3 * wall, "h2d_bytes": int(p24.nbytes),
                                    "hours_per_second_device": 24.0 / (tm["total_ms"] * 1e-3), "hours_per_second_e2e": 24.0 / wall}
        # ---- configs[4]: snippet forward, batch sweep 1 - 4096 on the resident 24-h recording ----
        ctx.upload_pcm(p24)
        ctx.spectrogram_resident(False)
        ctx.set_option("chunk", 4096)
        sizes = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096]
        ref, rows = {}, []
        for path, name in ((0, "fp32 CUDA cores"), (4, "split-fp16 tcgen05 (default)"), (3, "single-fp16 tcgen05 fused, calibrated (opt-in)"), (2, "bf16 tcgen05 layer-wise")):
            if path == 3:
                ctx.calibrate()                                   # replaces the resident recording
                ctx.upload_pcm(p24)
                ctx.spectrogram_resident(False)
            ctx.set_option("net_path", path)
            for n in sizes:
                if path in (0, 2) and n > 512:
                    continue                                      # the slow paths only serve as tolerance references
                res = ctx.forward_resident(0, n)                  # warm-up + result
                best = 1e30
                for _ in range(3):
                    ctx.forward_resident(0, n)
                    best = min(best, ctx.timings()["network_ms"])
                if path == 0:
                    ref[n] = res
                rows.append({"path": name, "batch": n, "ms": round(best, 4), "snippets_per_s": n / (best * 1e-3), "tflops": n * FLOP_PER_SNIPPET / (best * 1e-3) / 1e12,
                             "max_abs_dev_vs_fp32": float(np.abs(res - ref[n]).max()) if n in ref else None})
        out["batch_sweep"] = {"workload": "BASELINE configs[4]: orcai-V1 snippet forward, batch 1 - 4096 snippets of a resident recording; deviation of the per-snippet "
                                          "probabilities from the fp32 path (gate 1e-3); tensor-pipe use = tflops / measured bf16 peak", "rows": rows}
        del pin24, p24
    except Exception as e:  # noqa: BLE001 - the headline must not die with an optional config
        out["configs_error"] = f"{type(e).__name__}: {e}"
    rebind()
    ctx.set_option("chunk", args.chunk if args.chunk else 1024)
    return out


if __name__ == "__main__":
    main()
