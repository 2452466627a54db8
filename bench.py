#!/usr/bin/env python
"""Headline benchmark: htdemucs audio-seconds separated per second on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference algorithm on the host CPU

Workload (BASELINE.json configs[2]): htdemucs (random-init, 4 stems, 41.98 M parameters), one
synthetic stereo 44.1 kHz track of 64 x 7.8 s segments PER GPU (overlap 0.25, shifts=0), run
through ``apply_model``: segments are batched through the kernel engine, overlap-added on the
device, and -- for N > 1 -- sharded in contiguous blocks across ranks with a halo exchange and a
final all-reduce of the stems (weak scaling: the track grows with N).  A "step" is one
``apply_model`` pass over the whole track.

Prints ONE JSON line (rank 0).  ``value`` = track seconds / device time with the track resident
in HBM; ``e2e`` = the same through the public API from pinned host memory and back.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEGMENTS_PER_GPU = 64
SEG_LEN = 343980            # int(39/5 * 44100)
STRIDE = 257985             # int(0.75 * SEG_LEN)
SR = 44100
STRICT_MODE = "tf32x3"
METRIC = "htdemucs audio-seconds separated per second"
UNIT = "audio-s/s"


def synth_track(length: int, seed: int = 1234):
    import torch
    g = torch.Generator().manual_seed(seed)
    return 0.1 * torch.randn(1, 2, length, generator=g)


def workload_config(n_gpus: int, mode: str, batch: int) -> dict:
    return {"workload": f"htdemucs 4-stem random-init, one track of {SEGMENTS_PER_GPU}x7.8s segments per GPU "
                        f"(BASELINE configs[2]), apply_model overlap=0.25 shifts=0",
            "segments_per_gpu": SEGMENTS_PER_GPU, "track_seconds": n_gpus * SEGMENTS_PER_GPU * STRIDE / SR,
            "forward_batch": batch, "mode": mode, "parallelism": f"segments sharded x{n_gpus}",
            "l2": "inputs larger than L2 (132 MB track, multi-GB activations per step)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        busy = [c for c in sm if c >= 0.5 * (mx[0] if mx else 1)] or sm
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(label: str):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) and tensor-pipe activity of the kernels
    behind a profile label, from the committed ncu pass over one 16-segment forward
    (profiles/r01_ncu_forward_b16_tf32.json, made by tools/ncu_forward_table.py).  None when absent."""
    path = os.path.join(ROOT, "profiles", "r01_ncu_forward_b16_tf32.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        table = json.load(f)["kernels"]
    if label.startswith("conv_gemm_tc<"):
        tbk, tbn = label[len("conv_gemm_tc<"):-1].split(",")     # label = kernel template <TBK,TBN>
        rows = [r for r in table if r["kernel"].replace(" ", "") == f"conv_gemm_tc_persist_kernel<{tbk},{tbn},0>"]
    else:
        key = {"attention_tc": "attention_tc_kernel<0>"}.get(label, label)
        rows = [r for r in table if r["kernel"].startswith(key)]
    if not rows:
        return None
    n = sum(r["launches"] for r in rows)
    ms = sum(r["ms"] for r in rows)
    return {"bytes_per_launch": sum(r["dram_mb"] for r in rows) * 1e6 / n, "launches": n,
            "tensor_pipe_active": sum(r["tensor_pipe_active"] * r["ms"] for r in rows) / ms,
            "source": "profiles/r01_ncu_forward_b16_tf32.json"}


def measured_peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops_sustained"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------ CPU arms
def cpu_apply_seconds(length: int, reps: int, threads: int, want_output: bool = False):
    """Time the reference algorithm's apply_model on the host CPU: the unmodified reference when
    its tree is present (build container), else the oracle port (GPU box)."""
    import torch
    from oracle import refload
    from demucs_b200.config import htdemucs_config
    from demucs_b200.weights import init_weights
    torch.set_num_threads(threads)
    cfg = htdemucs_config()
    W = init_weights(cfg, 0)
    mix = synth_track(length)
    if refload.available():
        ref = refload.load()
        model = refload.build_reference_model(cfg, W)
        kind = "reference"

        def run():
            return ref.apply_model(model, mix, shifts=0, split=True, overlap=0.25, device="cpu")
    else:
        from oracle.apply_oracle import apply_model_oracle
        kind = "port"

        def run():
            return apply_model_oracle((W, cfg), mix, shifts=0, split=True, overlap=0.25)
    times, out = [], None
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            out = run()
            times.append(time.perf_counter() - t0)
    if want_output:
        return times, kind, (mix, out)
    return times, kind


def reference_arm(args) -> None:
    """--impl reference: the reference's CPU implementation of the path on this host's cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    length = 10 * SR                                     # BASELINE configs[0]: 10 s clip, 2 segments
    times, kind = cpu_apply_seconds(length, args.warmup + args.steps, threads)
    timed = times[args.warmup:]
    sec = sum(timed) / len(timed)
    value = (length / SR) / sec
    sample = f"{len(timed)} x apply_model on a 10 s clip (2 segments), {threads} threads, torch CPU fp32"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus, "cpu-fp32", 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def gpu_arm(args) -> None:
    import torch
    import torch.distributed as dist
    import demucs_b200 as D
    from demucs_b200 import _lib, perf

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    shard = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        from demucs_b200.distributed import Shard
        shard = Shard()

    model = D.htdemucs(mode=args.mode).to(dev)
    eng = model.engine()
    # the error-compensated arithmetic (3xTF32: per-stem rel-L2 <= 1e-4, north_star's fp32/TF32 tolerance) is timed
    # beside the headline mode (single-pass TF32: <= 1e-2, north_star's reduced-precision tolerance)
    strict_model = None
    if args.mode != STRICT_MODE and not args.no_strict:
        strict_model = D.htdemucs(mode=STRICT_MODE).to(dev)
    nseg = SEGMENTS_PER_GPU * world
    length = nseg * STRIDE
    host_mix = synth_track(length).pin_memory()
    dev_mix = host_mix.to(dev)
    kw = dict(shifts=0, split=True, overlap=0.25, batch_size=args.batch, shard=shard)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return D.apply_model(model, dev_mix, device=dev, **kw)

    def step_strict():
        return D.apply_model(strict_model, dev_mix, device=dev, **kw)

    # End-to-end step: host -> device copy of the input this rank consumes, apply_model, device -> host read of
    # the sample range this rank produced (N = 1: the whole track both ways).  Over all ranks the reads cover the
    # result exactly once; every rank still holds the complete result on its device after the gather.
    if shard is None:
        in_lo, in_hi, out_lo, out_hi = 0, length, 0, length
    else:
        lo_seg, hi_seg = shard.block(nseg)
        first = max(0, lo_seg - shard.halo(SEG_LEN, STRIDE))
        in_lo, in_hi = first * STRIDE, min(length, (hi_seg - 1) * STRIDE + SEG_LEN)
        out_lo, out_hi = lo_seg * STRIDE, (length if hi_seg == nseg else hi_seg * STRIDE)
    host_out = torch.empty(1, 4, 2, out_hi - out_lo).pin_memory()
    e2e_mix = torch.zeros_like(dev_mix)
    h2d_bytes = torch.tensor([2 * (in_hi - in_lo) * 4], dtype=torch.float64, device=dev)
    d2h_bytes = torch.tensor([host_out.numel() * 4], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(h2d_bytes)
        dist.all_reduce(d2h_bytes)

    host_in = host_mix[..., in_lo:in_hi].contiguous().pin_memory()   # this rank's input, resident in pinned memory

    def step_e2e():
        for c in range(2):                       # row-wise: contiguous pinned <-> contiguous device runs, plain async copies
            e2e_mix[0, c, in_lo:in_hi].copy_(host_in[0, c], non_blocking=True)
        out = D.apply_model(model, e2e_mix, device=dev, **kw)
        for s_ in range(4):
            for c in range(2):
                host_out[0, s_, c].copy_(out[0, s_, c, out_lo:out_hi], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(host_out[0, 0, 0, 0])

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = eng.launches
    ms = timed(step_device, args.steps)
    launches = (eng.launches - l0) * world
    clocks = sampler.stop() if sampler else None
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    track_s = length / SR
    value, e2e = track_s / (ms / 1e3), track_s / (ms_e2e / 1e3)
    strict = None
    if strict_model is not None:
        for _ in range(2):
            step_strict()
        ms_strict = timed(step_strict, args.steps)
        strict = {"mode": STRICT_MODE, "value": track_s / (ms_strict / 1e3), "unit": UNIT, "ms_per_step": ms_strict,
                  "tolerance": "per-stem rel-L2 <= 1e-4"}

    # per-kernel device time + algorithmic work of ONE more step, with events around every launch
    prof = perf.profile_step(eng, step_device)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    top = prof["dominant"]
    tensor_peak = peaks["bf16_tflops"] / 2.0            # TF32 = 1/2 of the measured bf16 GEMM rate
    if top["bound"] == "tensor":
        achieved, peak, unit = top["tflops"], tensor_peak, "TFLOP/s"
    else:
        achieved, peak, unit = top["gbs"], peaks["hbm_gbs"], "GB/s"
    ncu = ncu_traffic(top["name"]) if args.mode == "tf32" else None
    if ncu:   # the ncu pass ran 16-segment forwards; activations (all but the L2-resident weights) scale with the batch
        ncu["captured_at_batch"] = 16
        ncu["bytes_per_launch"] *= args.batch / 16
    roofline = {"kernel": top["name"], "bound": top["bound"], "achieved": achieved, "peak": peak, "unit": unit,
                "frac": achieved / peak, "traffic": ncu["bytes_per_launch"] if ncu else None,
                "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)", "ncu": ncu,
                "algorithmic_bytes_per_launch": top.get("bytes_per_launch"), "peak_source": peaks["source"],
                "share_of_step": top["share"], "launches_per_step": top["count"],
                "avg_launch_ms": top["ms"] / max(top["count"], 1),
                "kernels": prof["table"]}
    cpu_baseline, parity = None, None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        times, kind, (cpu_mix, cpu_out) = cpu_apply_seconds(10 * SR, 3, threads, want_output=True)
        sec = sorted(times[1:])[len(times[1:]) // 2]
        # the CPU run doubles as the checker: same weights, same clip, through the same public call
        parity = {"against": kind, "clip": "10 s (2 segments), shifts=0, overlap=0.25", "per_stem_rel_l2_max": {},
                  "tolerance": {"tf32": 1e-2, STRICT_MODE: 1e-4, "fp32": 1e-4}}
        for name, mdl in ((args.mode, model), (STRICT_MODE, strict_model)):
            if mdl is None:
                continue
            got = D.apply_model(mdl, cpu_mix.to(dev), shifts=0, split=True, overlap=0.25, device=dev).cpu()
            err = max(float((got[0, s] - cpu_out[0, s]).norm() / cpu_out[0, s].norm()) for s in range(got.shape[1]))
            parity["per_stem_rel_l2_max"][name] = err
        cpu_baseline = {"value": 10.0 / sec, "unit": UNIT, "cores": threads, "kind": kind,
                        "sample": "median of 2 x apply_model on a 10 s clip (2 segments) after 1 warm-up, "
                                  f"{threads} threads, torch {torch.__version__} CPU fp32"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"tf32": "tf32", "tf32x3": "tf32x3 (3-pass, fp32-accurate)"}.get(args.mode, "f32"),
            "data": "synthetic",
            "config": workload_config(world, args.mode, args.batch),
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": int(d2h_bytes)},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "strict": strict, "parity": parity,
            "model_roofline": {"segments_per_s": nseg / (ms / 1e3),
                               "tf32_roofline_segments_per_s_per_gpu": 1e3 / 0.742,
                               "frac": (nseg / world / (ms / 1e3)) / (1e3 / 0.742)}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default=os.environ.get("BD_MODE", "tf32"), choices=["fp32", "tf32", "tf32x3"])
    ap.add_argument("--no-strict", action="store_true", help="skip timing the error-compensated mode beside the headline")
    ap.add_argument("--batch", type=int, default=64, help="segments per forward (64 = one forward per step and GPU; ~36 GB)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
