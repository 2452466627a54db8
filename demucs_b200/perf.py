"""Per-kernel device timing of one step with CUDA events on the launching stream.

``Engine._k`` brackets every launch with a pair of events while a profile is active and carries
the launch's ALGORITHMIC flops / bytes (what the layer must move and multiply, not what the
kernel happened to do).  Used by bench.py for the ``roofline`` object and by tools/ scripts.
"""
from __future__ import annotations

import typing as tp

import torch

TF32_BALANCE = 100.0   # flop/byte above which a launch is tensor-bound (671 TF/s / 6.5 TB/s)


def profile_step(engine, fn: tp.Callable[[], tp.Any]) -> dict:
    engine._prof = []
    try:
        torch.cuda.synchronize()
        fn()
        torch.cuda.synchronize()
        records = engine._prof
    finally:
        engine._prof = None
    groups: tp.Dict[str, dict] = {}
    launches = []
    for name, e0, e1, flops, nbytes, detail in records:
        launches.append({"name": name, "detail": detail, "ms": e0.elapsed_time(e1), "flops": flops, "bytes": nbytes})
        g = groups.setdefault(name, {"name": name, "ms": 0.0, "count": 0, "flops": 0.0, "bytes": 0.0})
        g["ms"] += e0.elapsed_time(e1)
        g["count"] += 1
        g["flops"] += flops
        g["bytes"] += nbytes
    total = sum(g["ms"] for g in groups.values()) or 1.0
    table = []
    for g in sorted(groups.values(), key=lambda g: -g["ms"]):
        sec = g["ms"] / 1e3 or 1e-12
        intensity = g["flops"] / g["bytes"] if g["bytes"] else float("inf")
        table.append({"name": g["name"], "count": g["count"], "ms": round(g["ms"], 3),
                      "share": round(g["ms"] / total, 4), "tflops": round(g["flops"] / sec / 1e12, 3),
                      "gbs": round(g["bytes"] / sec / 1e9, 1),
                      "bytes_per_launch": g["bytes"] / g["count"], "flops_per_launch": g["flops"] / g["count"],
                      "bound": "tensor" if intensity >= TF32_BALANCE else "hbm"})
    return {"table": table, "dominant": table[0], "total_ms": total, "launches": launches}


def main():
    """python -m demucs_b200.perf [--batch B] [--mode fp32|tf32]: per-launch times of one forward."""
    import argparse
    from .htdemucs import htdemucs
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--mode", default="strict")
    ap.add_argument("--top", type=int, default=40)
    args = ap.parse_args()
    model = htdemucs(mode=args.mode).to("cuda")
    eng = model.engine()
    mix = 0.1 * torch.randn(args.batch, 2, eng.cfg.segment_length, device="cuda")
    for _ in range(2):
        eng.forward(mix)
    prof = profile_step(eng, lambda: eng.forward(mix))
    print(f"forward batch={args.batch} mode={args.mode}: {prof['total_ms']:.2f} ms in kernels, "
          f"{len(prof['launches'])} launches")
    for row in prof["table"]:
        print(f"  {row['name']:<26} n={row['count']:<4} {row['ms']:9.3f} ms {100 * row['share']:5.1f}%  "
              f"{row['tflops']:8.2f} TF/s {row['gbs']:8.1f} GB/s")
    merged = {}
    for l in prof["launches"]:
        k = (l["name"], l["detail"])
        m = merged.setdefault(k, {"ms": 0.0, "n": 0, "flops": 0.0, "bytes": 0.0})
        m["ms"] += l["ms"]; m["n"] += 1; m["flops"] += l["flops"]; m["bytes"] += l["bytes"]
    print("top launches:")
    for (name, detail), m in sorted(merged.items(), key=lambda kv: -kv[1]["ms"])[: args.top]:
        sec = m["ms"] / 1e3
        print(f"  {m['ms']:8.3f} ms x{m['n']:<3} {name:<22} {detail:<58} "
              f"{m['flops'] / sec / 1e12:7.2f} TF/s {m['bytes'] / sec / 1e9:7.1f} GB/s")


if __name__ == "__main__":
    main()
