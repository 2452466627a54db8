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
    for name, e0, e1, flops, nbytes in records:
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
                      "bound": "tensor" if intensity >= TF32_BALANCE else "hbm"})
    return {"table": table, "dominant": table[0], "total_ms": total}
