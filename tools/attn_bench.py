"""Micro-benchmark of bd_attention on the transformer shapes (CUDA events)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demucs_b200 import _lib  # noqa: E402

_lib.build()
DEV = "cuda:0"
MODES = sys.argv[1].split(",") if len(sys.argv) > 1 else ["tf32", "tf32x3"]
B, H, D = 16, 8, 512
for mname in MODES:
    math_mode = {"tf32": _lib.MATH_TF32, "tf32x3": _lib.MATH_TF32X3, "fp32": _lib.MATH_FP32}[mname]
    for Tq, Tk in ((2688, 2688), (1344, 1344), (2688, 1344), (1344, 2688)):
        q = torch.randn(B, Tq, D, device=DEV)
        kv = torch.randn(B, Tk, 2 * D, device=DEV)
        out = torch.empty(B, Tq, D, device=DEV)
        ws = torch.empty(max(1, _lib.call_value("bd_attention_workspace", B, H, Tq, Tk, math_mode)), device=DEV)
        args = (q.data_ptr(), kv.data_ptr(), kv.data_ptr() + 4 * D, out.data_ptr(), B, H, Tq, Tk, D, 2 * D, 2 * D, D,
                math_mode, ws.data_ptr(), 0)
        for _ in range(3):
            _lib.call("bd_attention", *args)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            _lib.call("bd_attention", *args)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{mname:7s} Tq={Tq} Tk={Tk}: {ms:7.3f} ms  {4.0 * B * Tq * Tk * D / ms / 1e9:7.1f} TF/s", flush=True)
