"""Micro-benchmark of bd_conv_gemm on the transformer / decoder shapes (CUDA events, L2 flushed by size)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demucs_b200 import _lib  # noqa: E402
from demucs_b200._lib import GemmDesc, ptr  # noqa: E402

_lib.build()
DEV = "cuda:0"
SHAPES = [("ffn1+gelu", 43008, 2048, 512, _lib.ACT_GELU), ("ffn2", 43008, 512, 2048, 0), ("qkv", 43008, 1536, 512, 0),
          ("proj", 43008, 512, 512, 0), ("proj_t", 21504, 512, 512, 0), ("thin_glu", 2752512, 96, 48, _lib.ACT_GLU),
          ("thin_k432", 2752512, 96, 432, _lib.ACT_GLU)]
ONLY = sys.argv[1].split(",") if len(sys.argv) > 1 else None
MODES = sys.argv[2].split(",") if len(sys.argv) > 2 else ["tf32", "tf32x3"]
for math_mode, mname in ((_lib.MATH_TF32, "tf32"), (_lib.MATH_TF32X3, "tf32x3")):
    if mname not in MODES:
        continue
    for name, M, N, K, act in SHAPES:
        if ONLY and name not in ONLY:
            continue
        x = torch.randn(M, K, device=DEV)
        w = torch.randn(N, K, device=DEV) / K ** 0.5
        b = torch.randn(N, device=DEV)
        out = torch.empty(M, N // 2 if act == _lib.ACT_GLU else N, device=DEV)
        d = GemmDesc()
        d.M, d.N, d.K, d.Cin, d.taps, d.I1, d.I0, d.m1, d.m0, d.J1, d.J0 = M, N, K, K, 1, 1, M, 1, 1, 1, M
        d.xs_0, d.xs_c, d.os_0 = K, 1, out.shape[1]
        d.x, d.w, d.bias, d.out, d.act, d.math = ptr(x), ptr(w), ptr(b), ptr(out), act, math_mode
        for _ in range(3):
            _lib.call("bd_conv_gemm", C.byref(d), 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            _lib.call("bd_conv_gemm", C.byref(d), 0)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{mname:7s} {name:10s} M={M} N={N} K={K}: {ms:7.3f} ms  {2.0 * M * N * K / ms / 1e9:7.1f} TF/s", flush=True)
