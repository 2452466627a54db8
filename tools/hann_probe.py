"""Probe for the round-1 Hann-window observation: in a fresh process, how far is torch.hann_window(4096) (float32, the
reference's window, spec.py:20,41) from the float64 closed form rounded once to float32 (what engine and oracle use)?
Run in a loop of fresh processes:  for i in $(seq 50); do python tools/hann_probe.py; done | sort | uniq -c"""
import numpy as np
import torch

k = np.arange(4096, dtype=np.float64)
exact = (0.5 - 0.5 * np.cos(2 * np.pi * k / 4096)).astype(np.float32)
w = torch.hann_window(4096).numpy()
dev = "cpu"
line = f"cpu max|torch - closed form| = {np.abs(w - exact).max():.3e} (rel to 1), sum = {float(w.astype(np.float64).sum()):.9f}"
if torch.cuda.is_available():
    wg = torch.hann_window(4096, device="cuda").cpu().numpy()
    line += f"; cuda max = {np.abs(wg - exact).max():.3e}, sum = {float(wg.astype(np.float64).sum()):.9f}"
print(line)
