#!/usr/bin/env python
"""Time psa_project alone (CUDA events): scripts/project_tune.py rows n_t n_sel [impl]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from psa_b200.engine import Engine  # noqa: E402

rows, n_t, n_sel = (int(v) for v in sys.argv[1:4])
impl = int(sys.argv[4]) if len(sys.argv) > 4 else 0
eng = Engine()
pitch = -(-n_sel // 64) * 64
ad = torch.randint(-128, 127, (4, rows, pitch), dtype=torch.int8, device=eng.device)
bd = torch.randint(-128, 127, (3, 4, n_t, pitch), dtype=torch.int8, device=eng.device)
ad[3] //= 2
bd[:, 3] //= 2
ex = torch.zeros((3, n_t), dtype=torch.int32, device=eng.device)
P = torch.empty((rows, 3, n_t), dtype=torch.float32, device=eng.device)
for _ in range(3):
    eng.project(ad, rows, rows, bd, ex, n_t, n_sel, pitch, P, n_t, impl=impl)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    eng.project(ad, rows, rows, bd, ex, n_t, n_sel, pitch, P, n_t, impl=impl)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
ops = 120.0 * (rows / 2) * n_t * n_sel
tiles = -(-rows // 128) * -(-n_t // 128) * 3
print(f"impl={impl} rows={rows} n_t={n_t} n_sel={n_sel}: {ms:.3f} ms  {ops / ms / 1e12:.2f} int8 POPS  "
      f"{ms * 1e3 / (tiles / 148):.2f} us per (128-frame) tile slot")
