#!/usr/bin/env python
"""Time psa_digitize alone (all atoms): scripts/digitize_tune.py n_t n_a"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from psa_b200.engine import Engine  # noqa: E402

n_t, n_a = int(sys.argv[1]), int(sys.argv[2])
eng = Engine()
data = torch.randn((n_t, n_a, 3), device=eng.device)
for _ in range(3):
    eng.digitize(data, None, None, n_a)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    eng.digitize(data, None, None, n_a)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"cluster={os.environ.get('PSA_DIG_CLUSTER', 'auto')} stage={'off' if os.environ.get('PSA_DIGITIZE_NO_STAGE') else 'on'} "
      f"n_t={n_t} n_a={n_a}: {ms:.3f} ms  {24.0 * n_t * n_a / ms / 1e6:.0f} GB/s (12 B read + 12 B written per frame-atom)")
