#!/usr/bin/env python
"""How long does the host take to ISSUE one default-workload step (no synchronisation), vs the GPU time?"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from psa_b200 import SEDCalculator  # noqa: E402
import synthetic as synth  # noqa: E402

cfg = synth.baseline_config("c2")
spec = cfg["spec"]
traj = spec.trajectory()
calc = SEDCalculator(traj, *spec.cells)
mags, kv = calc.get_k_path([1, 1, 0], 4.0, 200)
_ = calc.device_trajectory.positions, calc.device_trajectory.velocities


def step():
    calc.device_trajectory.reset_derived()
    return calc._calculate_device(kv, None, [1], "incoherent")


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    step()
t_issue = (time.perf_counter() - t0) / 20
torch.cuda.synchronize()
t_total = (time.perf_counter() - t0) / 20
print(f"host issue time per step {t_issue * 1e3:.3f} ms; wall per step incl. GPU {t_total * 1e3:.3f} ms")
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
