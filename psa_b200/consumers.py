"""Device-side consumers of a result (N3 of SURVEY.md 8f).

What the reference's plotter and GUI compute on the host from the whole ``(n_f, n_k)`` intensity array - the
intensity scaling (reference: src/psa/visualization/sed_plotter.py:160-181; ``_apply_intensity_scaling`` in
src/psa/gui/psa_gui.py), the global colour range over every frequency slice (psa_gui.py:2424-2441, ``np.nanmin`` /
``np.nanmax``) and the percentile colour limits (sed_plotter.py:211-215, ``np.percentile`` over the finite values) -
evaluated on the GPU, so that a 3.9 GB complex result never has to reach the host just to be reduced to a heat map
or to two numbers.  All arithmetic is in ``libpsa_b200.so``; this module sequences the calls.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

SCALE_MODES = {"linear": 0, "log": 1, "sqrt": 2, "dsqrt": 3}


def scale_intensity(eng, x: torch.Tensor, scale: str) -> None:
    """In place: 'log' = log10(max(x, 1e-12)), 'sqrt', 'dsqrt' = sqrt(sqrt(.)) of max(x, 0)."""
    mode = SCALE_MODES.get(str(scale).lower())
    if mode is None:
        raise ValueError(f"intensity scale must be one of {sorted(SCALE_MODES)}, got {scale!r}")
    eng._run("psa_scale_intensity", 1 if mode else 0, x.data_ptr(), x.numel(), mode, eng.stream())


def _key_to_float(key: int) -> np.float32:
    bits = (key ^ 0x80000000) if key & 0x80000000 else (~key & 0xFFFFFFFF)
    return np.frombuffer(struct.pack("<I", bits), np.float32)[0]


def nan_range(eng, x: torch.Tensor) -> Tuple[np.float32, np.float32, int]:
    """``(np.nanmin(x), np.nanmax(x), number of finite values)`` of a device float32 array."""
    out = eng.empty((4,), torch.int32)
    eng._run("psa_minmax", 1, x.data_ptr(), x.numel(), out.data_ptr(), eng.stream())
    raw = out.cpu().numpy().view(np.uint32)
    n_finite = int(raw[2]) | (int(raw[3]) << 32)
    if int(raw[0]) == 0xFFFFFFFF and int(raw[1]) == 0:                      # nothing but NaN (or empty)
        return np.float32(np.nan), np.float32(np.nan), n_finite
    return _key_to_float(int(raw[0])), _key_to_float(int(raw[1])), n_finite


def order_statistics(eng, x: torch.Tensor, ranks: Sequence[int]) -> List[np.float32]:
    """The values of the given 0-based ranks among the FINITE entries of ``x`` sorted ascending, exactly: a
    most-significant-byte-first radix select, four histogram passes for up to four ranks at once."""
    ranks = [int(r) for r in ranks]
    if not 1 <= len(ranks) <= 4:
        raise ValueError("1 to 4 ranks per call")
    m = len(ranks)
    prefixes = np.zeros(4, np.uint32)
    left = list(ranks)
    hist = eng.empty((m, 256), torch.int32)
    for done in (0, 8, 16, 24):
        pre_dev = eng.upload_small(prefixes)
        eng._run("psa_select_pass", 1, x.data_ptr(), x.numel(), done, pre_dev.data_ptr(), m, hist.data_ptr(), eng.stream())
        counts = hist.cpu().numpy().view(np.uint32).astype(np.int64)
        for j in range(m):
            cum = np.cumsum(counts[j])
            b = int(np.searchsorted(cum, left[j], side="right"))
            if b > 255:
                raise ValueError("rank beyond the number of finite values")
            left[j] -= int(cum[b - 1]) if b else 0
            prefixes[j] |= np.uint32(b << (24 - done))
    return [_key_to_float(int(prefixes[j])) for j in range(m)]


def percentiles(eng, x: torch.Tensor, qs: Sequence[float]) -> List[float]:
    """``np.percentile(x[isfinite(x)], q)`` (linear interpolation) for up to two percentiles, from exact order
    statistics found on the device, bit for bit: NumPy evaluates the quantile position of a float32 array in
    float32 (``q / float32(100)``, ``(n - 1) * q``), so the interpolation weight is quantised - mirrored here, as is
    its two-sided ``_lerp``.  Returns NaN when there is no finite value."""
    _, _, n = nan_range(eng, x)
    if n == 0:
        return [float("nan")] * len(qs)
    ranks, plan = [], []
    for q in qs:
        virt = (n - 1) * np.true_divide(float(q), np.float32(100))         # float32, like numpy/lib/_function_base_impl.py
        lo = min(max(int(np.floor(virt)), 0), n - 1)
        hi = min(lo + 1, n - 1)
        plan.append((len(ranks), np.float32(virt - np.float32(lo))))
        ranks += [lo, hi]
    vals: List[np.float32] = []
    for i in range(0, len(ranks), 4):
        vals += order_statistics(eng, x, ranks[i:i + 4])
    out = []
    for first, t in plan:
        a, b = vals[first], vals[first + 1]
        d = np.subtract(b, a)
        r = np.subtract(b, d * (1 - t)) if t >= 0.5 else np.add(a, d * t)
        out.append(float(r))
    return out


def intensity_stats(eng, x: torch.Tensor, vmin_percentile=None, vmax_percentile=None) -> Dict[str, float]:
    lo, hi, n = nan_range(eng, x)
    stats: Dict[str, float] = {"global_min": float(lo), "global_max": float(hi), "n_finite": n}
    qs = [q for q in (vmin_percentile, vmax_percentile) if q is not None]
    if qs:
        got = percentiles(eng, x, qs)
        if vmin_percentile is not None:
            stats["vmin"] = got.pop(0)
        if vmax_percentile is not None:
            stats["vmax"] = got.pop(0)
    return stats
