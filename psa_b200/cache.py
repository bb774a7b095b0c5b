"""The trajectory ``.npy`` cache: the step immediately before the hot path.

The reference's loader short-circuits OVITO when ``<stem>.positions.npy``, ``.velocities.npy``,
``.types.npy`` and ``.box_matrix.npy`` sit next to the trajectory file
(reference: src/psa/io/loader.py:48-76) and writes them with ``save_trajectory_npy``
(reference: src/psa/io/loader.py:363-387).  This module reads the same bundle, but maps the two big
arrays instead of reading them: ``SEDCalculator`` then streams them to the GPU through double-buffered
pinned staging (``psa_b200.engine.DeviceTrajectory``), so host memory never holds a second copy and the
copy engine overlaps the file reads.  Field values (timesteps, box lengths/tilts) follow the
reference's cache branch exactly.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict

import numpy as np

from .trajectory import Trajectory

_PARTS = ("positions", "velocities", "types", "box_matrix")


def cache_files(trajectory_file) -> Dict[str, Path]:
    path = Path(trajectory_file)
    stem = path.parent / path.stem
    return {name: stem.with_suffix(f".{name}.npy") for name in _PARTS}


def has_npy_cache(trajectory_file) -> bool:
    return all(f.exists() for f in cache_files(trajectory_file).values())


def load_npy_cache(trajectory_file, dt: float = 1.0, mmap: bool = True) -> Trajectory:
    """Trajectory from the cache bundle of ``trajectory_file`` (the file itself need not exist)."""
    if dt <= 0:
        raise ValueError("dt (timestep size) must be positive.")
    files = cache_files(trajectory_file)
    missing = [str(f) for f in files.values() if not f.exists()]
    if missing:
        raise FileNotFoundError(f"No complete .npy cache; missing: {missing}")
    mode = "c" if mmap else None           # copy-on-write map: writable view, file untouched
    pos = np.load(files["positions"], mmap_mode=mode)
    vel = np.load(files["velocities"], mmap_mode=mode)
    types = np.load(files["types"])
    box = np.load(files["box_matrix"])
    if box.shape != (3, 3):
        raise ValueError(f"Cached box_matrix has shape {box.shape}, expected (3,3).")
    lengths = np.array([box[0, 0], box[1, 1], box[2, 2]], dtype=np.float32)
    tilts = np.array([box[0, 1], box[0, 2], box[1, 2]], dtype=np.float32)
    steps = np.arange(pos.shape[0], dtype=np.float32) * dt
    return Trajectory(pos, vel, types, steps, box_matrix=box, box_lengths=lengths, box_tilts=tilts, dt_ps=dt)


def save_npy_cache(traj: Trajectory, trajectory_file) -> None:
    """Write the four arrays the loader looks for (not the derived mean/displacement files)."""
    files = cache_files(trajectory_file)
    next(iter(files.values())).parent.mkdir(parents=True, exist_ok=True)
    for name, path in files.items():
        np.save(path, getattr(traj, name))
