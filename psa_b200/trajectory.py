"""Input boundary type of the SED hot path.

Field-compatible with the reference's ``Trajectory`` dataclass
(reference: src/psa/core/trajectory.py:8-44) so that anything which builds a
PSA trajectory (the loader, user scripts) can hand it to
:class:`psa_b200.SEDCalculator` unchanged.  Only shape validation lives here;
the arrays stay NumPy arrays owned by the caller.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# (field, required shape suffix or ndim, message) checked in order; the first violation raises ValueError
_PER_FRAME_VECTORS = ("positions", "velocities")
_FIXED_SHAPES = (("box_matrix", (3, 3), "Box matrix must be 3x3, got {shape}"),
                 ("box_lengths", (3,), "Box lengths must be a 3-element array, got {shape}"),
                 ("box_tilts", (3,), "Box tilts must be a 3-element array, got {shape}"))


@dataclass
class Trajectory:
    positions: np.ndarray      # (n_frames, n_atoms, 3)
    velocities: np.ndarray     # (n_frames, n_atoms, 3)
    types: np.ndarray          # (n_atoms,)
    timesteps: np.ndarray      # (n_frames,)
    box_matrix: np.ndarray     # (3, 3), rows are the supercell vectors
    box_lengths: np.ndarray    # (3,)
    box_tilts: np.ndarray      # (3,)
    dt_ps: float               # sampling interval, picoseconds

    def __post_init__(self) -> None:
        first = next(self._violations(), None)
        if first is not None:
            raise ValueError(first)

    def _violations(self):
        for name in _PER_FRAME_VECTORS:
            arr = getattr(self, name)
            if arr.ndim != 3 or arr.shape[2] != 3:
                yield f"{name.capitalize()} must be 3D (frames, atoms, xyz) and last dimension must be 3."
        for name in ("types", "timesteps"):
            if getattr(self, name).ndim != 1:
                yield f"{name.capitalize()} must be 1D"
        counts = {name: getattr(self, name).shape[:2] for name in _PER_FRAME_VECTORS}
        if any(c[:1] != (len(self.timesteps),) for c in counts.values()):
            yield "Frame count mismatch: positions, velocities, timesteps."
        if any(c[1:2] != (len(self.types),) for c in counts.values()):
            yield "Atom count mismatch: positions, velocities, types."
        for name, want, msg in _FIXED_SHAPES:
            shape = getattr(self, name).shape
            if shape != want:
                yield msg.format(shape=shape)

    @property
    def n_frames(self) -> int:
        return len(self.timesteps)

    @property
    def n_atoms(self) -> int:
        return len(self.types)
