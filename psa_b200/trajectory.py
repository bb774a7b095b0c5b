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


def _require(cond: bool, msg: str) -> None:
    if not cond:
        raise ValueError(msg)


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
        for name in ("positions", "velocities"):
            arr = getattr(self, name)
            _require(arr.ndim == 3 and arr.shape[2] == 3,
                     f"{name.capitalize()} must be 3D (frames, atoms, xyz) and last dimension must be 3.")
        _require(self.types.ndim == 1, "Types must be 1D")
        _require(self.timesteps.ndim == 1, "Timesteps must be 1D")
        n_fr = len(self.timesteps)
        _require(self.positions.shape[0] == n_fr and self.velocities.shape[0] == n_fr,
                 "Frame count mismatch: positions, velocities, timesteps.")
        n_at = len(self.types)
        _require(self.positions.shape[1] == n_at and self.velocities.shape[1] == n_at,
                 "Atom count mismatch: positions, velocities, types.")
        _require(self.box_matrix.shape == (3, 3),
                 f"Box matrix must be 3x3, got {self.box_matrix.shape}")
        _require(self.box_lengths.shape == (3,),
                 f"Box lengths must be a 3-element array, got {self.box_lengths.shape}")
        _require(self.box_tilts.shape == (3,),
                 f"Box tilts must be a 3-element array, got {self.box_tilts.shape}")

    @property
    def n_frames(self) -> int:
        return len(self.timesteps)

    @property
    def n_atoms(self) -> int:
        return len(self.types)
