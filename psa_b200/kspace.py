"""Reciprocal lattice, k-paths and k-grids (host side, O(n_k)).

These arrays are the *input* of the CUDA path, so their float32 bit patterns
have to equal the reference's.  The arithmetic below restates, operation for
operation, the reference's

* primitive/reciprocal vectors   (reference: src/psa/core/sed_calculator.py:40-56)
* ``get_k_path``                 (reference: src/psa/core/sed_calculator.py:86-125)
* ``get_k_grid``                 (reference: src/psa/core/sed_calculator.py:127-180)

Units are rad/Angstrom (they go straight into ``exp(i k.r)``).
"""
from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np

from .directions import parse_direction

logger = logging.getLogger(__name__)


@dataclass(frozen=True)
class Lattice:
    """Primitive cell derived from a supercell box and its repeat counts."""
    a1: np.ndarray
    a2: np.ndarray
    a3: np.ndarray
    b1: np.ndarray
    b2: np.ndarray
    b3: np.ndarray
    recip_vecs_prim: np.ndarray  # (3,3) float32, rows b1,b2,b3

    @staticmethod
    def from_box(box_matrix: np.ndarray, nx: int, ny: int, nz: int) -> "Lattice":
        if not (nx > 0 and ny > 0 and nz > 0):
            raise ValueError("System dimensions (nx, ny, nz) must be positive.")
        nx, ny, nz = int(nx), int(ny), int(nz)   # plain ints keep a float32 box in float32
        # supercell vectors are the ROWS of box_matrix
        a1, a2, a3 = box_matrix[0, :] / nx, box_matrix[1, :] / ny, box_matrix[2, :] / nz
        if any(np.linalg.norm(v) < 1e-9 for v in (a1, a2, a3)):
            raise ValueError("One or more primitive vectors (a1,a2,a3) near zero. Check nx,ny,nz or box matrix.")
        vol = np.abs(np.dot(a1, np.cross(a2, a3)))
        if np.isclose(vol, 0):
            cell = np.vstack([a1, a2, a3])
            if np.linalg.matrix_rank(cell) < 3 or np.isclose(np.linalg.det(cell), 0):
                raise ValueError(f"Primitive cell vectors coplanar/collinear; volume zero ({vol:.2e}).")
            logger.warning("Primitive cell volume very small (%.2e).", vol)
        pref = 2 * np.pi / vol
        b1 = pref * np.cross(a2, a3)
        b2 = pref * np.cross(a3, a1)
        b3 = pref * np.cross(a1, a2)
        recip = np.vstack([b1, b2, b3]).astype(np.float32)
        return Lattice(a1, a2, a3, b1, b2, b3, recip)


def k_path(lattice: Lattice, direction_spec, bz_coverage: float, n_k: int,
           lat_param: Optional[float] = None) -> Tuple[np.ndarray, np.ndarray]:
    """``(k_mags (n_k,) f32, k_vecs (n_k,3) f32)`` along one direction from Gamma."""
    k_hat = parse_direction(direction_spec)

    if lat_param is None or lat_param <= 1e-6:
        # extent of the first zone along k_hat = largest |b_i . k_hat|
        proj = [abs(np.dot(k_hat, b)) for b in (lattice.b1, lattice.b2, lattice.b3)]
        extent = max(proj)
        if extent > 1e-6:
            logger.info("Using directional reciprocal lattice projection (%.3f 2pi/A) for k-path.", extent)
        else:
            len_a1 = np.linalg.norm(lattice.a1)
            if not len_a1 > 1e-6:
                raise ValueError("Invalid/small lattice_param for k-path & reciprocal projections too small "
                                 "for auto-detection.")
            extent = 2 * np.pi / len_a1
            logger.warning("Reciprocal projections too small, using |a1| fallback (%.3f A).", len_a1)
    else:
        extent = 2 * np.pi / lat_param
        logger.info("Using provided lattice parameter (%.3f A) for k-path.", lat_param)

    k_max = bz_coverage * extent
    if n_k < 1:
        raise ValueError("n_k (k-points) must be >= 1.")
    if n_k > 1:
        k_mags = np.linspace(0, k_max, n_k, dtype=np.float32)
    else:
        k_mags = np.array([0.0 if np.isclose(k_max, 0) else k_max], dtype=np.float32)
    k_vecs = np.outer(k_mags, k_hat).astype(np.float32)
    return k_mags, k_vecs


# plane name -> (column of the first range, column of the second range, fixed column)
_PLANE_AXES = {"xy": (0, 1, 2), "yz": (1, 2, 0), "zx": (2, 0, 1)}


def k_grid(plane: str, k_range_x: Sequence[float], k_range_y: Sequence[float],
           n_kx: int, n_ky: int, k_fixed_val: float = 0.0
           ) -> Tuple[np.ndarray, np.ndarray, Tuple[int, int]]:
    """Regular 2-D k-grid: ``(empty f32, k_vecs (n_kx*n_ky,3) f32, (n_kx,n_ky))``.

    The first range is the slow (outer) index, as in the reference.
    """
    if n_kx <= 0 or n_ky <= 0:
        raise ValueError("Number of k-points (n_kx, n_ky) must be positive.")
    axes = _PLANE_AXES.get(plane.lower())
    if axes is None:
        raise ValueError(f"Invalid plane specified: {plane}. Must be 'xy', 'yz', or 'zx'.")
    first = np.linspace(k_range_x[0], k_range_x[1], n_kx, dtype=np.float32)
    second = np.linspace(k_range_y[0], k_range_y[1], n_ky, dtype=np.float32)
    c_first, c_second, c_fixed = axes
    k_vecs = np.empty((n_kx * n_ky, 3), dtype=np.float32)
    k_vecs[:, c_first] = np.repeat(first, n_ky)
    k_vecs[:, c_second] = np.tile(second, n_kx)
    k_vecs[:, c_fixed] = np.float32(k_fixed_val)
    return np.array([], dtype=np.float32), k_vecs, (n_kx, n_ky)
