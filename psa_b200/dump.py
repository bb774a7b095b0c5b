"""LAMMPS text dump of reconstructed (iSED) frames.

Byte-for-byte the format the reference writes (reference: src/psa/io/writer.py:139-228) and its GUI
parses back (reference: src/psa/gui/psa_gui.py:1396-1455): per frame ``ITEM: TIMESTEP``, atom count,
orthogonal (``pp pp pp``) or triclinic (``xy xz yz pp pp pp``) box bounds with 8 decimals, then
``id type x y z`` with 6 decimals.  The formatting runs in the C++ library (``psa_write_dump``: a thread pool,
exact integer ``%.6f``), not in Python - the reference's one ``write`` per atom per frame is the bottleneck
of a batched reconstruction (6.4 M lines per point on the 64 000-atom config).
"""
from __future__ import annotations

import os
from pathlib import Path

import numpy as np

from . import _lib


def write_lammps_dump(filename: str, positions_tf: np.ndarray, types_tf: np.ndarray, box_matrix: np.ndarray,
                      threads: int = 0) -> None:
    frames = np.ascontiguousarray(positions_tf, dtype=np.float32)
    if frames.ndim != 3 or frames.shape[2] != 3:
        raise ValueError("positions_tf must have shape (n_frames, n_atoms, 3)")
    n_fr, n_at, _ = frames.shape
    types = np.ascontiguousarray(np.asarray(types_tf).astype(int), dtype=np.int32)
    if types.shape != (n_at,):
        raise ValueError("types_tf must have one entry per atom")
    box = np.ascontiguousarray(box_matrix, dtype=np.float32)
    if box.shape != (3, 3):
        raise ValueError("box_matrix must be 3x3")
    Path(filename).parent.mkdir(parents=True, exist_ok=True)
    _lib.call("psa_write_dump", os.fsencode(str(filename)), frames.ctypes.data, types.ctypes.data, n_fr, n_at,
              box.ctypes.data, int(threads))
