"""LAMMPS text dump of reconstructed (iSED) frames.

Byte-for-byte the format the reference writes (reference: src/psa/io/writer.py:139-228) and its GUI
parses back (reference: src/psa/gui/psa_gui.py:1396-1455): per frame ``ITEM: TIMESTEP``, atom count,
orthogonal (``pp pp pp``) or triclinic (``xy xz yz pp pp pp``) box bounds with 8 decimals, then
``id type x y z`` with 6 decimals.  Rows are formatted per frame with one vectorised ``%`` expansion
instead of one ``write`` per atom.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np


def write_lammps_dump(filename: str, positions_tf: np.ndarray, types_tf: np.ndarray, box_matrix: np.ndarray) -> None:
    n_fr, n_at, _ = positions_tf.shape
    Path(filename).parent.mkdir(parents=True, exist_ok=True)

    xhi, yhi, zhi = box_matrix[0, 0], box_matrix[1, 1], box_matrix[2, 2]
    xy, xz, yz = box_matrix[0, 1], box_matrix[0, 2], box_matrix[1, 2]
    triclinic = not (np.isclose(xy, 0.0) and np.isclose(xz, 0.0) and np.isclose(yz, 0.0))
    if triclinic:
        bounds = [(0.0 + min(0.0, xy, xz, xy + xz), xhi + max(0.0, xy, xz, xy + xz), xy),
                  (0.0 + min(0.0, yz), yhi + max(0.0, yz), xz),
                  (0.0, zhi, yz)]
        box_txt = "ITEM: BOX BOUNDS xy xz yz pp pp pp\n" + "".join(
            f"{lo:.8f} {hi:.8f} {tilt:.8f}\n" for lo, hi, tilt in bounds)
    else:
        box_txt = "ITEM: BOX BOUNDS pp pp pp\n" + "".join(
            f"{lo:.8f} {hi:.8f}\n" for lo, hi in ((0.0, xhi), (0.0, yhi), (0.0, zhi)))

    ids = np.arange(1, n_at + 1)
    types = np.asarray(types_tf).astype(int)
    row_fmt = "%d %d %.6f %.6f %.6f\n" * n_at
    with open(filename, "w") as fh:
        for i_fr in range(n_fr):
            fh.write(f"ITEM: TIMESTEP\n{i_fr}\nITEM: NUMBER OF ATOMS\n{n_at}\n")
            fh.write(box_txt)
            fh.write("ITEM: ATOMS id type x y z\n")
            xyz = positions_tf[i_fr].astype(np.float64)
            cols = np.empty((n_at, 5), dtype=object)
            cols[:, 0], cols[:, 1] = ids.tolist(), types.tolist()
            cols[:, 2], cols[:, 3], cols[:, 4] = xyz[:, 0].tolist(), xyz[:, 1].tolist(), xyz[:, 2].tolist()
            fh.write(row_fmt % tuple(cols.reshape(-1).tolist()))
