"""Atom-group resolution for the SED driver and for iSED (host side, integer logic).

Which atoms enter a projection, and whether several groups are summed as
amplitudes or as intensities, is decided here; the CUDA path only ever sees
explicit index lists.  The rules restate the reference's, quirks included:

* ``calculate``  (reference: src/psa/core/sed_calculator.py:208-266, 276)
* ``ised``       (reference: src/psa/core/sed_calculator.py:389-433)
"""
from __future__ import annotations

import logging
from typing import List, Optional, Sequence, Tuple

import numpy as np

logger = logging.getLogger(__name__)


def _atoms_of_types(types: np.ndarray, type_group: Sequence[int]) -> np.ndarray:
    return np.where(np.isin(types, type_group))[0]


def _check_bounds(idx: np.ndarray, n_atoms: int, msg: str) -> None:
    if np.any(idx >= n_atoms) or np.any(idx < 0):
        raise ValueError(msg)


def resolve_sed_groups(types: np.ndarray, n_atoms: int,
                       basis_atom_indices=None, basis_atom_types=None,
                       summation_mode: str = "coherent") -> List[np.ndarray]:
    """Index lists (one per group) for ``SEDCalculator.calculate``.

    * a flat list of types is one union group when coherent, one group per type when incoherent;
    * a list of lists is taken as given; unknown types are skipped with a warning;
    * types win over indices when both are passed;
    * nothing usable  ->  a single group holding every atom.
    """
    groups: List[np.ndarray] = []

    if basis_atom_types is not None:
        if basis_atom_indices is not None:
            logger.warning("Both basis_atom_types and basis_atom_indices provided. Using basis_atom_types.")
        type_groups: List[List[int]] = []
        if isinstance(basis_atom_types, list) and len(basis_atom_types) > 0:
            if all(isinstance(item, list) for item in basis_atom_types):
                type_groups = basis_atom_types
            elif all(isinstance(item, int) for item in basis_atom_types):
                if summation_mode == "incoherent":
                    type_groups = [[t] for t in basis_atom_types]
                else:
                    type_groups = [list(basis_atom_types)]
            else:
                raise ValueError("basis_atom_types must be a list of ints or a list of lists of ints.")
        elif isinstance(basis_atom_types, int):
            type_groups = [[basis_atom_types]]
        for type_group in type_groups:
            idx = _atoms_of_types(types, type_group)
            if idx.size > 0:
                groups.append(idx)
            else:
                logger.warning("No atoms found for type group %s. Skipping.", type_group)

    elif basis_atom_indices is not None:
        candidates: List[np.ndarray] = []
        if isinstance(basis_atom_indices, list):
            if len(basis_atom_indices) == 0:
                pass
            elif all(isinstance(item, list) for item in basis_atom_indices):
                candidates = [np.asarray(sub, dtype=int) for sub in basis_atom_indices]
            elif all(isinstance(item, int) for item in basis_atom_indices):
                candidates = [np.asarray(basis_atom_indices, dtype=int)]
            else:
                raise ValueError("basis_atom_indices must be a list of ints or a list of lists of ints.")
        elif isinstance(basis_atom_indices, np.ndarray):
            if basis_atom_indices.ndim == 1 and basis_atom_indices.size > 0:
                candidates = [basis_atom_indices.astype(int)]
            else:
                logger.warning("Unsupported np.ndarray format for basis_atom_indices. "
                               "Using all atoms if no other basis defined.")
        for idx in candidates:
            if idx.size == 0:
                continue
            _check_bounds(idx, n_atoms, "Atom indices in basis out of bounds.")
            groups.append(idx)

    if not groups:
        logger.debug("No specific basis provided or basis resulted in empty groups. "
                     "Using all %d atoms as a single group.", n_atoms)
        groups.append(np.arange(n_atoms))
        if summation_mode == "incoherent" and n_atoms > 0:
            logger.info("Using all atoms. Incoherent sum will effectively be a coherent sum of all atoms.")
    return groups


def plan_sed_groups(groups: List[np.ndarray], summation_mode: str) -> Tuple[bool, List[np.ndarray]]:
    """``(complex_output, projection_groups)``.

    Complex output (coherent, or fewer than two groups) projects the sorted
    union of all groups once; otherwise every group is projected separately and
    their intensities are added (reference: sed_calculator.py:276, 296-327).
    """
    if summation_mode == "coherent" or len(groups) <= 1:
        if len(groups) > 1:
            merged = np.unique(np.concatenate(groups)).astype(int)
        else:
            merged = groups[0]
        return True, [merged]
    return False, [g for g in groups if g.size > 0]


def resolve_ised_groups(types: np.ndarray, n_atoms: int,
                        basis_atom_idx_ised: Optional[list] = None,
                        basis_atom_types_ised: Optional[list] = None) -> List[np.ndarray]:
    """Reconstruction groups for ``ised``: a flat type list means one group PER type."""
    sys_types = types.astype(int)
    groups: List[np.ndarray] = []
    if basis_atom_idx_ised and len(basis_atom_idx_ised) > 0:
        if isinstance(basis_atom_idx_ised[0], list):
            for sub in basis_atom_idx_ised:
                arr = np.asarray(sub, dtype=int)
                _check_bounds(arr, n_atoms, f"Atom indices in group {sub} out of bounds.")
                if arr.size > 0:
                    groups.append(arr)
        else:
            arr = np.asarray(basis_atom_idx_ised, dtype=int)
            _check_bounds(arr, n_atoms, "Atom indices out of bounds.")
            if arr.size > 0:
                groups.append(arr)
        if basis_atom_types_ised and len(basis_atom_types_ised) > 0:
            logger.warning("iSED: atom_indices and atom_types provided. Using atom_indices.")
    elif basis_atom_types_ised and len(basis_atom_types_ised) > 0:
        nested = isinstance(basis_atom_types_ised[0], list)
        for entry in basis_atom_types_ised:
            idx = _atoms_of_types(sys_types, entry if nested else [entry])
            if idx.size > 0:
                groups.append(idx)
            else:
                logger.warning("No atoms for type%s %s in iSED.", " group" if nested else "", entry)
    else:
        groups.append(np.arange(n_atoms))
    return groups
