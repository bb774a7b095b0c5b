"""Result container of the SED hot path.

Field-for-field compatible with the reference ``SED`` dataclass
(reference: src/psa/core/sed.py:12-68): ``SEDPlotter`` and the GUI read
``sed / freqs / k_points / k_vectors / k_grid_shape / phase / is_complex`` and
nothing else, and ``save``/``load`` use the same ``<base>.<field>.npy`` bundle
so caches written by either implementation are interchangeable.

Additions (all optional, never required by a consumer of the reference type):

* mapping access (``result['sed']``, ``result.keys()``) because the README
  facade describes the result as a dict (reference: README.md:103-111);
* ``context``: the geometry the inverse projection needs (mean positions,
  types, box, k-direction, group lists) so that ``iSEDReconstructor(result)``
  can work from the result alone (reference: README.md:148-167).
"""
from __future__ import annotations

import logging
from dataclasses import dataclass, field, fields
from pathlib import Path
from typing import Any, Dict, Iterator, Optional, Tuple

import numpy as np

logger = logging.getLogger(__name__)

_REQUIRED = ("sed", "freqs", "k_points", "k_vectors")


@dataclass
class SED:
    sed: np.ndarray
    freqs: np.ndarray
    k_points: np.ndarray
    k_vectors: np.ndarray
    k_grid_shape: Optional[Tuple[int, ...]] = None
    phase: Optional[np.ndarray] = None
    is_complex: bool = True
    # not part of the reference dataclass; excluded from equality / repr
    context: Optional[Dict[str, Any]] = field(default=None, repr=False, compare=False)

    # -- reference property (sed.py:22-24); the 2-D "incoherent" quirk is kept on purpose
    @property
    def intensity(self) -> np.ndarray:
        mag = np.abs(self.sed)
        return np.sum(mag * mag, axis=-1).astype(np.float32)

    def __getstate__(self):
        """Pickle / copy without the back-reference to the calculator (a weakref, and GPU state behind it)."""
        state = dict(self.__dict__)
        if state.get("context"):
            state["context"] = {k: v for k, v in state["context"].items() if k != "calculator"}
        return state

    # -- dict-style access for the README facade
    def keys(self) -> Iterator[str]:
        return (f.name for f in fields(self) if f.name != "context")

    def __getitem__(self, key: str) -> Any:
        if key == "intensity":
            return self.intensity
        if key not in {f.name for f in fields(self)}:
            raise KeyError(key)
        return getattr(self, key)

    def __contains__(self, key: object) -> bool:
        return key == "intensity" or key in {f.name for f in fields(self)}

    # -- persistence: same file bundle as the reference (sed.py:26-68)
    def save(self, base_path: Path) -> None:
        base_path = Path(base_path)
        base_path.parent.mkdir(parents=True, exist_ok=True)
        for name in _REQUIRED:
            np.save(base_path.with_suffix(f".{name}.npy"), getattr(self, name))
        if self.k_grid_shape is not None:
            np.save(base_path.with_suffix(".k_grid_shape.npy"), np.array(self.k_grid_shape))
        if self.phase is not None:
            np.save(base_path.with_suffix(".phase.npy"), self.phase)
        logger.info("SED data saved: %s.*.npy", base_path.name)

    @staticmethod
    def load(base_path: Path) -> "SED":
        base_path = Path(base_path)
        missing = [n for n in _REQUIRED if not base_path.with_suffix(f".{n}.npy").exists()]
        if missing:
            raise FileNotFoundError(f"Required SED files missing for base: {base_path.name}")
        loaded = {n: np.load(base_path.with_suffix(f".{n}.npy")) for n in _REQUIRED}

        phase = None
        phase_file = base_path.with_suffix(".phase.npy")
        if phase_file.exists():
            try:
                phase = np.load(phase_file)
            except Exception as exc:  # unreadable cache entries are skipped, not fatal
                logger.warning("Could not load phase data from %s: %s", phase_file.name, exc)

        grid_shape = None
        grid_file = base_path.with_suffix(".k_grid_shape.npy")
        if grid_file.exists():
            try:
                grid_shape = tuple(int(v) for v in np.load(grid_file))
            except Exception as exc:
                logger.warning("Could not load k_grid_shape data from %s: %s", grid_file.name, exc)

        return SED(loaded["sed"], loaded["freqs"], loaded["k_points"], loaded["k_vectors"],
                   k_grid_shape=grid_shape, phase=phase)
