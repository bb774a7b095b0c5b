"""Direction specifications -> unit float32 3-vectors.

Host-side input of the k-path generator.  Behaviour follows the reference's
``parse_direction`` (reference: src/psa/utils/helpers.py:13-109), which is
pinned by its own tests (reference: tests/test_helpers.py:6-100): accepted
spellings, float32 arithmetic, error types and messages.  The result feeds
``np.outer(k_mags, k_hat)`` and therefore has to be bit-identical to the
reference's, because those k-vectors are the kernel's input.
"""
from __future__ import annotations

import logging
from typing import Any

import numpy as np

logger = logging.getLogger(__name__)

_S2 = 1 / np.sqrt(2)
_S3 = 1 / np.sqrt(3)

# axis letters and low-index Miller strings understood without parsing
_NAMED = {
    "x": (1, 0, 0), "y": (0, 1, 0), "z": (0, 0, 1),
    "100": (1, 0, 0), "010": (0, 1, 0), "001": (0, 0, 1),
    "xy": (_S2, _S2, 0), "yx": (_S2, _S2, 0), "110": (_S2, _S2, 0),
    "xz": (_S2, 0, _S2), "zx": (_S2, 0, _S2),
    "yz": (0, _S2, _S2), "zy": (0, _S2, _S2),
    "xyz": (_S3, _S3, _S3), "111": (_S3, _S3, _S3),
}


def _in_plane(angle_deg: float) -> np.ndarray:
    """Unit vector in the xy plane at ``angle_deg`` from +x."""
    rad = np.deg2rad(angle_deg)
    return np.array([np.cos(rad), np.sin(rad), 0.0], dtype=np.float32)


def _from_string(spec: str) -> np.ndarray:
    named = _NAMED.get(spec.lower())
    if named is not None:
        return np.array(named, dtype=np.float32)
    try:                                    # "37.5" -> in-plane angle
        return _in_plane(float(spec))
    except ValueError:
        pass
    parts = spec.replace(",", " ").split()  # "1 0 0" / "1,0,0" -> components
    if len(parts) == 3:
        try:
            return np.array([float(p) for p in parts], dtype=np.float32)
        except ValueError:
            pass
    raise ValueError(f"Unknown direction string: {spec}.")


def _from_sequence(spec: Any) -> np.ndarray:
    arr = np.asarray(spec, dtype=np.float32).squeeze()
    if arr.ndim == 0:
        return _in_plane(arr.item())
    if arr.ndim > 1:
        raise ValueError(f"Direction array has too many dims: {arr.ndim}, expected 0 or 1 (squeezed).")
    if arr.size == 1:
        return _in_plane(arr[0])
    if arr.size == 3:
        return arr
    raise ValueError(f"Direction array must have 1 (angle) or 3 (vector) components, got {arr.size}")


def _from_mapping(spec: dict) -> np.ndarray:
    if "angle" in spec:
        return _in_plane(float(spec["angle"]))
    if any(key in spec for key in ("h", "k", "l")):
        return np.array([float(spec.get(key, 0.0)) for key in ("h", "k", "l")], dtype=np.float32)
    raise ValueError("Direction dict must contain 'angle' or Miller indices ('h','k','l').")


def parse_direction(direction_spec: Any) -> np.ndarray:
    """Normalised float32 direction for a number (degrees in xy), string, 3-sequence or dict."""
    if isinstance(direction_spec, (int, float)):
        vec = _in_plane(float(direction_spec))
    elif isinstance(direction_spec, str):
        vec = _from_string(direction_spec)
    elif isinstance(direction_spec, (list, tuple, np.ndarray)):
        vec = _from_sequence(direction_spec)
    elif isinstance(direction_spec, dict):
        vec = _from_mapping(direction_spec)
    else:
        raise TypeError(f"Unsupported direction type: {type(direction_spec)}")

    if np.allclose(vec, 0, atol=1e-8):
        raise ValueError("Direction vector is zero. For k-path, direction must be non-zero if n_k > 1.")
    norm = np.linalg.norm(vec)
    if norm < 1e-9:
        logger.warning("Direction vector norm (%.2e) is very small, returning unnormalized vector.", norm)
        return vec
    return vec / norm
