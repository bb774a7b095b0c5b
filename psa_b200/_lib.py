"""ctypes binding of ``libpsa_b200.so`` (the C ABI declared in ``include/psa_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` (``nvcc -gencode
arch=compute_100a,code=sm_100a``).  There is no CPU or PyTorch fallback: if the
shared object is missing, or the device is not an sm_100 part, loading fails
loudly.  ``ctypes.CDLL`` releases the GIL for the duration of each call, so the
GUI's worker threads (reference: src/psa/gui/psa_gui.py:1015) stay responsive.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_void_p
from pathlib import Path

OK = 0
ERR_BAD_ARG = -1
ERR_CUDA = -2
ERR_UNSUPPORTED = -3

MODE_COHERENT = 0
MODE_INCOHERENT = 1
PROJECT_TENSOR = 0
PROJECT_SIMT = 1

LIB_PATH = Path(__file__).resolve().parent / "libpsa_b200.so"

_SIGNATURES = {
    "psa_version": (c_int, []),
    "psa_last_error": (c_char_p, []),
    "psa_device_check": (c_int, [c_int]),
    "psa_pitch": (c_int64, [c_int64]),
    "psa_mean_positions": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "psa_digitize": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64,
                             c_void_p, c_void_p, c_void_p]),
    "psa_phase_digits": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                 c_void_p, c_void_p]),
    "psa_project": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                            c_void_p, c_int64, c_int, c_void_p]),
    "psa_project_routed": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p,
                                   c_void_p, c_int, c_int64, c_void_p]),
    "psa_project_rows": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64,
                                 c_void_p, c_int64, c_int, c_void_p]),
    "psa_fft_plan_bytes": (c_int64, [c_int64]),
    "psa_fft_plan_init": (c_int, [c_int64, c_void_p, c_void_p]),
    "psa_fft_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64]),
    "psa_fft_sed": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64,
                            c_void_p, c_int, c_void_p, c_int64, c_int64, c_void_p]),
    "psa_chiral_phase": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "psa_intensity": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "psa_ised_absmax": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                c_int64, c_void_p, c_void_p]),
    "psa_ised_frames": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "psa_gather_bins": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "psa_disp_moments": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "psa_absmax": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "psa_scale_intensity": (c_int, [c_void_p, c_int64, c_int, c_void_p]),
    "psa_minmax": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "psa_select_pass": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "psa_mean_accumulate": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "psa_digitize_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p,
                                  c_void_p, c_int64, c_int64, c_void_p]),
    "psa_copy_rows": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "psa_digitize_rows_peers": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p,
                                        c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p]),
    "psa_write_dump": (c_int, [c_char_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int]),
    "psa_host_register": (c_int, [c_void_p, c_int64]),
    "psa_host_unregister": (c_int, [c_void_p]),
    "psa_ipc_export": (c_int, [c_void_p, c_void_p, c_void_p]),
    "psa_ipc_open": (c_int, [c_void_p, c_void_p]),
    "psa_ipc_close": (c_int, [c_void_p]),
}

EXPORTS = tuple(_SIGNATURES)

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library once; raise if it has not been built."""
    global _lib
    if _lib is None:
        path = Path(os.environ.get("PSA_B200_LIB", LIB_PATH))
        if not path.exists():
            raise RuntimeError(
                f"{path} not found: the CUDA library is not built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` at the repo root "
                "(needs nvcc; there is no CPU fallback).")
        lib = ctypes.CDLL(str(path))
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here = header and library out of sync
            fn.restype, fn.argtypes = restype, argtypes
        _lib = lib
    return _lib


def check(status: int) -> None:
    """Map a C status to the exception type the reference API uses for that failure."""
    if status == OK:
        return
    msg = load().psa_last_error().decode("utf-8", "replace")
    if status == ERR_BAD_ARG:
        raise ValueError(msg)
    if status == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args))
