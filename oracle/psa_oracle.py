"""CPU oracle for the PSA spectral-energy-density hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain NumPy, the algorithm of the reference
``psa.core.sed_calculator.SEDCalculator`` so that the CUDA path can be checked
on machines where the reference itself is absent (the GPU box).  Nothing under
``psa_b200/`` may import it; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do.

Pinning.  The reference's own tests hold NO golden vector for this path
(SURVEY.md section 8c), so the oracle is pinned against *outputs of the
reference itself*: ``oracle/make_golden.py`` imports ``/root/reference`` in the
build container, runs ``SEDCalculator.calculate / calculate_chiral_phase /
ised`` on seeded inputs and commits the results under ``tests/golden/``;
``tests/test_oracle.py`` requires this file to reproduce them bit for bit
(float32 path) on NumPy 2.3.x.  The third-party arithmetic underneath (NumPy ->
OpenBLAS sgemm/cgemm, pocketfft; unpinned in the reference's setup.py) is what
the image ships: NumPy 2.3.5 / OpenBLAS 0.3.30.

Two evaluations are provided:

* ``O-ref``  (``calculate``): the reference's float32/complex64 arithmetic, same
  NumPy calls in the same order -> defines parity.
* ``O-64``   (``calculate_fp64``): float64 accumulation and FFT fed the
  reference's *float32* mean positions, float32 ``k.r`` and complex64 phase
  table -> defines the truth for identical inputs, and therefore the
  reference's own rounding floor.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

C64 = np.complex64
F32 = np.float32


# --------------------------------------------------------------------------- stages

def mean_positions(positions: np.ndarray) -> np.ndarray:
    """Time-averaged positions in float32 (reference: sed_calculator.py:205, 384)."""
    return np.mean(positions, axis=0, dtype=F32)


def phase_table(k_vecs: np.ndarray, mean_pos_group: np.ndarray) -> np.ndarray:
    """exp(+i k.r) as complex64, theta formed in float32 (reference: sed_calculator.py:78)."""
    return np.exp(1j * np.dot(k_vecs, mean_pos_group.T))


def group_data(positions: np.ndarray, velocities: np.ndarray, idx: np.ndarray,
               mean_pos: np.ndarray, use_displacements: bool, weight: Optional[np.ndarray] = None) -> np.ndarray:
    """Per-group time series that gets projected (reference: sed_calculator.py:69-72).  ``weight`` (per-atom float32,
    e.g. sqrt(mass)) is the README-level extension the shipped source does not have: a float32 product."""
    data = positions[:, idx, :] - mean_pos[idx][None, :, :] if use_displacements else velocities[:, idx, :]
    if weight is not None:
        data = (data * weight[idx].astype(F32)[None, :, None]).astype(F32)
    return data


def group_sed(data: np.ndarray, k_vecs: np.ndarray, mean_pos_group: np.ndarray,
              window: Optional[np.ndarray] = None) -> np.ndarray:
    """Complex SED of one atom group, (n_t, n_k, 3) complex64 (reference: sed_calculator.py:58-84).  ``window``
    (float32 taper over the frames) is the README-level extension: the projected columns are multiplied by it
    before the FFT (the shipped source has no window = None)."""
    n_t = data.shape[0]
    n_k = len(k_vecs)
    if data.shape[1] == 0:
        return np.zeros((n_t, n_k, 3), C64)
    ph = phase_table(k_vecs, mean_pos_group)
    proj = np.zeros((n_t, n_k, 3), C64)
    for pol in range(3):
        proj[:, :, pol] = np.einsum("ta,ak->tk", data[:, :, pol], ph.T, optimize=True)
    if window is not None:
        proj = proj.astype(np.complex128) * window.astype(np.float64)[:, None, None]
    return (np.fft.fft(proj, axis=0) / n_t).astype(C64)


def group_sed_fp64(data: np.ndarray, k_vecs: np.ndarray, mean_pos_group: np.ndarray,
                   window: Optional[np.ndarray] = None) -> np.ndarray:
    """Same quantity in float64, fed the reference's complex64 phase table."""
    n_t = data.shape[0]
    ph = phase_table(k_vecs, mean_pos_group).astype(np.complex128)      # float32 inputs, exact widening
    proj = np.einsum("tap,ka->tkp", data.astype(np.float64), ph, optimize=True)
    if window is not None:
        proj = proj * window.astype(np.float64)[:, None, None]
    return np.fft.fft(proj, axis=0) / n_t


# --------------------------------------------------------------------------- group logic

def resolve_groups(types: np.ndarray, n_atoms: int, basis_atom_indices, basis_atom_types,
                   mode: str) -> List[np.ndarray]:
    """Atom groups of ``calculate`` (reference: sed_calculator.py:208-266)."""
    out: List[np.ndarray] = []
    if basis_atom_types is not None:
        tg: list = []
        if isinstance(basis_atom_types, list) and basis_atom_types:
            if all(isinstance(x, list) for x in basis_atom_types):
                tg = basis_atom_types
            elif all(isinstance(x, int) for x in basis_atom_types):
                tg = [[t] for t in basis_atom_types] if mode == "incoherent" else [list(basis_atom_types)]
            else:
                raise ValueError("basis_atom_types must be a list of ints or a list of lists of ints.")
        elif isinstance(basis_atom_types, int):
            tg = [[basis_atom_types]]
        for g in tg:
            sel = np.where(np.isin(types, g))[0]
            if sel.size:
                out.append(sel)
    elif basis_atom_indices is not None:
        cand: list = []
        if isinstance(basis_atom_indices, list):
            if basis_atom_indices and all(isinstance(x, list) for x in basis_atom_indices):
                cand = [np.asarray(s, dtype=int) for s in basis_atom_indices]
            elif basis_atom_indices and all(isinstance(x, int) for x in basis_atom_indices):
                cand = [np.asarray(basis_atom_indices, dtype=int)]
            elif basis_atom_indices:
                raise ValueError("basis_atom_indices must be a list of ints or a list of lists of ints.")
        elif isinstance(basis_atom_indices, np.ndarray):
            if basis_atom_indices.ndim == 1 and basis_atom_indices.size:
                cand = [basis_atom_indices.astype(int)]
        for c in cand:
            if c.size == 0:
                continue
            if np.any(c >= n_atoms) or np.any(c < 0):
                raise ValueError("Atom indices in basis out of bounds.")
            out.append(c)
    if not out:
        out.append(np.arange(n_atoms))
    return out


# --------------------------------------------------------------------------- drivers

def _calculate(positions, velocities, types, dt_ps, k_vecs, basis_atom_indices, basis_atom_types,
               summation_mode, use_displacements, k_chunk_size, fp64: bool, weight=None, window=None) -> Dict:
    if summation_mode not in ("coherent", "incoherent"):
        raise ValueError(f"summation_mode must be 'coherent' or 'incoherent', got {summation_mode}")
    n_t, n_atoms = positions.shape[0], positions.shape[1]
    if n_t == 0 or n_atoms == 0:
        return dict(sed=np.zeros((0, 0, 3), C64), freqs=np.array([], F32), is_complex=True)
    mean_pos = mean_positions(positions)
    freqs = np.fft.fftfreq(n_t, d=dt_ps)
    groups = resolve_groups(types, n_atoms, basis_atom_indices, basis_atom_types, summation_mode)
    n_k = len(k_vecs)
    is_complex = summation_mode == "coherent" or len(groups) <= 1
    cdt, rdt = (np.complex128, np.float64) if fp64 else (C64, F32)
    one = group_sed_fp64 if fp64 else group_sed
    out = np.zeros((n_t, n_k, 3), cdt) if is_complex else np.zeros((n_t, n_k), rdt)
    chunk = min(max(1, k_chunk_size), n_k) if n_k else 1
    for k0 in range(0, n_k, chunk):
        kv = k_vecs[k0:k0 + chunk]
        if is_complex:
            idx = np.unique(np.concatenate(groups)).astype(int) if len(groups) > 1 else groups[0]
            if idx.size == 0:
                continue
            data = group_data(positions, velocities, idx, mean_pos, use_displacements, weight)
            out[:, k0:k0 + chunk, :] = one(data, kv, mean_pos[idx], window)
        else:
            acc = np.zeros((n_t, kv.shape[0]), rdt)
            for idx in groups:
                if idx.size == 0:
                    continue
                data = group_data(positions, velocities, idx, mean_pos, use_displacements, weight)
                acc += np.sum(np.abs(one(data, kv, mean_pos[idx], window)) ** 2, axis=-1)
            out[:, k0:k0 + chunk] = acc
    return dict(sed=out, freqs=freqs, is_complex=is_complex, mean_pos=mean_pos, groups=groups)


def calculate(positions, velocities, types, dt_ps, k_vecs, basis_atom_indices=None,
              basis_atom_types=None, summation_mode="coherent", use_displacements=False,
              k_chunk_size=500, weight=None, window=None) -> Dict:
    """O-ref: ``SEDCalculator.calculate`` in the reference's float32 arithmetic
    (reference: sed_calculator.py:182-336).  Returns ``sed`` (complex64 (n_f,n_k,3) or float32
    (n_f,n_k)), ``freqs`` (float64, fftfreq order), ``is_complex``."""
    return _calculate(positions, velocities, types, dt_ps, k_vecs, basis_atom_indices,
                      basis_atom_types, summation_mode, use_displacements, k_chunk_size, fp64=False,
                      weight=weight, window=window)


def calculate_fp64(positions, velocities, types, dt_ps, k_vecs, basis_atom_indices=None,
                   basis_atom_types=None, summation_mode="coherent", use_displacements=False,
                   k_chunk_size=500, weight=None, window=None) -> Dict:
    """O-64: same driver, float64 contraction + FFT on the reference's float32 inputs."""
    return _calculate(positions, velocities, types, dt_ps, k_vecs, basis_atom_indices,
                      basis_atom_types, summation_mode, use_displacements, k_chunk_size, fp64=True,
                      weight=weight, window=window)


def intensity(sed: np.ndarray) -> np.ndarray:
    """Sum over the last axis of |sed|^2 as float32 (reference: sed.py:22-24)."""
    return np.sum(np.abs(sed) ** 2, axis=-1).astype(F32)


def chiral_phase(z1: np.ndarray, z2: np.ndarray, opt: str = "C") -> np.ndarray:
    """Folded phase difference of two polarisation components (reference: sed_calculator.py:338-371)."""
    if z1.shape != z2.shape:
        raise ValueError("Z1 and Z2 shapes must match for chiral phase.")
    if z1.size == 0:
        return np.array([], F32).reshape(z1.shape)
    if opt == "C":
        d = np.angle(z1) - np.angle(z2)
        d = (d + np.pi) % (2 * np.pi) - np.pi
        hi = d > (np.pi / 2)
        d[hi] = np.pi - d[hi]
        lo = d < (-np.pi / 2)
        d[lo] = -np.pi - d[lo]
        return d.astype(F32)
    # options "A"/"B": the reference loops over scalars of the arrays' own precision
    out = np.zeros(z1.shape, F32)
    r1, i1, r2, i2 = z1.real, z1.imag, z2.real, z2.imag
    m1, m2 = r1 * r1 + i1 * i1, r2 * r2 + i2 * i2
    ok = ~((m1 < 1e-18) | (m2 < 1e-18))
    den = np.sqrt(m1[ok]) * np.sqrt(m2[ok])
    if opt == "A":
        out[ok] = np.arccos(np.clip((r1[ok] * r2[ok] + i1[ok] * i2[ok]) / den, -1.0, 1.0))
    elif opt == "B":
        out[ok] = np.arcsin(np.clip((r1[ok] * i2[ok] - i1[ok] * r2[ok]) / den, -1.0, 1.0))
    return out


def k_path(recip_b: Sequence[np.ndarray], a1: np.ndarray, k_hat: np.ndarray, bz_coverage: float,
           n_k: int, lat_param: Optional[float] = None) -> Tuple[np.ndarray, np.ndarray]:
    """k magnitudes / vectors of a path (reference: sed_calculator.py:86-125); ``k_hat`` already unit."""
    if lat_param is None or lat_param <= 1e-6:
        ext = max(abs(np.dot(k_hat, b)) for b in recip_b)
        if not ext > 1e-6:
            ext = 2 * np.pi / np.linalg.norm(a1)
    else:
        ext = 2 * np.pi / lat_param
    k_max = bz_coverage * ext
    mags = (np.linspace(0, k_max, n_k, dtype=F32) if n_k > 1
            else np.array([0.0 if np.isclose(k_max, 0) else k_max], F32))
    return mags, np.outer(mags, k_hat).astype(F32)


def ised(positions, velocities, types, dt_ps, k_hat, k_mags, k_vecs, k_target, w_target,
         groups: List[np.ndarray], rescale_factor=1.0, n_frames=100, use_displacements=False) -> Dict:
    """Inverse projection (reference: sed_calculator.py:436-534) for already-resolved ``groups``
    and an already-built k-path; returns the frames that the reference would dump."""
    avg = mean_positions(positions)
    n_atoms = positions.shape[1]
    wig = np.zeros((n_frames, n_atoms, 3), F32)
    tau = np.linspace(0, 2 * np.pi, n_frames, endpoint=False)
    xproj = np.dot(avg, k_hat)
    k_idx = int(np.argmin(np.abs(k_mags - k_target)))
    k_act = k_mags[k_idx]
    auto = isinstance(rescale_factor, str) and rescale_factor.lower() == "auto"
    max_amp, std_sum, n_sum, w_idx = 0.0, 0.0, 0, 0
    for idx in groups:
        if idx.size == 0:
            continue
        res = calculate(positions, velocities, types, dt_ps, k_vecs, basis_atom_indices=idx,
                        summation_mode="coherent", use_displacements=use_displacements)
        w_idx = int(np.argmin(np.abs(res["freqs"] - w_target)))
        for pol in range(3):
            amp = res["sed"][w_idx, k_idx, pol]
            wig[:, idx, pol] += np.real(amp * np.exp(1j * tau[:, None] - 1j * k_act * xproj[idx][None, :]))
        if auto:
            max_amp = max(max_amp, float(np.amax(np.abs(wig[:, idx, :]))))
            std_sum += float(np.std(positions[:, idx, :] - avg[None, idx, :])) * len(idx)
            n_sum += len(idx)
    touched = np.unique(np.concatenate(groups)) if groups else np.array([], int)
    if touched.size:
        if auto:
            if max_amp > 1e-9:
                wig[:, touched, :] /= max_amp
                scale = std_sum / n_sum if n_sum > 0 else 0.0
                if scale > 1e-9:
                    wig[:, touched, :] *= scale
        elif isinstance(rescale_factor, (int, float)):
            wig[:, touched, :] *= rescale_factor
    return dict(frames=avg[None, :, :] + wig, k_idx=k_idx, w_idx=w_idx, k_actual=k_act)


# --------------------------------------------------------------------------- parity metrics

def rel_err_above(new_i: np.ndarray, ref_i: np.ndarray, frac_of_peak: float) -> Dict[str, float]:
    """max / p99 / median of |new-ref|/ref over bins with ref > frac_of_peak * max(ref)."""
    ref64, new64 = ref_i.astype(np.float64), new_i.astype(np.float64)
    mask = ref64 > frac_of_peak * ref64.max()
    if not mask.any():
        return dict(max=0.0, p99=0.0, median=0.0, n=0)
    rel = np.abs(new64[mask] - ref64[mask]) / ref64[mask]
    return dict(max=float(rel.max()), p99=float(np.percentile(rel, 99)),
                median=float(np.median(rel)), n=int(mask.sum()))


def peak_indices(inten: np.ndarray) -> Tuple[Tuple[int, int], np.ndarray]:
    """Global (f,k) argmax and per-k argmax over the non-negative-frequency half."""
    n_f = inten.shape[0]
    g = np.unravel_index(int(np.argmax(inten)), inten.shape)
    per_k = np.argmax(inten[: n_f // 2 + 1], axis=0)
    return (int(g[0]), int(g[1])), per_k


def parity_report(i_new: np.ndarray, i_ref: np.ndarray, i_64: np.ndarray,
                  thresholds=(1e-4, 1e-5, 1e-6)) -> Dict:
    """The three distances of SURVEY.md section 8d: new<->ref, new<->O-64, ref<->O-64."""
    rep: Dict = {}
    for thr in thresholds:
        rep[thr] = dict(new_ref=rel_err_above(i_new, i_ref, thr),
                        new_o64=rel_err_above(i_new, i_64, thr),
                        ref_o64=rel_err_above(i_ref, i_64, thr))
    g_new, pk_new = peak_indices(i_new)
    g_ref, pk_ref = peak_indices(i_ref)
    strong = i_ref[: i_ref.shape[0] // 2 + 1].max(axis=0) > 1e-4 * i_ref.max()
    rep["global_peak_equal"] = g_new == g_ref
    rep["per_k_peak_equal"] = bool(np.array_equal(pk_new[strong], pk_ref[strong]))
    rep["global_peak"] = g_ref
    return rep
