"""``SEDCalculator`` - drop-in for the reference's calculator, backed by the sm_100a kernels.

Two surfaces are exposed on the same class:

* the **source API** that the reference's GUI, CLI and examples call
  (reference: src/psa/core/sed_calculator.py:18-590): ``SEDCalculator(traj, nx, ny, nz,
  use_displacements, dt_ps)``, ``.a1 .. .b3 .recip_vecs_prim .dt_ps .traj``, ``get_k_path``,
  ``get_k_grid``, ``calculate``, ``calculate_chiral_phase``, ``ised``;
* the **README facade** (reference: README.md:83-169): ``calculate_kpath_sed``,
  ``calculate_kgrid_sed``, ``calculate_chiral_sed`` and ``iSEDReconstructor``.

Same argument meaning, same result shapes/dtypes, same exceptions.  Host logic (k-points,
group rules, frequency axis) is NumPy; everything O(n_t x n_atoms) runs on the GPU through
``psa_b200.engine``.  There is no CPU fallback.
"""
from __future__ import annotations

import logging
import weakref
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from . import groups as grp
from . import kspace
from .directions import parse_direction
from .dump import write_lammps_dump
from .engine import DeviceTrajectory, Engine, HostTarget, effective_k_chunk, sed_on_device
from .sed import SED
from .trajectory import Trajectory

logger = logging.getLogger(__name__)

_CHIRAL_AXES = {"x": (1, 2), "y": (0, 2), "z": (0, 1)}   # reference: psa_gui.py:976-982
_ISED_BATCH_BYTES = 16 << 30                              # frames of one batched iSED launch (device memory)


def make_window(spec, n_t: int) -> Optional[np.ndarray]:
    """float32 taper over the frames: ``None`` (rectangular = the reference, sed_calculator.py:83), a name
    ('hann', 'hamming', 'blackman'; periodic form, the right one for spectral estimation) or an array of n_t values."""
    if spec is None:
        return None
    if isinstance(spec, str):
        x = 2.0 * np.pi * np.arange(n_t, dtype=np.float64) / max(n_t, 1)
        name = spec.lower()
        if name in ("rect", "rectangular", "boxcar", "none"):
            return None
        if name == "hann":
            w = 0.5 - 0.5 * np.cos(x)
        elif name == "hamming":
            w = 0.54 - 0.46 * np.cos(x)
        elif name == "blackman":
            w = 0.42 - 0.5 * np.cos(x) + 0.08 * np.cos(2.0 * x)
        else:
            raise ValueError(f"unknown window {spec!r} (use 'hann', 'hamming', 'blackman' or an array of n_frames values)")
        return w.astype(np.float32)
    w = np.asarray(spec, dtype=np.float32).reshape(-1)
    if w.size != n_t:
        raise ValueError(f"window has {w.size} values, the trajectory has {n_t} frames")
    return np.ascontiguousarray(w)


def mass_weights(masses, types: np.ndarray) -> Optional[np.ndarray]:
    """Per-atom float32 sqrt(mass) from a per-atom array or a ``{type: mass}`` mapping (``None`` = unweighted)."""
    if masses is None:
        return None
    n_atoms = len(types)
    if isinstance(masses, dict):
        per_atom = np.empty(n_atoms, np.float64)
        for t in np.unique(types):
            key = int(t)
            if key not in masses:
                raise ValueError(f"masses has no entry for atom type {key}")
            per_atom[types == t] = float(masses[key])
    else:
        per_atom = np.asarray(masses, dtype=np.float64).reshape(-1)
        if per_atom.size != n_atoms:
            raise ValueError(f"masses has {per_atom.size} values, the trajectory has {n_atoms} atoms")
    if not np.all(np.isfinite(per_atom)) or np.any(per_atom <= 0):
        raise ValueError("masses must be positive and finite")
    return np.sqrt(per_atom).astype(np.float32)


class SEDCalculator:
    def __init__(self, traj: Optional[Trajectory] = None, nx: int = 1, ny: int = 1, nz: int = 1,
                 use_displacements: bool = False, dt_ps: Optional[float] = None, *,
                 positions: Optional[np.ndarray] = None, velocities: Optional[np.ndarray] = None,
                 masses=None, types: Optional[np.ndarray] = None,
                 lattice: Optional[np.ndarray] = None, device: Optional[int] = None, window=None):
        """``masses`` (per-atom array or ``{type: mass}``) switches on the README's mass weighting: the projected
        series becomes sqrt(m_a) v_a (float32 product).  ``window`` ('hann' | 'hamming' | 'blackman' | array of
        n_frames values) tapers every projected column before the time FFT.  Both default to ``None`` = exactly the
        shipped reference arithmetic, which has neither (sed_calculator.py:72, 83)."""
        if traj is None:
            traj, dt_ps = self._traj_from_arrays(positions, velocities, types, lattice, dt_ps), None
        if not (nx > 0 and ny > 0 and nz > 0):
            raise ValueError("System dimensions (nx, ny, nz) must be positive.")
        self.traj = traj
        self.use_displacements = use_displacements

        if dt_ps is not None:
            logger.warning("Explicitly providing dt_ps to SEDCalculator is deprecated. The provided dt_ps will "
                           "override the Trajectory's dt_ps.")
            self.dt_ps = dt_ps
        elif getattr(traj, "dt_ps", None) is not None:
            self.dt_ps = traj.dt_ps
        else:
            raise ValueError("Timestep dt_ps not found in Trajectory object and not provided to SEDCalculator.")
        if self.dt_ps <= 0:
            raise ValueError("Timestep dt_ps must be positive.")

        lattice_obj = kspace.Lattice.from_box(traj.box_matrix, nx, ny, nz)
        self._lattice = lattice_obj
        self.a1, self.a2, self.a3 = lattice_obj.a1, lattice_obj.a2, lattice_obj.a3
        self.b1, self.b2, self.b3 = lattice_obj.b1, lattice_obj.b2, lattice_obj.b3
        self.recip_vecs_prim = lattice_obj.recip_vecs_prim

        self._weight = mass_weights(masses, np.asarray(traj.types))
        self._window = make_window(window, traj.n_frames)
        self._device_index = device
        self._engine: Optional[Engine] = None
        self._dev_traj: Optional[DeviceTrajectory] = None

    # ------------------------------------------------------------------ construction helpers
    @staticmethod
    def _traj_from_arrays(positions, velocities, types, lattice, dt_ps) -> Trajectory:
        if positions is None or velocities is None or types is None or lattice is None:
            raise ValueError("Pass either a Trajectory or positions=, velocities=, types= and lattice=.")
        if dt_ps is None:
            raise ValueError("Timestep dt_ps not found in Trajectory object and not provided to SEDCalculator.")
        box = np.asarray(lattice, dtype=np.float32)
        return Trajectory(positions=np.asarray(positions), velocities=np.asarray(velocities),
                          types=np.asarray(types), timesteps=np.arange(len(positions)), box_matrix=box,
                          box_lengths=np.array([box[0, 0], box[1, 1], box[2, 2]], np.float32),
                          box_tilts=np.array([box[0, 1], box[0, 2], box[1, 2]], np.float32), dt_ps=float(dt_ps))

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(self._device_index)
        return self._engine

    @property
    def device_trajectory(self) -> DeviceTrajectory:
        if self._dev_traj is None:
            self._dev_traj = DeviceTrajectory(self.engine, self.traj.positions, self.traj.velocities,
                                              weight=self._weight, window=self._window)
        return self._dev_traj

    def release_device_memory(self) -> None:
        """Drop every device buffer (the next call re-uploads the trajectory).  Call this after modifying the
        trajectory's arrays in place: uploads, mean positions and digit planes are cached per calculator."""
        self._dev_traj = None

    # ------------------------------------------------------------------ k-space (host)
    def get_k_path(self, direction_spec, bz_coverage: float, n_k: int, lat_param: Optional[float] = None
                   ) -> Tuple[np.ndarray, np.ndarray]:
        return kspace.k_path(self._lattice, direction_spec, bz_coverage, n_k, lat_param)

    def get_k_grid(self, plane: str, k_range_x: Tuple[float, float], k_range_y: Tuple[float, float],
                   n_kx: int, n_ky: int, k_fixed_val: float = 0.0):
        return kspace.k_grid(plane, k_range_x, k_range_y, n_kx, n_ky, k_fixed_val)

    # ------------------------------------------------------------------ SED
    def calculate(self, k_points_mags: np.ndarray, k_vectors_3d: np.ndarray,
                  basis_atom_indices=None, basis_atom_types=None, summation_mode: str = "coherent",
                  k_grid_shape: Optional[Tuple[int, int]] = None, k_chunk_size: int = 500) -> SED:
        if summation_mode not in ("coherent", "incoherent"):
            raise ValueError(f"summation_mode must be 'coherent' or 'incoherent', got {summation_mode}")
        n_t, n_atoms = self.traj.n_frames, self.traj.n_atoms
        if n_t == 0 or n_atoms == 0:
            logger.warning("Cannot calculate SED: 0 frames or 0 atoms.")
            return SED(np.array([], dtype=np.complex64).reshape(0, 0, 3), np.array([], dtype=np.float32),
                       k_points_mags, k_vectors_3d, k_grid_shape=k_grid_shape, is_complex=True, phase=None)

        sed_host, complex_out, groups = self._calculate_device(k_vectors_3d, basis_atom_indices, basis_atom_types,
                                                               summation_mode, k_chunk_size, to_host=True)
        freqs = np.fft.fftfreq(n_t, d=self.dt_ps)
        return SED(sed_host, freqs, k_points_mags, k_vectors_3d, k_grid_shape=k_grid_shape,
                   is_complex=complex_out, phase=None, context=self._context(groups))

    def _calculate_device(self, k_vectors_3d, basis_atom_indices, basis_atom_types, summation_mode, k_chunk_size=500,
                          to_host: bool = False):
        """Device-resident result of ``calculate`` (what ``bench.py`` times as the kernel-only path), or with
        ``to_host`` the host array.  A k-set longer than one chunk is streamed: each chunk's spectra go to
        pinned host memory while the next chunk is computed, and nothing result-sized stays on the device."""
        groups = grp.resolve_sed_groups(self.traj.types, self.traj.n_atoms, basis_atom_indices,
                                        basis_atom_types, summation_mode)
        complex_out, proj_groups = grp.plan_sed_groups(groups, summation_mode)
        k_vecs = np.ascontiguousarray(np.asarray(k_vectors_3d, dtype=np.float32).reshape(-1, 3))
        k_chunk = max(1, int(k_chunk_size))
        n_t, n_k = self.traj.n_frames, k_vecs.shape[0]
        with torch.cuda.device(self.engine.device):
            if to_host and n_k > effective_k_chunk(k_chunk, n_k):
                host = torch.empty((n_t, n_k, 3) if complex_out else (n_t, n_k),
                                   dtype=torch.complex64 if complex_out else torch.float32, pin_memory=True)
                sed_on_device(self.device_trajectory, k_vecs, proj_groups, complex_out, self.use_displacements,
                              k_chunk=k_chunk, host_out=HostTarget(host.data_ptr(), n_k, 0, n_t, host))
                self.engine.copy_stream.synchronize()
                return host.numpy(), complex_out, groups
            out = sed_on_device(self.device_trajectory, k_vecs, proj_groups, complex_out,
                                self.use_displacements, k_chunk=k_chunk)
            return (self._to_host(out) if to_host else out), complex_out, groups

    def calculate_intensity(self, k_points_mags: np.ndarray, k_vectors_3d: np.ndarray, basis_atom_indices=None,
                            basis_atom_types=None, summation_mode: str = "coherent",
                            k_grid_shape: Optional[Tuple[int, int]] = None, k_chunk_size: int = 500,
                            max_freq: Optional[float] = None, intensity_scale: str = "linear",
                            vmin_percentile: Optional[float] = None, vmax_percentile: Optional[float] = None,
                            global_range: bool = False) -> SED:
        """The heat-map the plotter and the GUI reduce a result to (reference: sed_plotter.py:127-130,
        psa_gui.py:2196-2214, 2424-2441), produced on the device: ``sum_pol |S|^2`` as float32
        ``(n_f, n_k)`` - for a coherent selection it comes straight out of the FFT kernel's |.|^2 epilogue,
        the complex spectra are never stored - cropped to ``0 <= f <= max_freq`` when given.  Only that
        array crosses PCIe (C4: 0.16 GB instead of 3.9 GB).  ``is_complex`` is False; ``freqs`` matches
        the rows.

        ``intensity_scale`` ('linear' | 'log' | 'sqrt' | 'dsqrt', reference: sed_plotter.py:160-181), the percentile
        colour limits (``np.percentile`` of the finite values, sed_plotter.py:211-215) and ``global_range`` (``nanmin`` /
        ``nanmax`` over every slice, psa_gui.py:2424-2441) are evaluated on the device as well; they land in
        ``result.context['stats']`` (``vmin``, ``vmax``, ``global_min``, ``global_max``)."""
        from . import consumers
        if summation_mode not in ("coherent", "incoherent"):
            raise ValueError(f"summation_mode must be 'coherent' or 'incoherent', got {summation_mode}")
        n_t = self.traj.n_frames
        freqs = np.fft.fftfreq(n_t, d=self.dt_ps) if n_t else np.array([], dtype=np.float64)
        n_rows = n_t if max_freq is None else int(np.count_nonzero((freqs >= 0) & (freqs <= max_freq)))
        if n_t == 0 or self.traj.n_atoms == 0:
            return SED(np.zeros((0, 0), np.float32), freqs, k_points_mags, k_vectors_3d, k_grid_shape=k_grid_shape,
                       is_complex=False, phase=None)
        groups = grp.resolve_sed_groups(self.traj.types, self.traj.n_atoms, basis_atom_indices,
                                        basis_atom_types, summation_mode)
        _, proj_groups = grp.plan_sed_groups(groups, summation_mode)
        k_vecs = np.ascontiguousarray(np.asarray(k_vectors_3d, dtype=np.float32).reshape(-1, 3))
        k_chunk, n_k = max(1, int(k_chunk_size)), k_vecs.shape[0]
        if str(intensity_scale).lower() not in consumers.SCALE_MODES:
            raise ValueError(f"intensity scale must be one of {sorted(consumers.SCALE_MODES)}, got {intensity_scale!r}")
        on_device = (str(intensity_scale).lower() != "linear" or vmin_percentile is not None
                     or vmax_percentile is not None or global_range)
        stats = None
        with torch.cuda.device(self.engine.device):
            if on_device:                                   # the whole (cropped) map stays on the device for the reductions
                out = sed_on_device(self.device_trajectory, k_vecs, proj_groups, False, self.use_displacements,
                                    k_chunk=k_chunk)
                crop = out[:n_rows]
                consumers.scale_intensity(self.engine, crop, intensity_scale)
                stats = consumers.intensity_stats(self.engine, crop, vmin_percentile, vmax_percentile)
                inten = self._to_host(crop)
            elif n_k > effective_k_chunk(k_chunk, n_k):
                host = torch.empty((n_rows, n_k), dtype=torch.float32, pin_memory=True)
                sed_on_device(self.device_trajectory, k_vecs, proj_groups, False, self.use_displacements,
                              k_chunk=k_chunk, host_out=HostTarget(host.data_ptr(), n_k, 0, n_rows, host))
                self.engine.copy_stream.synchronize()
                inten = host.numpy()
            else:
                out = sed_on_device(self.device_trajectory, k_vecs, proj_groups, False, self.use_displacements,
                                    k_chunk=k_chunk)
                inten = self._to_host(out[:n_rows])
        ctx = self._context(groups)
        if stats is not None:
            ctx["stats"] = stats
            ctx["intensity_scale"] = str(intensity_scale).lower()
        return SED(inten, freqs[:n_rows], k_points_mags, k_vectors_3d, k_grid_shape=k_grid_shape,
                   is_complex=False, phase=None, context=ctx)

    def _to_host(self, dev: torch.Tensor) -> np.ndarray:
        host = torch.empty(dev.shape, dtype=dev.dtype, pin_memory=True)
        host.copy_(dev, non_blocking=True)
        torch.cuda.current_stream(dev.device).synchronize()
        return host.numpy()

    def _context(self, groups: List[np.ndarray]) -> Dict:
        """What ``iSEDReconstructor(result)`` needs: geometry plus a WEAK reference to this calculator (a result must
        not keep the trajectory's device buffers alive, and stays picklable - ``SED.__getstate__`` drops the reference)."""
        return dict(calculator=weakref.ref(self), groups=groups, types=self.traj.types, box_matrix=self.traj.box_matrix)

    # ------------------------------------------------------------------ chirality
    def calculate_chiral_phase(self, Z1: np.ndarray, Z2: np.ndarray, angle_range_opt: str = "C") -> np.ndarray:
        if Z1.shape != Z2.shape:
            raise ValueError("Z1 and Z2 shapes must match for chiral phase.")
        if Z1.size == 0:
            return np.array([], dtype=np.float32).reshape(Z1.shape)
        if angle_range_opt not in ("A", "B", "C"):
            logger.warning("Unknown angle_range_opt '%s'. Angle=0.", angle_range_opt)
            return np.zeros(Z1.shape, dtype=np.float32)
        eng = self.engine
        with torch.cuda.device(eng.device):
            z1 = torch.from_numpy(np.ascontiguousarray(Z1, dtype=np.complex64)).to(eng.device)
            z2 = torch.from_numpy(np.ascontiguousarray(Z2, dtype=np.complex64)).to(eng.device)
            out = eng.empty(Z1.shape, torch.float32)
            eng.chiral_phase(z1, z2, z1.numel(), 1, 1, angle_range_opt, out)
            return self._to_host(out)

    def _chiral_phase_of_result(self, sed_dev: torch.Tensor, pair: Tuple[int, int]) -> torch.Tensor:
        """Phase of two polarisation planes of a device-resident (n_f, n_k, 3) result, no host round trip."""
        eng = self.engine
        n_f, n_k, _ = sed_dev.shape
        out = eng.empty((n_f, n_k), torch.float32)
        flat = torch.view_as_real(sed_dev).view(-1, 2)
        eng.chiral_phase(flat[pair[0]:], flat[pair[1]:], n_f * n_k, 3, 3, "C", out)
        return out

    # ------------------------------------------------------------------ README facade (reference: README.md:83-169)
    def calculate_kpath_sed(self, direction, bz_coverage: float = 1.0, n_k: int = 100, basis_atom_types=None,
                            summation_mode: str = "coherent", basis_atom_indices=None, lat_param=None) -> SED:
        k_mags, k_vecs = self.get_k_path(direction, bz_coverage, n_k, lat_param)
        res = self.calculate(k_mags, k_vecs, basis_atom_indices=basis_atom_indices,
                             basis_atom_types=basis_atom_types, summation_mode=summation_mode)
        res.context["k_hat"] = parse_direction(direction)
        return res

    def calculate_kgrid_sed(self, plane: str = "xy", k_ranges: Sequence[float] = (-1, 1, -1, 1), n_kx: int = 50,
                            n_ky: int = 50, k_fixed: float = 0.0, basis_atom_types=None,
                            summation_mode: str = "coherent", basis_atom_indices=None) -> SED:
        k_mags, k_vecs, shape = self.get_k_grid(plane, (k_ranges[0], k_ranges[1]), (k_ranges[2], k_ranges[3]),
                                                n_kx, n_ky, k_fixed)
        return self.calculate(k_mags, k_vecs, basis_atom_indices=basis_atom_indices,
                              basis_atom_types=basis_atom_types, summation_mode=summation_mode, k_grid_shape=shape)

    def calculate_chiral_sed(self, direction, bz_coverage: float = 1.0, n_k: int = 100, chiral_axis: str = "z",
                             basis_atom_types=None, basis_atom_indices=None) -> SED:
        """Coherent SED plus the folded phase between the two polarisations normal to ``chiral_axis``
        (the GUI's chirality option, reference: psa_gui.py:957-991)."""
        pair = _CHIRAL_AXES.get(str(chiral_axis).lower(), _CHIRAL_AXES["z"])
        k_mags, k_vecs = self.get_k_path(direction, bz_coverage, n_k)
        n_t = self.traj.n_frames
        out_dev, complex_out, groups = self._calculate_device(k_vecs, basis_atom_indices, basis_atom_types, "coherent")
        with torch.cuda.device(self.engine.device):
            phase_dev = self._chiral_phase_of_result(out_dev, pair)
            sed_host, phase_host = self._to_host(out_dev), self._to_host(phase_dev)
        ctx = self._context(groups)
        ctx["k_hat"] = parse_direction(direction)
        return SED(sed_host, np.fft.fftfreq(n_t, d=self.dt_ps), k_mags, k_vecs, k_grid_shape=None,
                   phase=phase_host, is_complex=complex_out, context=ctx)

    # ------------------------------------------------------------------ iSED
    def reconstruct(self, k_dir_spec, targets: Sequence[Tuple[float, float]], char_len_k_path: Optional[float],
                    nk_on_path: int = 100, bz_cov_ised: float = 1.0, basis_atom_idx_ised=None,
                    basis_atom_types_ised=None, rescale_factor: Union[str, float] = 1.0,
                    n_recon_frames: int = 100, keep_on_device: bool = False) -> List[Dict]:
        """Batched inverse projection: one entry per ``(k_target, w_target)`` with the reconstructed
        frames ``(n_recon_frames, n_atoms, 3)`` float32 plus the matched indices.  The projection of all
        distinct matched k-points is done in one pass per atom group instead of one full SED per group
        per call (reference: sed_calculator.py:451-499).  ``keep_on_device`` leaves ``frames`` as CUDA
        tensors (for device-side consumers; the host copy of many large frame sets is bound by the
        first-touch cost of fresh host memory, not by the GPU)."""
        traj = self.traj
        n_atoms = traj.n_atoms
        k_hat = parse_direction(k_dir_spec)
        recon_groups = grp.resolve_ised_groups(traj.types, n_atoms, basis_atom_idx_ised, basis_atom_types_ised)
        if not recon_groups:
            logger.error("iSED: No atom groups for reconstruction. Aborting.")
            return []
        k_mags, k_vecs = self.get_k_path(k_hat, bz_cov_ised, nk_on_path, lat_param=char_len_k_path)
        freqs = np.fft.fftfreq(traj.n_frames, d=self.dt_ps)
        k_idx = [int(np.argmin(np.abs(k_mags - kt))) for kt, _ in targets]
        w_idx = [int(np.argmin(np.abs(freqs - wt))) for _, wt in targets]
        uniq_k = sorted(set(k_idx))
        col_of = {k: i for i, k in enumerate(uniq_k)}
        auto = isinstance(rescale_factor, str) and rescale_factor.lower() == "auto"

        eng, dtraj = self.engine, self.device_trajectory
        results: List[Dict] = []
        n_pts, n_grp, n_fr = len(targets), len(recon_groups), int(n_recon_frames)
        if n_pts == 0:
            return results
        # groups of every atom as a CSR list, in the reference's group-loop order; a group counts once per atom
        # (the reference's fancy-indexed `+=` ignores duplicate indices, sed_calculator.py:499)
        members = [np.unique(g) for g in recon_groups]
        counts = np.zeros(n_atoms + 1, np.int64)
        for m in members:
            counts[m + 1] += 1
        member_off = np.cumsum(counts)
        member_grp = np.zeros(max(1, int(member_off[-1])), np.int32)
        fill = member_off[:-1].copy()
        for gi, m in enumerate(members):
            member_grp[fill[m]] = gi
            fill[m] += 1
        with torch.cuda.device(eng.device):
            # amplitudes S_g[w, k, pol] of every group at the matched bins only
            w_dev = eng.upload_small(np.asarray(w_idx, np.int32))
            c_dev = eng.upload_small(np.asarray([col_of[k] for k in k_idx], np.int32))
            amp = eng.empty((n_pts, n_grp, 3), torch.complex64)
            for gi, g in enumerate(recon_groups):
                sed_g = sed_on_device(dtraj, k_vecs[uniq_k], [g], True, self.use_displacements)
                eng._run("psa_gather_bins", 1, sed_g.data_ptr(), len(uniq_k), w_dev.data_ptr(), c_dev.data_ptr(), n_pts,
                         n_grp * 3, amp.data_ptr() + gi * 3 * 8, eng.stream())
                del sed_g
            mean = dtraj.mean
            khat_dev = eng.upload_small(np.ascontiguousarray(k_hat, np.float32))
            kact = np.asarray([k_mags[k] for k in k_idx], np.float32)
            kact_dev = eng.upload_small(kact)
            off_dev = eng.upload_small(member_off.astype(np.int32))
            grp_dev = eng.upload_small(member_grp)
            batch_args = (mean.data_ptr(), khat_dev.data_ptr())

            div, mul = np.ones(n_pts, np.float32), np.ones(n_pts, np.float32)
            if auto:                                                           # reference: sed_calculator.py:502-524
                num, den = 0.0, 0
                mom = eng.empty((2,), torch.float64)
                for g in recon_groups:
                    whole = g.size == n_atoms and np.array_equal(g, np.arange(n_atoms))   # whole rows: vector loads
                    idx_dev = None if whole else eng.upload_small(np.ascontiguousarray(g, np.int32))
                    eng._run("psa_disp_moments", 1, dtraj.positions.data_ptr(), mean.data_ptr(),
                             None if whole else idx_dev.data_ptr(),
                             traj.n_frames, n_atoms, int(g.size), mom.data_ptr(), eng.stream())
                    s1, s2 = mom.cpu().tolist()
                    n_el = traj.n_frames * int(g.size) * 3
                    var = max(s2 / n_el - (s1 / n_el) ** 2, 0.0)
                    num += float(np.sqrt(var)) * int(g.size)
                    den += int(g.size)
                std_scale = num / den if den > 0 else 0.0
                wmax = eng.empty((n_pts,), torch.float32)
                eng._run("psa_ised_absmax", 1, *batch_args, kact_dev.data_ptr(), amp.data_ptr(), off_dev.data_ptr(),
                         grp_dev.data_ptr(), n_grp, n_atoms, n_fr, n_pts, wmax.data_ptr(), eng.stream())
                wmax_host = self._to_host(wmax)
                for p in range(n_pts):
                    if wmax_host[p] > 1e-9:
                        div[p] = wmax_host[p]
                        if std_scale > 1e-9:
                            mul[p] = np.float32(std_scale)
                    else:
                        logger.warning("iSED: Max wiggle amp near zero. Auto-rescaling ineffective.")
            elif isinstance(rescale_factor, (int, float)):
                mul[:] = np.float32(rescale_factor)
            div_dev, mul_dev = eng.upload_small(div), eng.upload_small(mul)

            per_point = n_fr * n_atoms * 12
            batch = max(1, min(n_pts, _ISED_BATCH_BYTES // max(per_point, 1)))
            for p0 in range(0, n_pts, batch):
                p1 = min(n_pts, p0 + batch)
                frames = eng.empty((p1 - p0, n_fr, n_atoms, 3), torch.float32)
                eng._run("psa_ised_frames", 1, *batch_args, kact_dev.data_ptr() + 4 * p0,
                         amp.data_ptr() + p0 * n_grp * 3 * 8, off_dev.data_ptr(), grp_dev.data_ptr(), n_grp, n_atoms, n_fr,
                         p1 - p0, div_dev.data_ptr() + 4 * p0, mul_dev.data_ptr() + 4 * p0, frames.data_ptr(), eng.stream())
                if keep_on_device:
                    host_frames = frames
                elif n_pts == 1:
                    host_frames = self._to_host(frames)
                else:       # many frame sets: one plain host array (a pinned buffer of this size costs more than the copy)
                    host_frames = np.empty(tuple(frames.shape), np.float32)
                    torch.from_numpy(host_frames).copy_(frames)
                for ti in range(p0, p1):
                    kt, wt = targets[ti]
                    results.append(dict(frames=host_frames[ti - p0], k_index=k_idx[ti], w_index=w_idx[ti],
                                        k_actual=float(kact[ti]), w_actual=float(freqs[w_idx[ti]]), k_target=kt, w_target=wt))
        return results

    def ised(self, k_dir_spec, k_target: float, w_target: float, char_len_k_path: float, nk_on_path: int = 100,
             bz_cov_ised: float = 1.0, basis_atom_idx_ised: Optional[List[int]] = None,
             basis_atom_types_ised: Optional[List[int]] = None, rescale_factor: Union[str, float] = 1.0,
             n_recon_frames: int = 100, dump_filepath: str = "iSED_reconstruction.dump",
             plot_dir_ised: Optional[Path] = None, plot_max_freq: Optional[float] = None,
             plot_theme: str = "light") -> None:
        """Reconstruct the motion of one (k, omega) mode and write it as a LAMMPS dump
        (reference: sed_calculator.py:373-538).  Plotting of the input spectrum is left to the
        reference's ``SEDPlotter`` (out of scope here); a request for it is logged and skipped."""
        logger.info("Starting iSED reconstruction.")
        res = self.reconstruct(k_dir_spec, [(k_target, w_target)], char_len_k_path, nk_on_path, bz_cov_ised,
                               basis_atom_idx_ised, basis_atom_types_ised, rescale_factor, n_recon_frames)
        if not res:
            return
        write_lammps_dump(dump_filepath, res[0]["frames"], self.traj.types.astype(int), self.traj.box_matrix)
        logger.info("iSED reconstruction saved: %s", dump_filepath)
        if plot_dir_ised:
            logger.warning("iSED input-spectrum plot requested; plotting is not part of psa_b200 "
                           "(use psa.visualization.SEDPlotter on calculate_kpath_sed's result).")


class iSEDReconstructor:
    """README facade (reference: README.md:148-167) over :meth:`SEDCalculator.reconstruct`."""

    def __init__(self, sed_result: SED):
        ctx = getattr(sed_result, "context", None)
        if not ctx or "calculator" not in ctx:
            raise ValueError("iSEDReconstructor needs a result produced by psa_b200.SEDCalculator "
                             "(it carries the trajectory geometry).")
        self._sed = sed_result
        calc = ctx["calculator"]
        self._calc: SEDCalculator = calc() if isinstance(calc, weakref.ref) else calc
        if self._calc is None:
            raise ValueError("the SEDCalculator that produced this result no longer exists (iSED needs its trajectory)")
        self._k_hat = ctx.get("k_hat")
        if self._k_hat is None:
            vecs = np.asarray(sed_result.k_vectors, dtype=np.float32)
            norms = np.linalg.norm(vecs, axis=1)
            if not (norms > 0).any():
                raise ValueError("cannot infer the k-path direction from the result")
            self._k_hat = vecs[int(np.argmax(norms))] / norms.max()
        self._groups = ctx.get("groups")

    def reconstruct_motion(self, k_target: float, omega_target: float, n_frames: int = 100,
                           rescale_factor: Union[str, float] = 1.0) -> np.ndarray:
        k_mags = np.asarray(self._sed.k_points)
        if k_mags.size < 1:
            raise ValueError("iSED needs a k-path result (k_points is empty for k-grids).")
        k_max = float(k_mags[-1])
        lat_param = 2 * np.pi / k_max if k_max > 0 else None      # reproduces the result's own k-path
        idx_groups = [list(map(int, g)) for g in self._groups] if self._groups else None
        out = self._calc.reconstruct(self._k_hat, [(k_target, omega_target)], lat_param, nk_on_path=len(k_mags),
                                     bz_cov_ised=1.0, basis_atom_idx_ised=idx_groups,
                                     rescale_factor=rescale_factor, n_recon_frames=n_frames)
        return out[0]["frames"]

    def save_trajectory(self, motion: np.ndarray, path: str, format: str = "lammps") -> None:
        if format != "lammps":
            raise ValueError("only format='lammps' is supported")
        write_lammps_dump(path, motion, self._calc.traj.types.astype(int), self._calc.traj.box_matrix)
