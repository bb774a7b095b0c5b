"""Seeded synthetic MD trajectories for parity tests and for ``bench.py``.

There is no network and no real trajectory in the image, so every workload
(the five BASELINE.json configs and the small test cases) is generated here:
an ideal lattice plus a handful of plane-wave phonon-like modes plus white
noise.  Displacements are ``u = A e cos(q.r0 - w t + phi)``, velocities are the
analytic time derivative, both get independent Gaussian noise.  Mode wave
vectors are commensurate with the supercell, so the SED shows sharp peaks at
known (k, omega).

Frames are produced in fixed blocks of ``BLOCK`` frames, each with its own
counter-derived RNG stream, so any time slice ``[t0, t1)`` can be generated
independently (the 25 GB config streams through pinned memory chunk by chunk)
and gives bit-identical data regardless of how it is chunked.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from psa_b200.trajectory import Trajectory

BLOCK = 256  # frames per RNG block


@dataclass
class SyntheticSpec:
    name: str
    r0: np.ndarray                 # (n_atoms, 3) float64 ideal sites
    types: np.ndarray              # (n_atoms,) int32
    box_matrix: np.ndarray         # (3,3) float32, rows = supercell vectors
    cells: Tuple[int, int, int]    # (nx, ny, nz) for SEDCalculator
    n_frames: int
    dt_ps: float
    seed: int
    # modes: q (M,3) rad/A, freq (M,) THz, phi (M,), pol_re/pol_im (M,3) (complex polarisation), amp (M,)
    q: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))
    freq: np.ndarray = field(default_factory=lambda: np.zeros(0))
    phi: np.ndarray = field(default_factory=lambda: np.zeros(0))
    pol_re: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))
    pol_im: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))
    amp: np.ndarray = field(default_factory=lambda: np.zeros(0))
    sigma_u: float = 0.005
    sigma_v: float = 0.25

    @property
    def n_atoms(self) -> int:
        return self.r0.shape[0]

    # ---- spatial factors: u[t,a,:] = sum_m Re( amp_m * pol_m * exp(i(q_m.r_a + phi_m - w_m t)) )
    def _spatial(self) -> Tuple[np.ndarray, np.ndarray]:
        """W_cos, W_sin of shape (M, n_atoms*3): u = cos(wt) @ W_cos + sin(wt) @ W_sin."""
        alpha = self.r0 @ self.q.T + self.phi[None, :]            # (n_a, M)
        ca, sa = np.cos(alpha).T, np.sin(alpha).T                 # (M, n_a)
        # Re(p e^{i alpha} e^{-i w t}) = cos(wt) Re(p e^{i alpha}) + sin(wt) Im(p e^{i alpha})
        re = self.pol_re[:, None, :] * ca[:, :, None] - self.pol_im[:, None, :] * sa[:, :, None]
        im = self.pol_re[:, None, :] * sa[:, :, None] + self.pol_im[:, None, :] * ca[:, :, None]
        a = self.amp[:, None, None]
        m = self.q.shape[0]
        return (a * re).reshape(m, -1), (a * im).reshape(m, -1)

    def frames(self, t0: int, t1: int, out_pos: Optional[np.ndarray] = None,
               out_vel: Optional[np.ndarray] = None, threads: int = 8
               ) -> Tuple[np.ndarray, np.ndarray]:
        """Positions and velocities (float32, shape (t1-t0, n_atoms, 3)) of frames [t0, t1).

        ``t0`` must be a multiple of ``BLOCK`` so chunked generation is reproducible.
        """
        if t0 % BLOCK:
            raise ValueError(f"t0 must be a multiple of {BLOCK}")
        n_a = self.n_atoms
        shape = (t1 - t0, n_a, 3)
        pos = out_pos if out_pos is not None else np.empty(shape, np.float32)
        vel = out_vel if out_vel is not None else np.empty(shape, np.float32)
        w_cos, w_sin = self._spatial()
        omega = 2 * np.pi * self.freq                              # rad/ps
        r0_flat = self.r0.reshape(1, -1)

        def one_block(b0: int) -> None:
            b1 = min(b0 + BLOCK, t1)
            rng = np.random.default_rng([self.seed, b0 // BLOCK])
            t = (np.arange(b0, b1) * self.dt_ps)[:, None] * omega[None, :]   # (nb, M)
            ct, st = np.cos(t), np.sin(t)
            nb = b1 - b0
            u = ct @ w_cos + st @ w_sin                                       # (nb, 3 n_a)
            # d/dt [cos(wt) Wc + sin(wt) Ws] = w (-sin(wt) Wc + cos(wt) Ws)
            v = (-st * omega) @ w_cos + (ct * omega) @ w_sin
            noise = rng.standard_normal((2, nb, 3 * n_a), dtype=np.float32)
            pos[b0 - t0:b1 - t0] = (r0_flat + u + self.sigma_u * noise[0]).reshape(nb, n_a, 3)
            vel[b0 - t0:b1 - t0] = (v + self.sigma_v * noise[1]).reshape(nb, n_a, 3)

        starts = list(range(t0, t1, BLOCK))
        if threads > 1 and len(starts) > 1:
            with ThreadPoolExecutor(max_workers=threads) as pool:
                list(pool.map(one_block, starts))
        else:
            for b0 in starts:
                one_block(b0)
        return pos, vel

    def trajectory(self, threads: int = 8) -> Trajectory:
        pos, vel = self.frames(0, self.n_frames, threads=threads)
        return self.wrap(pos, vel)

    def wrap(self, pos: np.ndarray, vel: np.ndarray) -> Trajectory:
        box = self.box_matrix
        return Trajectory(positions=pos, velocities=vel, types=self.types,
                          timesteps=np.arange(pos.shape[0]), box_matrix=box,
                          box_lengths=np.array([box[0, 0], box[1, 1], box[2, 2]], np.float32),
                          box_tilts=np.array([box[1, 0], box[2, 0], box[2, 1]], np.float32),
                          dt_ps=self.dt_ps)


# ------------------------------------------------------------------ lattices

SI_A = 5.431
_SI_FCC = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]], float)


def si_diamond(n: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """8-atom cubic cell repeated n^3 times; type 1 = fcc sublattice, type 2 = the (1/4,1/4,1/4) one."""
    cell = np.concatenate([_SI_FCC, _SI_FCC + 0.25])
    cell_types = np.array([1] * 4 + [2] * 4, np.int32)
    ijk = np.stack(np.meshgrid(*(np.arange(n),) * 3, indexing="ij"), -1).reshape(-1, 3)
    r0 = ((ijk[:, None, :] + cell[None, :, :]) * SI_A).reshape(-1, 3)
    types = np.tile(cell_types, len(ijk))
    box = (np.eye(3) * n * SI_A).astype(np.float32)
    return r0, types, box


GR_A = 2.46


def graphene(n: int, height: float = 20.0) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """n x n hexagonal cells, 2 atoms each (type 1 / type 2), sheet at z = height/2."""
    a1 = GR_A * np.array([1.0, 0.0, 0.0])
    a2 = GR_A * np.array([0.5, np.sqrt(3) / 2, 0.0])
    basis = np.stack([np.zeros(3), (a1 + a2) / 3])
    ij = np.stack(np.meshgrid(np.arange(n), np.arange(n), indexing="ij"), -1).reshape(-1, 2)
    origin = ij[:, :1] * a1[None, :] + ij[:, 1:] * a2[None, :]
    r0 = (origin[:, None, :] + basis[None, :, :]).reshape(-1, 3)
    r0[:, 2] = height / 2
    types = np.tile(np.array([1, 2], np.int32), len(ij))
    box = np.stack([n * a1, n * a2, [0.0, 0.0, height]]).astype(np.float32)
    return r0, types, box


# ------------------------------------------------------------------ mode sets

def _recip(box: np.ndarray) -> np.ndarray:
    """Rows = reciprocal vectors of the SUPERCELL (2 pi convention)."""
    return 2 * np.pi * np.linalg.inv(box.astype(float)).T


def _add_modes(spec: SyntheticSpec, rng: np.random.Generator, int_q: List[Tuple[int, int, int]],
               circular: bool = False) -> None:
    g = _recip(spec.box_matrix)
    m = len(int_q)
    spec.q = np.array(int_q, float) @ g
    n_t, dt = spec.n_frames, spec.dt_ps
    df = 1.0 / (n_t * dt)
    f = rng.uniform(1.0, 15.0, m)
    on_bin = np.arange(m) % 2 == 0                 # alternate exact-bin and off-bin frequencies
    f[on_bin] = np.round(f[on_bin] / df) * df
    spec.freq = f
    spec.phi = rng.uniform(0, 2 * np.pi, m)
    pol = rng.standard_normal((m, 3))
    pol /= np.linalg.norm(pol, axis=1, keepdims=True)
    spec.pol_re, spec.pol_im = pol, np.zeros((m, 3))
    if circular:                                   # x +- i y: handed in-plane rotation
        hand = np.where(np.arange(m) % 2 == 0, 1.0, -1.0)
        spec.pol_re = np.tile(np.array([1.0, 0.0, 0.0]), (m, 1)) / np.sqrt(2)
        spec.pol_im = np.stack([np.zeros(m), hand, np.zeros(m)], 1) / np.sqrt(2)
    spec.amp = np.full(m, 0.05)


def si_spec(name: str, n_cells: int, n_frames: int, seed: int, n_modes: int = 8) -> SyntheticSpec:
    r0, types, box = si_diamond(n_cells)
    spec = SyntheticSpec(name, r0, types, box, (n_cells,) * 3, n_frames, 0.002, seed)
    rng = np.random.default_rng([seed, 0xC0FFEE])
    n = n_cells
    # commensurate wave vectors: multiples of the supercell reciprocal vectors,
    # on [100], [110], [111] and a few generic points
    cand = [(n // 4, 0, 0), (n // 2, 0, 0), (3 * n // 4, 0, 0), (n // 4, n // 4, 0),
            (n // 2, n // 2, 0), (n // 4, n // 4, n // 4), (n // 2, n // 2, n // 2),
            (1, 0, 0), (1, 2, 0), (2, 1, 3), (0, 3, 1), (3 * n // 4, 3 * n // 4, 0)]
    _add_modes(spec, rng, cand[:max(1, min(n_modes, len(cand)))])
    return spec


def graphene_spec(name: str, n_cells: int, n_frames: int, seed: int, n_modes: int = 8) -> SyntheticSpec:
    r0, types, box = graphene(n_cells)
    spec = SyntheticSpec(name, r0, types, box, (n_cells, n_cells, 1), n_frames, 0.005, seed)
    rng = np.random.default_rng([seed, 0xC0FFEE])
    n = n_cells
    cand = [(n // 5, 0, 0), (n // 2, 0, 0), (n // 10, 0, 0), (n // 5, n // 5, 0), (0, n // 2, 0),
            (1, 0, 0), (2, 1, 0), (3 * n // 10, 0, 0), (n // 4, n // 10, 0), (1, 3, 0)]
    _add_modes(spec, rng, cand[:max(1, min(n_modes, len(cand)))], circular=True)
    return spec


# ------------------------------------------------------------------ BASELINE.json configs

def baseline_config(name: str, n_frames: Optional[int] = None, n_cells: Optional[int] = None) -> Dict:
    """Workload description for BASELINE.json ``configs[i]`` (``c1``..``c5``).

    ``n_frames`` / ``n_cells`` shrink a config for tests; the defaults are the published sizes.
    Returns ``{'spec': SyntheticSpec, 'kind': 'kpath'|'kgrid'|'chiral', ...call arguments...}``.
    """
    name = name.lower()
    if name == "c1":
        spec = si_spec("c1", n_cells or 8, n_frames or 8192, seed=1)
        return dict(spec=spec, kind="kpath", paths=[dict(direction=[1, 0, 0], n_k=100)],
                    bz_coverage=4.0, summation_mode="coherent", basis_atom_types=None)
    if name == "c2":
        spec = si_spec("c2", n_cells or 8, n_frames or 16384, seed=2)
        return dict(spec=spec, kind="kpath", paths=[dict(direction=[1, 1, 0], n_k=200)],
                    bz_coverage=4.0, summation_mode="incoherent", basis_atom_types=[1])
    if name == "c3":
        spec = graphene_spec("c3", n_cells or 50, n_frames or 16384, seed=3)
        return dict(spec=spec, kind="chiral", paths=[dict(direction=[1, 0, 0], n_k=80)],
                    bz_coverage=4.0, summation_mode="coherent", basis_atom_types=None, chiral_axis="z")
    if name == "c4":
        spec = si_spec("c4", n_cells or 12, n_frames or 16384, seed=4)
        return dict(spec=spec, kind="kgrid", plane="xy", k_ranges=(-3.5, 3.5, -3.5, 3.5),
                    n_kx=100, n_ky=100, k_fixed=0.0, summation_mode="coherent", basis_atom_types=None)
    if name == "c5":
        spec = si_spec("c5", n_cells or 20, n_frames or 32768, seed=5, n_modes=12)
        return dict(spec=spec, kind="kpath",
                    paths=[dict(direction=d, n_k=256) for d in ([1, 0, 0], [1, 1, 0], [1, 1, 1])],
                    bz_coverage=4.0, summation_mode="coherent", basis_atom_types=None,
                    ised_points=64, ised_frames=100)
    raise ValueError(f"unknown baseline config {name!r}")
