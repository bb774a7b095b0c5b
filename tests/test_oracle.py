"""The oracle must reproduce the REAL reference's outputs (tests/golden, made by
oracle/make_golden.py) - bit for bit on the float32 path when NumPy/OpenBLAS match
the recorded environment, and to rounding otherwise."""
import numpy as np
import pytest

from oracle import psa_oracle as O


def _same_env(gold) -> bool:
    return np.__version__.split(".")[:2] == eval(str(gold["_env"]))["numpy"].split(".")[:2]


def _check(new, ref, exact):
    assert new.shape == ref.shape and new.dtype == ref.dtype
    if exact:
        np.testing.assert_array_equal(new, ref)
    else:
        scale = np.abs(ref).max()
        np.testing.assert_allclose(new, ref, rtol=0, atol=2e-6 * scale)


SED_CASES = {
    "coh_all_100": ("kpath_100_vecs", {}),
    "coh_all_110": ("kpath_110_vecs", {}),
    "coh_all_111": ("kpath_111_vecs", {}),
    "coh_all_100_chunk5": ("kpath_100_vecs", dict(k_chunk_size=5)),
    "coh_types12": ("kpath_110_vecs", dict(basis_atom_types=[1, 2], summation_mode="coherent")),
    "inc_types12": ("kpath_110_vecs", dict(basis_atom_types=[1, 2], summation_mode="incoherent")),
    "inc_types1": ("kpath_110_vecs", dict(basis_atom_types=[1], summation_mode="incoherent")),
    "inc_types_nested": ("kpath_110_vecs", dict(basis_atom_types=[[1, 2]], summation_mode="incoherent")),
    "inc_types_unknown": ("kpath_100_vecs", dict(basis_atom_types=[7], summation_mode="incoherent")),
    "inc_types_1_and_unknown": ("kpath_100_vecs", dict(basis_atom_types=[1, 7], summation_mode="incoherent")),
    "inc_indices": ("kpath_100_vecs", dict(basis_atom_indices=[[0, 1, 5, 9], [2, 3, 40]],
                                           summation_mode="incoherent")),
    "coh_indices_union": ("kpath_100_vecs", dict(basis_atom_indices=[[0, 1, 5, 9], [2, 3, 5]],
                                                 summation_mode="coherent")),
    "coh_indices_flat_dup": ("kpath_100_vecs", dict(basis_atom_indices=[3, 1, 1, 20])),
    "coh_indices_ndarray": ("kpath_100_vecs", dict(basis_atom_indices=np.array([4, 8, 15, 16, 23, 42]))),
    "inc_all": ("kpath_100_vecs", dict(summation_mode="incoherent")),
    "kgrid_xy": ("kgrid_xy_vecs", {}),
}


@pytest.mark.parametrize("name", sorted(SED_CASES))
def test_calculate_matches_reference(gold_si, name):
    kkey, kw = SED_CASES[name]
    g = gold_si
    res = O.calculate(g["positions"], g["velocities"], g["types"], float(g["dt_ps"]), g[kkey], **kw)
    assert res["is_complex"] == bool(g[f"cplx_{name}"])
    _check(res["sed"], g[f"sed_{name}"], _same_env(g))
    np.testing.assert_array_equal(res["freqs"], g["freqs"])


def test_displacement_mode(gold_si):
    g = gold_si
    res = O.calculate(g["positions"], g["velocities"], g["types"], float(g["dt_ps"]),
                      g["kpath_100_vecs"], use_displacements=True)
    _check(res["sed"], g["sed_disp_coh_all_100"], _same_env(g))
    res = O.calculate(g["positions"], g["velocities"], g["types"], float(g["dt_ps"]),
                      g["kpath_110_vecs"], basis_atom_types=[1, 2], summation_mode="incoherent",
                      use_displacements=True)
    _check(res["sed"], g["sed_disp_inc_types12"], _same_env(g))


def test_odd_frame_count(gold_si):
    g = gold_si
    res = O.calculate(g["positions"][:250], g["velocities"][:250], g["types"], float(g["dt_ps"]),
                      g["kpath_100_vecs"])
    _check(res["sed"], g["sed_odd250_coh_all_100"], _same_env(g))


def test_intensity(gold_si):
    np.testing.assert_array_equal(O.intensity(gold_si["sed_coh_all_100"]), gold_si["intensity_coh_all_100"])


def test_fp64_oracle_is_close_to_reference(gold_si):
    g = gold_si
    r64 = O.calculate_fp64(g["positions"], g["velocities"], g["types"], float(g["dt_ps"]), g["kpath_110_vecs"])
    ref = g["sed_coh_all_110"]
    assert r64["sed"].dtype == np.complex128
    assert np.abs(r64["sed"] - ref).max() < 5e-6 * np.abs(ref).max()


@pytest.mark.parametrize("axis,pair", [("x", (1, 2)), ("y", (0, 2)), ("z", (0, 1))])
def test_chiral_phase_c(gold_gr, axis, pair):
    s = gold_gr["sed_coh"]
    ph = O.chiral_phase(s[:, :, pair[0]], s[:, :, pair[1]], "C")
    _check(ph, gold_gr[f"phase_C_{axis}"], _same_env(gold_gr))
    assert np.abs(ph).max() <= np.pi / 2 + 1e-6


@pytest.mark.parametrize("opt", ["A", "B"])
def test_chiral_phase_ab(gold_gr, opt):
    s = gold_gr["sed_coh"]
    ph = O.chiral_phase(s[:, :, 0], s[:, :, 1], opt)
    np.testing.assert_allclose(ph, gold_gr[f"phase_{opt}_z"], atol=2e-3)   # acos/asin are ill-conditioned at +-1
    assert np.median(np.abs(ph - gold_gr[f"phase_{opt}_z"])) < 1e-6


def test_chiral_phase_shape_mismatch():
    with pytest.raises(ValueError):
        O.chiral_phase(np.zeros((2, 2), np.complex64), np.zeros((2, 3), np.complex64))
    assert O.chiral_phase(np.zeros((0, 4), np.complex64), np.zeros((0, 4), np.complex64)).shape == (0, 4)


def _ised_args(g):
    k_hat = np.array([1, 0, 0], np.float32)
    mags, vecs = O.k_path(None, None, k_hat, 1.0, 9, lat_param=5.431)
    np.testing.assert_array_equal(mags, g["kpath_lat_mags"])
    return k_hat, mags, vecs


def test_ised_types_float(gold_si):
    g = gold_si
    k_hat, mags, vecs = _ised_args(g)
    groups = [np.where(g["types"] == 1)[0], np.where(g["types"] == 2)[0]]
    out = O.ised(g["positions"], g["velocities"], g["types"], float(g["dt_ps"]), k_hat, mags, vecs,
                 float(g["ised_k_target"]), float(g["ised_w_target"]), groups, rescale_factor=0.5, n_frames=8)
    _check(out["frames"], g["ised_types_float"], _same_env(g))


def test_ised_auto(gold_si):
    g = gold_si
    k_hat, mags, vecs = _ised_args(g)
    out = O.ised(g["positions"], g["velocities"], g["types"], float(g["dt_ps"]), k_hat, mags, vecs,
                 float(g["ised_k_target"]), float(g["ised_w_target"]), [np.arange(len(g["types"]))],
                 rescale_factor="auto", n_frames=8)
    _check(out["frames"], g["ised_all_auto"], _same_env(g))
    groups = [np.array([0, 1, 2, 3]), np.array([10, 11, 12])]
    out = O.ised(g["positions"], g["velocities"], g["types"], float(g["dt_ps"]), k_hat, mags, vecs,
                 0.3, float(g["ised_w_target"]) * 0.5, groups, rescale_factor="auto", n_frames=8)
    _check(out["frames"], g["ised_idx_auto"], _same_env(g))


def test_parity_report_shape():
    rng = np.random.default_rng(0)
    i_ref = rng.random((16, 5)).astype(np.float32)
    rep = O.parity_report(i_ref * (1 + 1e-7), i_ref, i_ref.astype(np.float64))
    assert rep["global_peak_equal"] and rep["per_k_peak_equal"]
    assert rep[1e-6]["new_ref"]["max"] < 1e-6


# ------------------------------------------------------------------ randomized differential check (build container only)
def _random_case(seed):
    rng = np.random.default_rng(seed)
    n_t = int(rng.choice([7, 16, 30, 64, 100, 129]))
    n_a = int(rng.integers(3, 40))
    cells = tuple(int(c) for c in rng.integers(1, 4, 3))
    box = np.diag(rng.uniform(8.0, 25.0, 3)).astype(np.float32)
    if seed % 2:
        box[0, 1], box[0, 2], box[1, 2] = rng.uniform(-2, 2, 3).astype(np.float32)     # tilted cell
    pos = (rng.random((n_t, n_a, 3)) * 20 + rng.standard_normal((n_t, n_a, 3)) * 0.05).astype(np.float32)
    vel = rng.standard_normal((n_t, n_a, 3)).astype(np.float32)
    types = rng.integers(1, 4, n_a)
    kw = [dict(), dict(summation_mode="incoherent"), dict(basis_atom_types=[1, 2], summation_mode="incoherent"),
          dict(basis_atom_types=[2, 3], summation_mode="coherent"), dict(basis_atom_types=[[1], [2, 3]], summation_mode="incoherent"),
          dict(basis_atom_indices=[[0, 1], [2]], summation_mode="incoherent"),
          dict(basis_atom_indices=list(map(int, rng.choice(n_a, 3, replace=False))))][seed % 7]
    direction = [[1, 0, 0], [1, 1, 0], [1, 1, 1], [0, 1, 2]][seed % 4]
    return pos, vel, types, box, cells, float(rng.choice([0.001, 0.002, 0.005])), kw, direction, bool(seed % 3 == 0)


@pytest.mark.parametrize("seed", range(14))
def test_oracle_equals_the_real_reference_on_random_inputs(seed):
    """Seeded random trajectories, cells (incl. tilted), group rules, modes and frame counts (odd and not a power of
    two): the oracle's k-path and SED equal the imported reference bit for bit.  Skipped where /root/reference is
    absent (the GPU box)."""
    from oracle.ref_import import load_reference
    psa = load_reference()
    if psa is None:
        pytest.skip("reference tree not on this machine")
    pos, vel, types, box, cells, dt, kw, direction, disp = _random_case(seed)
    n_t = pos.shape[0]
    traj = psa.Trajectory(pos, vel, types, np.arange(n_t), box, np.diag(box).copy(),
                          np.array([box[0, 1], box[0, 2], box[1, 2]], np.float32), dt)
    calc = psa.SEDCalculator(traj, *cells, use_displacements=disp)
    mags, kv = calc.get_k_path(direction, 2.5, 9)
    ref = calc.calculate(mags, kv, **kw)
    new = O.calculate(pos, vel, types, dt, kv, use_displacements=disp, **kw)
    assert new["is_complex"] == ref.is_complex and new["sed"].dtype == ref.sed.dtype
    np.testing.assert_array_equal(new["sed"], ref.sed)
    np.testing.assert_array_equal(new["freqs"], ref.freqs)
    # host-side k-path of the drop-in, same inputs
    from psa_b200 import kspace
    lat = kspace.Lattice.from_box(box, *cells)
    m2, kv2 = kspace.k_path(lat, direction, 2.5, 9, None)
    np.testing.assert_array_equal(m2, mags)
    np.testing.assert_array_equal(kv2, kv)


@pytest.mark.parametrize("seed", range(8))
def test_oracle_ised_and_chiral_equal_the_real_reference_on_random_inputs(seed):
    """iSED frames (flat type list = one group per type, index lists, 'auto' and numeric rescale) and the three
    chiral-phase options on random inputs, oracle vs the imported reference, bit for bit."""
    from oracle.ref_import import load_reference
    psa = load_reference()
    if psa is None:
        pytest.skip("reference tree not on this machine")
    from oracle.make_golden import _ised_frames
    from psa_b200 import kspace
    pos, vel, types, box, cells, dt, _, _, _ = _random_case(100 + seed)
    n_t, n_a = pos.shape[:2]
    traj = psa.Trajectory(pos, vel, types, np.arange(n_t), box, np.diag(box).copy(),
                          np.array([box[0, 1], box[0, 2], box[1, 2]], np.float32), dt)
    calc = psa.SEDCalculator(traj, *cells)
    rng = np.random.default_rng(seed)
    # chiral phase
    z1 = (rng.standard_normal((6, 5)) + 1j * rng.standard_normal((6, 5))).astype(np.complex64)
    z2 = (rng.standard_normal((6, 5)) + 1j * rng.standard_normal((6, 5))).astype(np.complex64)
    z2[0, 0] = 0
    for opt in "ABC":
        np.testing.assert_array_equal(O.chiral_phase(z1, z2, opt), calc.calculate_chiral_phase(z1, z2, opt), err_msg=opt)
    # iSED
    direction = [[1, 0, 0], [1, 1, 0]][seed % 2]
    char_len, nk = 5.0 + seed, 7
    spec = [dict(basis_atom_types_ised=[1, 2]), dict(basis_atom_idx_ised=[[0, 1], [2]]), dict()][seed % 3]
    rescale = ["auto", 0.5, 2][seed % 3]
    k_target, w_target = 0.4 + 0.1 * seed, 20.0 + 10 * seed
    want = _ised_frames(psa, calc, k_dir_spec=direction, k_target=k_target, w_target=w_target,
                        char_len_k_path=char_len, nk_on_path=nk, bz_cov_ised=1.0, rescale_factor=rescale,
                        n_recon_frames=5, **spec)
    from psa_b200 import groups as G
    from psa_b200 import parse_direction
    k_hat = parse_direction(direction)
    lat = kspace.Lattice.from_box(box, *cells)
    mags, vecs = kspace.k_path(lat, k_hat, 1.0, nk, char_len)
    groups = G.resolve_ised_groups(types, n_a, spec.get("basis_atom_idx_ised"), spec.get("basis_atom_types_ised"))
    got = O.ised(pos, vel, types, dt, k_hat, mags, vecs, k_target, w_target, groups, rescale_factor=rescale, n_frames=5)
    np.testing.assert_array_equal(got["frames"], want)


def test_numpy_complex64_fft_is_the_rounded_float64_transform():
    """Why the CUDA FFT is carried in float64: the reference's `np.fft.fft(..., axis=0)` on complex64 data
    (sed_calculator.py:83) returns, in the NumPy it runs on, exactly the float64 transform rounded once to complex64 -
    every bin accurate to its own magnitude, weak bins next to a strong line included.  A float32-butterfly transform
    cannot match that, whatever its twiddles."""
    rng = np.random.default_rng(4)
    for n in (250, 1000, 4096, 16384):
        z = (rng.standard_normal((n, 5)) + 1j * rng.standard_normal((n, 5))).astype(np.complex64)
        z[:, 0] += (30 * np.exp(2j * np.pi * 37 * np.arange(n) / n)).astype(np.complex64)       # a line 30x over the noise
        got = np.fft.fft(z, axis=0)
        assert got.dtype == np.complex64
        want = np.fft.fft(z.astype(np.complex128), axis=0).astype(np.complex64)
        np.testing.assert_array_equal(got, want)
