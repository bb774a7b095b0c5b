"""Generate ``tests/golden/*.npz`` by running the REAL reference on seeded inputs.

TEST INFRASTRUCTURE ONLY.  Run in the build container (where ``/root/reference``
is mounted):

    python oracle/make_golden.py

The reference's own tests contain no vector for the SED path, so
these files are what pins ``oracle/psa_oracle.py`` - and, through it, the CUDA
path.  Each file stores the inputs (so the tests do not depend on the synthetic
generator staying frozen) and the reference's outputs for every API quirk listed
in SURVEY.md 8a (rows A8-A13).  Environment recorded in the files: NumPy /
OpenBLAS versions, because the float32 results depend on them.
"""
from __future__ import annotations

import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle.ref_import import load_reference  # noqa: E402
import synthetic as synth  # noqa: E402

OUT = ROOT / "tests" / "golden"


def _env() -> dict:
    info = {"numpy": np.__version__}
    try:
        from threadpoolctl import threadpool_info
        for lib in threadpool_info():
            if lib.get("internal_api") == "openblas":
                info["openblas"] = f"{lib.get('version')} {lib.get('architecture')}"
    except Exception:
        pass
    return info


def _ref_traj(psa, traj):
    return psa.Trajectory(positions=traj.positions, velocities=traj.velocities, types=traj.types,
                          timesteps=traj.timesteps, box_matrix=traj.box_matrix,
                          box_lengths=traj.box_lengths, box_tilts=traj.box_tilts, dt_ps=traj.dt_ps)


def _ised_frames(psa, calc, **kw) -> np.ndarray:
    """Run reference ``ised`` and capture the frames it hands to its dump writer."""
    mod = sys.modules["psa.core.sed_calculator"]
    captured = {}
    orig = mod.out_to_qdump

    def sink(filename, positions_tf, types_tf, box_matrix):
        captured["frames"] = np.array(positions_tf)
        captured["types"] = np.array(types_tf)

    mod.out_to_qdump = sink
    try:
        with tempfile.TemporaryDirectory() as tmp:
            calc.ised(dump_filepath=str(Path(tmp) / "x.dump"), plot_dir_ised=None, **kw)
    finally:
        mod.out_to_qdump = orig
    return captured["frames"]


def si_case(psa) -> dict:
    spec = synth.si_spec("gold_si", n_cells=2, n_frames=256, seed=11, n_modes=5)
    traj = spec.trajectory(threads=1)
    calc = psa.SEDCalculator(_ref_traj(psa, traj), *spec.cells)
    g = dict(positions=traj.positions, velocities=traj.velocities, types=traj.types,
             box_matrix=traj.box_matrix, dt_ps=np.float64(traj.dt_ps), cells=np.array(spec.cells))
    g["recip_vecs_prim"] = calc.recip_vecs_prim
    g["a1"], g["b1"] = np.asarray(calc.a1), np.asarray(calc.b1)

    for tag, direction in (("100", [1, 0, 0]), ("110", [1, 1, 0]), ("111", "111")):
        mags, vecs = calc.get_k_path(direction, 4.0, 12)
        g[f"kpath_{tag}_mags"], g[f"kpath_{tag}_vecs"] = mags, vecs
    mags, vecs = calc.get_k_path("x", 1.0, 9, lat_param=synth.SI_A)
    g["kpath_lat_mags"], g["kpath_lat_vecs"] = mags, vecs
    mags1, vecs1 = calc.get_k_path([0, 1, 0], 2.0, 1)
    g["kpath_nk1_mags"], g["kpath_nk1_vecs"] = mags1, vecs1
    for plane in ("xy", "yz", "zx"):
        _, gv, shape = calc.get_k_grid(plane, (-1.5, 2.0), (-0.5, 1.0), 4, 3, 0.25)
        g[f"kgrid_{plane}_vecs"], g[f"kgrid_{plane}_shape"] = gv, np.array(shape)

    k100, v100 = g["kpath_100_mags"], g["kpath_100_vecs"]
    k110, v110 = g["kpath_110_mags"], g["kpath_110_vecs"]
    run = lambda kv, **kw: calc.calculate(np.zeros(len(kv), np.float32), kv, **kw)  # noqa: E731
    cases = {
        "coh_all_100": run(v100),
        "coh_all_110": run(v110),
        "coh_all_111": run(g["kpath_111_vecs"]),
        "coh_all_100_chunk5": run(v100, k_chunk_size=5),
        "coh_types12": run(v110, basis_atom_types=[1, 2], summation_mode="coherent"),
        "inc_types12": run(v110, basis_atom_types=[1, 2], summation_mode="incoherent"),
        "inc_types1": run(v110, basis_atom_types=[1], summation_mode="incoherent"),
        "inc_types_nested": run(v110, basis_atom_types=[[1, 2]], summation_mode="incoherent"),
        "inc_types_unknown": run(v100, basis_atom_types=[7], summation_mode="incoherent"),
        "inc_types_1_and_unknown": run(v100, basis_atom_types=[1, 7], summation_mode="incoherent"),
        "inc_indices": run(v100, basis_atom_indices=[[0, 1, 5, 9], [2, 3, 40]], summation_mode="incoherent"),
        "coh_indices_union": run(v100, basis_atom_indices=[[0, 1, 5, 9], [2, 3, 5]], summation_mode="coherent"),
        "coh_indices_flat_dup": run(v100, basis_atom_indices=[3, 1, 1, 20], summation_mode="coherent"),
        "coh_indices_ndarray": run(v100, basis_atom_indices=np.array([4, 8, 15, 16, 23, 42])),
        "inc_all": run(v100, summation_mode="incoherent"),
        "kgrid_xy": calc.calculate(np.array([], np.float32), g["kgrid_xy_vecs"], k_grid_shape=(4, 3)),
    }
    for name, res in cases.items():
        g[f"sed_{name}"] = res.sed
        g[f"cplx_{name}"] = np.array(res.is_complex)
    g["freqs"] = cases["coh_all_100"].freqs
    g["intensity_coh_all_100"] = cases["coh_all_100"].intensity

    calc_d = psa.SEDCalculator(_ref_traj(psa, traj), *spec.cells, use_displacements=True)
    g["sed_disp_coh_all_100"] = calc_d.calculate(k100, v100).sed
    g["sed_disp_inc_types12"] = calc_d.calculate(k110, v110, basis_atom_types=[1, 2],
                                                 summation_mode="incoherent").sed

    # iSED: float rescale with per-type groups, 'auto' rescale with all atoms, index groups
    common = dict(k_dir_spec=[1, 0, 0], char_len_k_path=synth.SI_A, nk_on_path=9, bz_cov_ised=1.0,
                  n_recon_frames=8)
    inten = cases["coh_all_100"].intensity
    f_pk = int(np.argmax(inten[1:inten.shape[0] // 2, :].max(axis=1))) + 1
    w_t = float(g["freqs"][f_pk])
    g["ised_w_target"], g["ised_k_target"] = np.float64(w_t), np.float64(0.58)
    g["ised_types_float"] = _ised_frames(psa, calc, k_target=0.58, w_target=w_t,
                                         basis_atom_types_ised=[1, 2], rescale_factor=0.5, **common)
    g["ised_all_auto"] = _ised_frames(psa, calc, k_target=0.58, w_target=w_t,
                                      rescale_factor="auto", **common)
    g["ised_idx_auto"] = _ised_frames(psa, calc, k_target=0.3, w_target=w_t * 0.5,
                                      basis_atom_idx_ised=[[0, 1, 2, 3], [10, 11, 12]],
                                      rescale_factor="auto", **common)

    # odd frame count (non power of two FFT length)
    traj_odd = spec.wrap(traj.positions[:250], traj.velocities[:250])
    calc_odd = psa.SEDCalculator(_ref_traj(psa, traj_odd), *spec.cells)
    g["sed_odd250_coh_all_100"] = calc_odd.calculate(k100, v100).sed
    return g


def graphene_case(psa) -> dict:
    spec = synth.graphene_spec("gold_gr", n_cells=6, n_frames=128, seed=12, n_modes=4)
    traj = spec.trajectory(threads=1)
    calc = psa.SEDCalculator(_ref_traj(psa, traj), *spec.cells)
    g = dict(positions=traj.positions, velocities=traj.velocities, types=traj.types,
             box_matrix=traj.box_matrix, dt_ps=np.float64(traj.dt_ps), cells=np.array(spec.cells))
    g["recip_vecs_prim"] = calc.recip_vecs_prim
    mags, vecs = calc.get_k_path([1, 0, 0], 4.0, 10)
    g["kpath_mags"], g["kpath_vecs"] = mags, vecs
    res = calc.calculate(mags, vecs, summation_mode="coherent")
    g["sed_coh"] = res.sed
    g["freqs"] = res.freqs
    for axis, (i, j) in (("x", (1, 2)), ("y", (0, 2)), ("z", (0, 1))):
        g[f"phase_C_{axis}"] = calc.calculate_chiral_phase(res.sed[:, :, i], res.sed[:, :, j], "C")
    g["phase_A_z"] = calc.calculate_chiral_phase(res.sed[:, :, 0], res.sed[:, :, 1], "A")
    g["phase_B_z"] = calc.calculate_chiral_phase(res.sed[:, :, 0], res.sed[:, :, 1], "B")
    return g


def dump_cases(psa) -> None:
    """Text fixtures of the reference's iSED dump writer (src/psa/io/writer.py:139-228): an orthogonal and a
    triclinic box, written by ``out_to_qdump`` itself from seeded frames (the inputs are regenerated by the test)."""
    from psa.io.writer import out_to_qdump  # type: ignore
    for name, box in DUMP_BOXES.items():
        frames, types = dump_inputs()
        out_to_qdump(str(OUT / f"dump_{name}.txt"), frames, types, np.array(box, np.float32))
        print(f"wrote dump_{name}.txt")


DUMP_BOXES = {"ortho": [[10.862, 0, 0], [0, 10.862, 0], [0, 0, 21.724]],
              "triclinic": [[12.3, 1.5, -0.75], [0, 10.6524, 2.25], [0, 0, 20.0]]}


def dump_inputs():
    rng = np.random.default_rng(2024)
    frames = (rng.random((3, 7, 3)) * 12 - 1).astype(np.float32)
    frames[1, 2] = [0.0, -0.0, 1e-7]
    frames[2, 6] = [123456.789, -9.87654321e-5, 3.0]
    return frames, np.array([1, 2, 1, 2, 3, 1, 2])


def main() -> None:
    psa = load_reference()
    if psa is None:
        raise SystemExit("reference tree not found (set PSA_REFERENCE_SRC)")
    OUT.mkdir(parents=True, exist_ok=True)
    dump_cases(psa)
    env = _env()
    for name, builder in (("si_small", si_case), ("graphene_small", graphene_case)):
        data = builder(psa)
        data["_env"] = np.array(repr(env))
        np.savez_compressed(OUT / f"{name}.npz", **data)
        size = (OUT / f"{name}.npz").stat().st_size
        print(f"wrote {name}.npz: {len(data)} arrays, {size / 1e6:.2f} MB, env {env}")


if __name__ == "__main__":
    main()
