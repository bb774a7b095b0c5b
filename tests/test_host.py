"""Host-side logic of the drop-in: directions, k-space, group rules, the SED container.
Mirrors the reference's own unit tests (tests/test_helpers.py, test_sed.py, test_trajectory.py)
and pins k-paths / k-grids / lattice vectors against the real reference's outputs (tests/golden)."""
from pathlib import Path

import numpy as np
import pytest

from psa_b200 import SED, Trajectory, parse_direction
from psa_b200 import groups as G
from psa_b200 import kspace

S2, S3 = 1 / np.sqrt(2), 1 / np.sqrt(3)


@pytest.mark.parametrize("spec,expected", [
    ("x", [1, 0, 0]), ("y", [0, 1, 0]), ("z", [0, 0, 1]), ("100", [1, 0, 0]),
    ("xy", [S2, S2, 0]), ("110", [S2, S2, 0]), ("xyz", [S3, S3, S3]), ("111", [S3, S3, S3]),
    ("0,1,0", [0, 1, 0]), (" 1 0 0 ", [1, 0, 0]),
    (0, [1, 0, 0]), (90, [0, 1, 0]), (45, [S2, S2, 0]), ("180.0", [-1, 0, 0]),
    ([1, 0, 0], [1, 0, 0]), ((0, 5, 0), [0, 1, 0]), (np.array([1, 1, 1]), [S3, S3, S3]),
    ([45], [S2, S2, 0]), (np.array(60.0), [0.5, np.sqrt(3) / 2, 0]),
    ({"angle": 30}, [np.sqrt(3) / 2, 0.5, 0]), ({"h": 1, "k": 0, "l": 0}, [1, 0, 0]),
    ({"h": 1, "k": 1, "l": 0}, [S2, S2, 0]), ({"h": 0, "k": 0, "l": 2}, [0, 0, 1]),
])
def test_parse_direction(spec, expected):
    out = parse_direction(spec)
    assert out.dtype == np.float32
    np.testing.assert_allclose(out, np.array(expected, np.float32), atol=1e-6)


@pytest.mark.parametrize("bad", ["invalid_string", [1, 2], [1, 2, 3, 4], np.array([[1, 0, 0], [0, 1, 0]]),
                                 {"a": 1, "b": 2}, [0, 0, 0], np.array([1e-8, 1e-9, 1e-10], np.float32)])
def test_parse_direction_invalid(bad):
    with pytest.raises(ValueError):
        parse_direction(bad)


def test_parse_direction_type_error():
    with pytest.raises(TypeError, match="Unsupported direction type: <class 'NoneType'>"):
        parse_direction(None)


def test_lattice_and_kpaths_match_reference(gold_si):
    g = gold_si
    lat = kspace.Lattice.from_box(g["box_matrix"], *g["cells"])
    np.testing.assert_array_equal(lat.recip_vecs_prim, g["recip_vecs_prim"])
    np.testing.assert_array_equal(lat.a1, g["a1"])
    np.testing.assert_array_equal(lat.b1, g["b1"])
    for tag, d in (("100", [1, 0, 0]), ("110", [1, 1, 0]), ("111", "111")):
        mags, vecs = kspace.k_path(lat, d, 4.0, 12)
        np.testing.assert_array_equal(mags, g[f"kpath_{tag}_mags"])
        np.testing.assert_array_equal(vecs, g[f"kpath_{tag}_vecs"])
        assert mags.dtype == vecs.dtype == np.float32
    mags, vecs = kspace.k_path(lat, "x", 1.0, 9, lat_param=5.431)
    np.testing.assert_array_equal(mags, g["kpath_lat_mags"])
    np.testing.assert_array_equal(vecs, g["kpath_lat_vecs"])
    mags, vecs = kspace.k_path(lat, [0, 1, 0], 2.0, 1)
    np.testing.assert_array_equal(mags, g["kpath_nk1_mags"])
    np.testing.assert_array_equal(vecs, g["kpath_nk1_vecs"])
    with pytest.raises(ValueError):
        kspace.k_path(lat, "x", 1.0, 0)


def test_kgrid_matches_reference(gold_si):
    for plane in ("xy", "yz", "zx"):
        empty, vecs, shape = kspace.k_grid(plane, (-1.5, 2.0), (-0.5, 1.0), 4, 3, 0.25)
        assert empty.size == 0 and empty.dtype == np.float32
        np.testing.assert_array_equal(vecs, gold_si[f"kgrid_{plane}_vecs"])
        assert shape == tuple(gold_si[f"kgrid_{plane}_shape"])
    with pytest.raises(ValueError):
        kspace.k_grid("xx", (0, 1), (0, 1), 2, 2)
    with pytest.raises(ValueError):
        kspace.k_grid("xy", (0, 1), (0, 1), 0, 2)


def test_graphene_lattice(gold_gr):
    lat = kspace.Lattice.from_box(gold_gr["box_matrix"], *gold_gr["cells"])
    np.testing.assert_array_equal(lat.recip_vecs_prim, gold_gr["recip_vecs_prim"])
    mags, vecs = kspace.k_path(lat, [1, 0, 0], 4.0, 10)
    np.testing.assert_array_equal(vecs, gold_gr["kpath_vecs"])


def test_lattice_errors():
    box = np.eye(3, dtype=np.float32) * 10
    with pytest.raises(ValueError):
        kspace.Lattice.from_box(box, 0, 1, 1)
    flat = box.copy(); flat[2] = 0
    with pytest.raises(ValueError):
        kspace.Lattice.from_box(flat, 1, 1, 1)


# ---- group rules
TYPES = np.array([1, 2, 1, 2, 1, 3], np.int32)


def test_groups_types_flat():
    coh = G.resolve_sed_groups(TYPES, 6, basis_atom_types=[1, 2], summation_mode="coherent")
    assert len(coh) == 1 and list(coh[0]) == [0, 1, 2, 3, 4]
    inc = G.resolve_sed_groups(TYPES, 6, basis_atom_types=[1, 2], summation_mode="incoherent")
    assert [list(x) for x in inc] == [[0, 2, 4], [1, 3]]
    cplx, proj = G.plan_sed_groups(inc, "incoherent")
    assert not cplx and len(proj) == 2


def test_groups_single_group_incoherent_is_complex():
    one = G.resolve_sed_groups(TYPES, 6, basis_atom_types=[1], summation_mode="incoherent")
    cplx, proj = G.plan_sed_groups(one, "incoherent")
    assert cplx and list(proj[0]) == [0, 2, 4]
    nested = G.resolve_sed_groups(TYPES, 6, basis_atom_types=[[1, 2]], summation_mode="incoherent")
    assert G.plan_sed_groups(nested, "incoherent")[0]


def test_groups_unknown_type_falls_back_to_all():
    out = G.resolve_sed_groups(TYPES, 6, basis_atom_types=[7], summation_mode="incoherent")
    assert len(out) == 1 and list(out[0]) == list(range(6))


def test_groups_indices_and_errors():
    out = G.resolve_sed_groups(TYPES, 6, basis_atom_indices=[[0, 1], [2, 3]], summation_mode="incoherent")
    assert [list(x) for x in out] == [[0, 1], [2, 3]]
    cplx, proj = G.plan_sed_groups(G.resolve_sed_groups(TYPES, 6, basis_atom_indices=[[0, 1], [1, 3]]), "coherent")
    assert cplx and list(proj[0]) == [0, 1, 3]
    dup = G.resolve_sed_groups(TYPES, 6, basis_atom_indices=[3, 1, 1])
    assert list(dup[0]) == [3, 1, 1]                      # a single flat list keeps order and duplicates
    with pytest.raises(ValueError):
        G.resolve_sed_groups(TYPES, 6, basis_atom_indices=[0, 6])
    with pytest.raises(ValueError):
        G.resolve_sed_groups(TYPES, 6, basis_atom_types=[1, [2]])
    both = G.resolve_sed_groups(TYPES, 6, basis_atom_indices=[0], basis_atom_types=[2])
    assert list(both[0]) == [1, 3]                        # types win


def test_ised_groups_flat_types_are_per_type():
    out = G.resolve_ised_groups(TYPES, 6, basis_atom_types_ised=[1, 2])
    assert [list(x) for x in out] == [[0, 2, 4], [1, 3]]
    assert len(G.resolve_ised_groups(TYPES, 6)) == 1
    assert G.resolve_ised_groups(TYPES, 6, basis_atom_types_ised=[9]) == []
    with pytest.raises(ValueError):
        G.resolve_ised_groups(TYPES, 6, basis_atom_idx_ised=[0, 99])


# ---- SED container (reference tests/test_sed.py)
@pytest.fixture
def sed_data():
    rng = np.random.default_rng(3)
    s = (rng.random((10, 5, 3)) + 1j * rng.random((10, 5, 3))).astype(np.complex64)
    return dict(sed=s, freqs=np.linspace(0, 10, 10, dtype=np.float32),
                k_points=np.linspace(0, 1, 5, dtype=np.float32),
                k_vectors=rng.random((5, 3)).astype(np.float32),
                phase=rng.random((10, 5)).astype(np.float32))


def test_sed_intensity_and_dict_access(sed_data):
    obj = SED(**sed_data)
    expected = np.sum(np.abs(sed_data["sed"]) ** 2, axis=-1).astype(np.float32)
    np.testing.assert_array_equal(obj.intensity, expected)
    assert obj["sed"] is obj.sed and "phase" in obj and set(obj.keys()) >= {"sed", "freqs", "is_complex"}
    np.testing.assert_array_equal(obj["intensity"], expected)
    with pytest.raises(KeyError):
        obj["nope"]
    empty = SED(sed=np.array([]).reshape(0, 0, 3), freqs=np.array([]), k_points=np.array([]),
                k_vectors=np.array([]).reshape(0, 3))
    assert empty.intensity.shape == (0, 0)


def test_sed_save_load(sed_data, tmp_path):
    obj = SED(**sed_data, k_grid_shape=(5, 1))
    obj.save(tmp_path / "run1")
    for suffix in (".sed.npy", ".freqs.npy", ".k_points.npy", ".k_vectors.npy", ".phase.npy", ".k_grid_shape.npy"):
        assert (tmp_path / "run1").with_suffix(suffix).exists()
    back = SED.load(tmp_path / "run1")
    np.testing.assert_array_equal(back.sed, obj.sed)
    np.testing.assert_array_equal(back.phase, obj.phase)
    assert back.k_grid_shape == (5, 1)
    with pytest.raises(FileNotFoundError):
        SED.load(tmp_path / "missing")


def test_trajectory_validation():
    ok = dict(positions=np.zeros((4, 3, 3), np.float32), velocities=np.zeros((4, 3, 3), np.float32),
              types=np.ones(3, np.int32), timesteps=np.arange(4), box_matrix=np.eye(3, dtype=np.float32),
              box_lengths=np.ones(3, np.float32), box_tilts=np.zeros(3, np.float32), dt_ps=0.001)
    t = Trajectory(**ok)
    assert t.n_frames == 4 and t.n_atoms == 3
    for key, bad in (("positions", np.zeros((4, 3, 2))), ("velocities", np.zeros((5, 3, 3))),
                     ("types", np.ones((3, 1))), ("timesteps", np.arange(5)),
                     ("box_matrix", np.eye(2)), ("box_lengths", np.ones(2)), ("box_tilts", np.ones(4))):
        with pytest.raises(ValueError):
            Trajectory(**{**ok, key: bad})


def test_npy_cache_roundtrip_matches_reference_loader(tmp_path, gold_si):
    """N1: the .npy cache bundle is read (memory-mapped) with the reference loader's field semantics."""
    from psa_b200 import Trajectory
    from psa_b200 import cache
    g = gold_si
    box = g["box_matrix"]
    traj = Trajectory(g["positions"][:16], g["velocities"][:16], g["types"], np.arange(16), box,
                      np.diag(box).copy(), np.zeros(3, np.float32), 0.002)
    fake = tmp_path / "run.lammpstrj"
    assert not cache.has_npy_cache(fake)
    with pytest.raises(FileNotFoundError):
        cache.load_npy_cache(fake, 0.002)
    cache.save_npy_cache(traj, fake)
    assert cache.has_npy_cache(fake)
    back = cache.load_npy_cache(fake, dt=0.002)
    assert isinstance(back.positions, np.memmap) and back.positions.flags.writeable
    np.testing.assert_array_equal(back.positions, traj.positions)
    np.testing.assert_array_equal(back.velocities, traj.velocities)
    np.testing.assert_array_equal(back.timesteps, np.arange(16, dtype=np.float32) * 0.002)
    assert back.dt_ps == 0.002 and back.n_atoms == traj.n_atoms
    # same answer as the reference's own loader, when the reference is on this machine
    from oracle.ref_import import load_reference
    psa = load_reference()
    if psa is not None:
        fake.write_text("placeholder: the loader only checks that the file exists before using the cache\n")
        ref = psa.TrajectoryLoader(str(fake), dt=0.002).load()
        for name in ("positions", "velocities", "types", "timesteps", "box_matrix", "box_lengths", "box_tilts"):
            np.testing.assert_array_equal(getattr(back, name), getattr(ref, name), err_msg=name)
        assert ref.dt_ps == back.dt_ps
    with pytest.raises(ValueError):
        cache.load_npy_cache(fake, dt=0.0)


def test_mode_validation_needs_no_gpu():
    """Argument errors surface before any device work (reference: sed_calculator.py:190-191)."""
    from psa_b200 import SEDCalculator
    import synthetic as synth
    spec = synth.si_spec("tiny", n_cells=1, n_frames=8, seed=1)
    calc = SEDCalculator(spec.trajectory(threads=1), *spec.cells)
    mags, vecs = calc.get_k_path([1, 0, 0], 1.0, 3)
    for call in (calc.calculate, calc.calculate_intensity):
        with pytest.raises(ValueError, match="summation_mode"):
            call(mags, vecs, summation_mode="partially coherent")
    assert calc._engine is None                      # nothing touched the GPU


@pytest.mark.parametrize("name", ["ortho", "triclinic"])
def test_dump_writer_is_byte_identical_to_the_reference(name, tmp_path):
    """N2: ``write_lammps_dump`` against text written by the reference's own ``out_to_qdump``
    (reference: src/psa/io/writer.py:139-228; fixtures from oracle/make_golden.py:dump_cases)."""
    from oracle.make_golden import DUMP_BOXES, dump_inputs
    from psa_b200.dump import write_lammps_dump
    frames, types = dump_inputs()
    out = tmp_path / "d.dump"
    write_lammps_dump(str(out), frames, types, np.array(DUMP_BOXES[name], np.float32))
    want = (Path(__file__).parent / "golden" / f"dump_{name}.txt").read_bytes()
    assert out.read_bytes() == want


@pytest.mark.parametrize("seed", range(6))
def test_lattice_and_k_grid_equal_the_real_reference_on_random_inputs(seed):
    """Primitive / reciprocal vectors for random (tilted) cells and k-grids on every plane, against the imported
    reference, bit for bit.  Skipped where /root/reference is absent."""
    from oracle.ref_import import load_reference
    psa = load_reference()
    if psa is None:
        pytest.skip("reference tree not on this machine")
    rng = np.random.default_rng(seed)
    box = np.diag(rng.uniform(6.0, 30.0, 3)).astype(np.float32)
    box[0, 1], box[0, 2], box[1, 2] = rng.uniform(-3, 3, 3).astype(np.float32)
    cells = tuple(int(c) for c in rng.integers(1, 6, 3))
    n_t, n_a = 4, 3
    z = np.zeros((n_t, n_a, 3), np.float32)
    traj = psa.Trajectory(z, z, np.ones(n_a, int), np.arange(n_t), box, np.diag(box).copy(),
                          np.array([box[0, 1], box[0, 2], box[1, 2]], np.float32), 0.002)
    ref = psa.SEDCalculator(traj, *cells)
    lat = kspace.Lattice.from_box(box, *cells)
    for name in ("a1", "a2", "a3", "b1", "b2", "b3", "recip_vecs_prim"):
        a, b = getattr(lat, name), getattr(ref, name)
        assert a.dtype == b.dtype, name
        np.testing.assert_array_equal(a, b, err_msg=name)
    plane = ["xy", "yz", "zx"][seed % 3]
    rx, ry = tuple(rng.uniform(-4, 4, 2)), tuple(rng.uniform(-4, 4, 2))
    nkx, nky, kfix = int(rng.integers(1, 6)), int(rng.integers(1, 6)), float(rng.uniform(-1, 1))
    want = ref.get_k_grid(plane, rx, ry, nkx, nky, kfix)
    got = kspace.k_grid(plane, rx, ry, nkx, nky, kfix)
    assert got[2] == want[2] and got[0].dtype == want[0].dtype and got[1].dtype == want[1].dtype
    np.testing.assert_array_equal(got[0], want[0])
    np.testing.assert_array_equal(got[1], want[1])


def test_four_step_fft_thread_code_on_the_host(tmp_path):
    """psa_b200/csrc/fft4.cuh is plain C++ when compiled without nvcc: run both stages of the four-step FFT thread by
    thread on the CPU (index maps, twiddles, butterflies) against a direct float64 transform, for the three lengths
    the kernel serves.  Every output must be written exactly once."""
    import shutil
    import subprocess
    from pathlib import Path
    gxx = shutil.which("g++")
    if gxx is None:
        import pytest
        pytest.skip("g++ not available")
    src = Path(__file__).resolve().parent / "fft4_host_check.cpp"
    exe = tmp_path / "fft4_host_check"
    subprocess.run([gxx, "-O2", "-std=c++17", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr


def test_dump_writer_formats_every_float32_like_python(tmp_path):
    """The C++ writer's integer `%.6f` against CPython's own formatting of the same float32 values: exact ties of the
    binary value (x.xxxxxx5 with a finite binary expansion), negatives that round to -0.000000, denormals, large
    magnitudes (C library path), NaN and infinities, random values over 60 orders of magnitude."""
    import numpy as np
    from psa_b200.dump import write_lammps_dump
    rng = np.random.default_rng(17)
    special = np.array([0.0, -0.0, 0.5e-6, -0.5e-6, 1.5e-6, 2.5e-6, 0.0000005, 0.0000015, 0.125, 0.0078125, 1.0000005,
                        -1e-7, -4.9e-7, 1e-45, -1e-45, 1e-38, 123456.7890625, 999999.9999995, 1048576.5, 8388607.5,
                        16777216.0, 1.0995116e12, 3.4e38, -3.4e38, np.nan, np.inf, -np.inf, 0.9999995, 0.99999951,
                        2.0 ** -20, 2.0 ** -21, 3 * 2.0 ** -21, 5 * 2.0 ** -22, 7.0000005, 43.4300005], np.float32)
    rand = (rng.standard_normal(4000) * 10.0 ** rng.integers(-30, 30, 4000)).astype(np.float32)
    vals = np.concatenate([special, rand, rng.uniform(-50, 50, 3000).astype(np.float32)])
    vals = np.concatenate([vals, np.zeros((-len(vals)) % 3, np.float32)])
    frames = vals.reshape(1, -1, 3)
    n_at = frames.shape[1]
    types = (np.arange(n_at) % 3 + 1).astype(np.int32)
    box = np.array([[10.5, 0.0, 0.0], [0.0, 11.25, 0.0], [0.0, 0.0, 12.0]], np.float32)
    out = tmp_path / "f.dump"
    write_lammps_dump(str(out), frames, types, box, threads=3)
    lines = out.read_text().splitlines()
    assert lines[:9] == ["ITEM: TIMESTEP", "0", "ITEM: NUMBER OF ATOMS", str(n_at), "ITEM: BOX BOUNDS pp pp pp",
                         "0.00000000 10.50000000", "0.00000000 11.25000000", "0.00000000 12.00000000",
                         "ITEM: ATOMS id type x y z"]
    for a, line in enumerate(lines[9:]):
        x, y, z = (float(v) for v in frames[0, a])
        assert line == f"{a + 1} {int(types[a])} {x:.6f} {y:.6f} {z:.6f}", (a, line)
    # many frames, more frames than threads and fewer: same bytes whatever the thread count
    many = rng.standard_normal((7, 33, 3)).astype(np.float32)
    write_lammps_dump(str(tmp_path / "a.dump"), many, np.ones(33, int), box, threads=1)
    write_lammps_dump(str(tmp_path / "b.dump"), many, np.ones(33, int), box, threads=16)
    assert (tmp_path / "a.dump").read_bytes() == (tmp_path / "b.dump").read_bytes()


def test_cli_host_logic(tmp_path):
    """N4: config merge, direction labels, basis resolution and the missing-cache error of the batch driver
    (reference: src/psa/cli.py:38-56, 79-89, 108-119)."""
    import numpy as np
    import pytest
    from psa_b200 import cli
    cfg = cli.update_dict_recursively({"a": {"b": 1, "c": 2}, "d": 3}, {"a": {"b": 5}, "e": {"f": 1}})
    assert cfg == {"a": {"b": 5, "c": 2}, "d": 3, "e": {"f": 1}}
    assert cli.direction_label([1, 0, 0], 1) == "1.00_0.00_0.00" and cli.direction_label(45, 2) == "45.0deg"
    assert cli.direction_label("x y/z", 3) == "x_y-z" and cli.direction_label({"h": 1, "k": 1}, 4) == "h1_k1_l0"
    types = np.array([1, 1, 2, 2, 3])
    idx, sfx = cli.resolve_basis(types, 5, {"atom_indices": None, "atom_types": [2, 3]})
    assert idx.tolist() == [2, 3, 4] and sfx == "_typebasis2_3"
    idx, sfx = cli.resolve_basis(types, 5, {"atom_indices": [0, 4], "atom_types": [2]})
    assert idx.tolist() == [0, 4] and sfx == "_idxbasis"
    assert cli.resolve_basis(types, 5, {"atom_indices": None, "atom_types": [9]}) == (None, "")
    with pytest.raises(ValueError):
        cli.resolve_basis(types, 5, {"atom_indices": [7], "atom_types": None})
    assert cli.main(["--trajectory", str(tmp_path / "none.lammpstrj"), "--output-dir", str(tmp_path / "o")]) == 1


def test_k_chunk_planning():
    """Equal chunks by default; with a frame-range hint (pipelined multi-GPU exchange) the first chunk is a whole number
    of 128-row tiles that fills whole waves of CTA pairs per range launch."""
    from psa_b200.engine import effective_k_chunk, plan_k_chunks
    assert effective_k_chunk(500, 10000) == 1000 and effective_k_chunk(500, 200) == 200 and effective_k_chunk(7, 100) == 7
    assert plan_k_chunks(10000, 500) == [(k, 1000) for k in range(0, 10000, 1000)]
    assert plan_k_chunks(0, 500) == [] and plan_k_chunks(7, 3) == [(0, 3), (3, 3), (6, 1)]
    first = plan_k_chunks(1250, 500, (2048, 74))
    assert first == [(0, 576), (576, 674)]                 # 9 row tiles x 8 frame tiles x 3 = 216 tiles = 2.92 waves of 74
    assert plan_k_chunks(200, 500, (8192, 74)) == [(0, 200)]
    for n_k, hint in ((2500, (4096, 74)), (5000, (8192, 74)), (96, (1365, 74))):
        chunks = plan_k_chunks(n_k, 500, hint)
        assert chunks[0][0] == 0 and sum(nk for _, nk in chunks) == n_k
        assert all(b[0] == a[0] + a[1] for a, b in zip(chunks, chunks[1:]))


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver times next to the GPU arm) runs without a GPU and prints ONE
    JSON line with the contract's keys; its `config` is the object the GPU arm prints for the same workload."""
    import json
    import subprocess
    import sys
    root = Path(__file__).resolve().parents[1]
    out = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--frames", "256", "--cells", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=str(root))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["n_gpus"] == 1
    assert line["metric"].endswith("atoms/sec") and line["unit"] == "k*t*atom/s" and line["value"] > 0
    assert set(line["config"]) == {"workload", "desc"} and line["config"]["workload"] == "c4"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "k-points" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0
