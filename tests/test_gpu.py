"""Parity tests proper: every CUDA kernel, called through the C ABI, against (1) the bit-exact
integer model, (2) the oracle, (3) the golden outputs of the real reference.  Run with `-m gpu`."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
import intmodel as M  # noqa: E402
from oracle import psa_oracle as O  # noqa: E402
import synthetic as synth  # noqa: E402

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def eng():
    from psa_b200 import _lib
    from psa_b200.engine import Engine
    assert _lib.load().psa_version() >= 100
    return Engine()


def dev(eng, arr):
    return torch.from_numpy(np.ascontiguousarray(arr)).to(eng.device)


# ------------------------------------------------------------------ ingest
# (n_atoms % 4 == 0 -> TMA kernel, otherwise the register-staged one; ragged last tiles in both directions)
@pytest.mark.parametrize("n_t,n_a", [(256, 64), (3001, 37), (17, 1000), (3001, 44), (1000, 4), (2, 8), (64 * 9, 20)])
def test_mean_positions_bit_exact(eng, n_t, n_a):
    rng = np.random.default_rng(n_t)
    pos = (rng.random((n_t, n_a, 3)) * 43.4 + rng.standard_normal((n_t, n_a, 3)) * 0.01).astype(np.float32)
    got = eng.mean_positions(dev(eng, pos)).cpu().numpy()
    np.testing.assert_array_equal(got, np.mean(pos, axis=0, dtype=np.float32))


def test_mean_positions_unaligned_base_uses_fallback(eng):
    """A trajectory whose first element is not 16-byte aligned cannot be described by a tensor map."""
    rng = np.random.default_rng(3)
    n_t, n_a = 300, 16
    flat = torch.from_numpy(rng.standard_normal(n_t * n_a * 3 + 1).astype(np.float32)).to(eng.device)
    pos = flat[1:].view(n_t, n_a, 3)
    assert pos.data_ptr() % 16 != 0
    got = eng.mean_positions(pos).cpu().numpy()
    np.testing.assert_array_equal(got, np.mean(pos.cpu().numpy(), axis=0, dtype=np.float32))


def test_ingest_by_frame_ranges_is_bit_identical(eng):
    """The multi-GPU sliced ingest: running sums continued over frame ranges and row-range digitising give
    the bits of the one-shot kernels."""
    rng = np.random.default_rng(11)
    n_t, n_a = 1000, 48
    pos = (rng.random((n_t, n_a, 3)) * 20 + rng.standard_normal((n_t, n_a, 3)) * 0.05).astype(np.float32)
    d = dev(eng, pos)
    want = eng.mean_positions(d)
    acc = None
    cuts = [0, 333, 334, 900, 1000]
    for i, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
        acc = eng.mean_accumulate(d[a:b].contiguous(), acc, n_t if i == len(cuts) - 2 else 0)
    assert torch.equal(acc, want)
    np.testing.assert_array_equal(acc.cpu().numpy(), np.mean(pos, axis=0, dtype=np.float32))
    idx = dev(eng, np.array([1, 5, 6, 7, 40, 41], np.int32))
    for mean, sel, n_sel in ((None, None, n_a), (want, idx, 6)):
        dig, expo, pitch = eng.digitize(d, mean, sel, n_sel)
        dig2, expo2 = torch.zeros_like(dig), torch.zeros_like(expo)
        for a, b in zip(cuts[:-1], cuts[1:]):
            eng.digitize_rows(d[a:b].contiguous(), mean, sel, n_sel, pitch, dig2, expo2, n_t, a)
        assert torch.equal(expo2, expo)
        assert torch.equal(dig2[..., :n_sel], dig[..., :n_sel])


# "staged": a gathered selection whose frame row is 16-byte aligned goes through the bulk-copy (shared memory) path
@pytest.mark.parametrize("case", ["all", "subset", "displacement", "zeros", "staged", "staged_displacement"])
def test_digitize_exact(eng, case):
    rng = np.random.default_rng(5)
    n_t, n_a = 33, (204 if case.startswith("staged") else 203)
    data = (rng.standard_normal((n_t, n_a, 3)) * np.array([3.0, 0.02, 700.0])).astype(np.float32)
    idx, mean = None, None
    if case == "subset":
        idx = np.array([5, 0, 77, 77, 202, 13, 14, 15, 100], np.int32)
    if case.startswith("staged"):
        idx = np.concatenate([np.arange(0, 204, 2), [203, 1, 1]]).astype(np.int32)
    if case in ("displacement", "staged_displacement"):
        mean = (rng.random((n_a, 3)) * 3).astype(np.float32)
    if case == "zeros":
        data[:, :, 1] = 0.0
        data[7] = 0.0
    sel = data if idx is None else data[:, idx, :]
    if mean is not None:
        sel = sel - (mean if idx is None else mean[idx])[None]
    n_sel = sel.shape[1]
    dig, expo, pitch = eng.digitize(dev(eng, data), None if mean is None else dev(eng, mean),
                                    None if idx is None else dev(eng, idx), n_sel)
    x_ref, e_ref = M.digitize(sel)
    np.testing.assert_array_equal(expo.cpu().numpy(), e_ref)
    d = dig.cpu().numpy()                                     # (3, 4, n_t, pitch)
    assert pitch % 64 == 0 and pitch >= n_sel
    assert not d[:, :, :, n_sel:].any()                       # padding is zero
    for pol in range(3):
        np.testing.assert_array_equal(d[pol, :, :, :n_sel], M.balanced_digits(x_ref[pol]))


@pytest.mark.parametrize("n_a,with_mean", [(17072, False), (17076, True), (40000, False)])
def test_digitize_long_rows_cluster_path(eng, n_a, with_mean):
    """Rows longer than 200 KB: a cluster of eight CTAs stages one frame in shared memory (single HBM read),
    maxima exchanged through distributed shared memory.  Same digits as the integer model."""
    rng = np.random.default_rng(n_a)
    n_t = 5
    data = (rng.standard_normal((n_t, n_a, 3)) * np.array([3.0, 0.02, 700.0])).astype(np.float32)
    data[2, n_a - 1, 2] = 5000.0                      # the row maximum sits in the last CTA's slice
    data[3, 0, 0] = -77.0                             # ... and in the first one's
    mean = (rng.random((n_a, 3)) * 3).astype(np.float32) if with_mean else None
    sel = data if mean is None else data - mean[None]
    dig, expo, pitch = eng.digitize(dev(eng, data), None if mean is None else dev(eng, mean), None, n_a)
    x_ref, e_ref = M.digitize(sel)
    np.testing.assert_array_equal(expo.cpu().numpy(), e_ref)
    d = dig.cpu().numpy()
    assert not d[:, :, :, n_a:].any()
    for pol in range(3):
        np.testing.assert_array_equal(d[pol, :, :, :n_a], M.balanced_digits(x_ref[pol]))


def test_phase_digits(eng):
    rng = np.random.default_rng(6)
    n_a, n_k = 500, 23
    mean = (rng.random((n_a, 3)) * 43.0).astype(np.float32)
    kv = (rng.standard_normal((n_k, 3)) * 2.5).astype(np.float32)
    kv[0] = 0.0                                                # Gamma: cos = 1 exactly, sin = 0
    idx = rng.permutation(n_a)[:301].astype(np.int32)
    for sel_idx in (None, idx):
        msel = mean if sel_idx is None else mean[sel_idx]
        n_sel = msel.shape[0]
        pitch = -(-n_sel // 64) * 64
        rows_alloc = 2 * n_k + 6
        ad = eng.phase_digits(dev(eng, kv), dev(eng, mean), None if sel_idx is None else dev(eng, sel_idx),
                              n_sel, pitch, rows_alloc).cpu().numpy()
        got = M.digits_to_int(ad[:, :2 * n_k, :n_sel])
        want = M.phase_ints(kv, msel)
        assert got[0].min() == got[0].max() == 2 ** 30 and not got[1].any()
        diff = np.abs(got - want)
        assert (diff == 0).mean() > 0.995, (diff == 0).mean()   # rest: fma-vs-sgemm / double-rounding corner cases
        assert diff.max() <= 2 ** 8, diff.max()                 # at most a few float32 ulps of a value near 1
        assert not ad[:, :2 * n_k, n_sel:].any()


# ------------------------------------------------------------------ projection
def _proj_inputs(rng, rows, n_t, n_sel):
    xa = rng.integers(-2 ** 30, 2 ** 30, (rows, n_sel), endpoint=True)
    xb = rng.integers(-2 ** 30 + 1, 2 ** 30, (3, n_t, n_sel))
    e = rng.integers(-12, 9, (3, n_t)).astype(np.int32)
    return xa, xb, e


def _run_project(eng, xa, xb, e, impl, rows_alloc=None):
    from psa_b200 import _lib
    rows, n_sel = xa.shape
    n_t = xb.shape[1]
    pitch = -(-n_sel // 64) * 64
    rows_alloc = rows_alloc or rows
    ad = np.zeros((4, rows_alloc, pitch), np.int8)
    ad[:, :rows, :n_sel] = M.balanced_digits(xa)
    bd = np.zeros((3, 4, n_t, pitch), np.int8)
    for pol in range(3):
        bd[pol, :, :, :n_sel] = M.balanced_digits(xb[pol])
    ldp = -(-n_t // 4) * 4
    P = torch.full((rows, 3, ldp), float("nan"), dtype=torch.float32, device=eng.device)
    eng.project(dev(eng, ad), rows, rows_alloc, dev(eng, bd), dev(eng, e), n_t, n_sel, pitch, P, ldp, impl=impl)
    torch.cuda.synchronize()
    return P.cpu().numpy()[:, :, :n_t]


SHAPES = [(2, 16, 64), (24, 128, 64), (128, 128, 256), (200, 300, 1000), (130, 77, 65), (256, 512, 4096)]


@pytest.mark.parametrize("rows,n_t,n_sel", SHAPES)
def test_project_simt_matches_integer_model(eng, rows, n_t, n_sel):
    from psa_b200 import _lib
    xa, xb, e = _proj_inputs(np.random.default_rng(rows + n_t), rows, n_t, n_sel)
    got = _run_project(eng, xa, xb, e, _lib.PROJECT_SIMT, rows_alloc=rows + 10)
    np.testing.assert_array_equal(got, M.project(xa, xb, e))


@pytest.mark.parametrize("rows,n_t,n_sel", SHAPES + [(400, 1024, 2048), (16, 256, 128)])
def test_project_tensor_matches_integer_model(eng, rows, n_t, n_sel):
    """tcgen05 cta_group::2 kernel (two CTAs share one 256-frame tile): bit-exact against the integer model."""
    from psa_b200 import _lib
    xa, xb, e = _proj_inputs(np.random.default_rng(rows * 5 + n_t), rows, n_t, n_sel)
    got = _run_project(eng, xa, xb, e, _lib.PROJECT_TENSOR, rows_alloc=rows + 10)
    np.testing.assert_array_equal(got, M.project(xa, xb, e))


def test_projection_by_frame_ranges_is_bit_identical(eng):
    """psa_project_rows: projecting a trajectory range by range (ragged ranges, not tile-aligned) gives the bits of the
    one-shot call - what lets a multi-GPU run project the frames that have arrived while the rest is in flight."""
    from psa_b200 import _lib
    rows, n_t, n_sel = 200, 1000, 300
    xa, xb, e = _proj_inputs(np.random.default_rng(77), rows, n_t, n_sel)
    want = M.project(xa, xb, e)
    pitch = -(-n_sel // 64) * 64
    ad = np.zeros((4, rows, pitch), np.int8)
    ad[:, :, :n_sel] = M.balanced_digits(xa)
    bd = np.zeros((3, 4, n_t, pitch), np.int8)
    for pol in range(3):
        bd[pol, :, :, :n_sel] = M.balanced_digits(xb[pol])
    ad_d, bd_d, e_d = dev(eng, ad), dev(eng, bd), dev(eng, e)
    for impl in (_lib.PROJECT_TENSOR, _lib.PROJECT_SIMT):
        P = torch.full((rows, 3, n_t), float("nan"), dtype=torch.float32, device=eng.device)
        for t0, t1 in ((700, 1000), (0, 256), (256, 700)):                   # any order
            eng.project(ad_d, rows, rows, bd_d, e_d, n_t, n_sel, pitch, P, n_t, impl=impl, t_range=(t0, t1))
        np.testing.assert_array_equal(P.cpu().numpy(), want)


@pytest.mark.parametrize("bounds", [(0, 130, 130, 296, 400), (0, 400), (0, 2, 398, 400), (0, 0, 144, 400, 400)])
def test_projection_routed_by_row_ranges_is_bit_identical(eng, bounds):
    """psa_project_routed: one launch, every row range stored as rows 0.. of its own destination (here: separate local
    buffers standing in for the peers' IPC-mapped ones, each at a frame offset inside a wider row) - the bits of the
    one-destination call.  Boundaries inside a 16-row store group, a 2-row owner and empty owners included."""
    import ctypes
    from psa_b200 import _lib
    rows, n_t, n_sel = 400, 600, 200
    xa, xb, e = _proj_inputs(np.random.default_rng(len(bounds)), rows, n_t, n_sel)
    want = M.project(xa, xb, e)
    pitch = -(-n_sel // 64) * 64
    ad = np.zeros((4, rows + 6, pitch), np.int8)
    ad[:, :rows, :n_sel] = M.balanced_digits(xa)
    bd = np.zeros((3, 4, n_t, pitch), np.int8)
    for pol in range(3):
        bd[pol, :, :, :n_sel] = M.balanced_digits(xb[pol])
    ad_d, bd_d, e_d = dev(eng, ad), dev(eng, bd), dev(eng, e)
    ldp, f0 = 1024, 212                                             # the owner's row pitch and this rank's first frame
    n_dest = len(bounds) - 1
    bufs = [torch.full((max(bounds[q + 1] - bounds[q], 1), 3, ldp), float("nan"), dtype=torch.float32, device=eng.device)
            for q in range(n_dest)]
    dests = (ctypes.c_void_p * n_dest)(*[b.data_ptr() + 4 * f0 for b in bufs])
    begin = (ctypes.c_int64 * (n_dest + 1))(*bounds)
    _lib.call("psa_project_routed", ad_d.data_ptr(), rows, rows + 6, bd_d.data_ptr(), e_d.data_ptr(), n_t, n_sel, pitch,
              ctypes.addressof(dests), ctypes.addressof(begin), n_dest, ldp, eng.stream())
    for q in range(n_dest):
        got = bufs[q].cpu().numpy()
        r0, r1 = bounds[q], bounds[q + 1]
        np.testing.assert_array_equal(got[:r1 - r0, :, f0:f0 + n_t], want[r0:r1])
        assert np.isnan(got[:, :, :f0]).all() and np.isnan(got[:, :, f0 + n_t:]).all()      # nothing outside the range
    # a bad table is refused before anything is launched
    begin_bad = (ctypes.c_int64 * (n_dest + 1))(*([0] * n_dest + [rows - 1]))
    with pytest.raises(ValueError):
        _lib.call("psa_project_routed", ad_d.data_ptr(), rows, rows + 6, bd_d.data_ptr(), e_d.data_ptr(), n_t, n_sel, pitch,
                  ctypes.addressof(dests), ctypes.addressof(begin_bad), n_dest, ldp, eng.stream())


def test_first_chunk_follows_an_arrival_schedule(gold_si):
    """Digit planes installed with an arrival schedule (frame ranges + events, as the pipelined multi-GPU exchange does):
    the first k-chunk is projected range by range, later chunks in one launch - same bits as the plain path."""
    calc = _calc(gold_si)
    kv = np.concatenate([gold_si["kpath_110_vecs"]] * 12)                   # 144 k-points: more than one chunk below
    plain = calc.calculate(np.zeros(len(kv)), kv, k_chunk_size=500).sed
    dtraj = calc.device_trajectory
    idx_dev, n_sel, pitch, dig, expo = dtraj.group(None, False)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(calc.engine.device))
    dtraj.install_group(None, False, dig, expo, arrivals=[(128, 256, None), (0, 64, ev), (64, 128, None)])
    again = calc.calculate(np.zeros(len(kv)), kv, k_chunk_size=100).sed      # 2 chunks of 72: first one by ranges
    np.testing.assert_array_equal(again, plain)
    assert dtraj.pop_arrivals(dig) is None                                   # consumed


@pytest.mark.parametrize("pinned", [True, False])
def test_streamed_ingest_is_bit_identical(monkeypatch, pinned):
    """One GPU, velocities still in host memory: frame ranges are copied + digitised on a side stream while the leading
    k-chunks are projected range by range (engine.IngestStream).  Same bits as upload-then-compute, for page-locked and
    pageable sources, one and two atom groups, several streamed chunks and a ragged last range."""
    from psa_b200 import SEDCalculator, Trajectory
    from psa_b200 import engine as E
    spec = synth.si_spec("stream", 5, 2304 + 100, seed=5)                # 1000 atoms, 2404 frames: 10 ranges, last ragged
    traj = spec.trajectory()
    pos, vel = traj.positions, traj.velocities
    if pinned:
        pos_t, vel_t = torch.from_numpy(pos).pin_memory(), torch.from_numpy(vel).pin_memory()
        traj = Trajectory(pos_t.numpy(), vel_t.numpy(), traj.types, traj.timesteps, traj.box_matrix, traj.box_lengths,
                          traj.box_tilts, traj.dt_ps)
    calc = SEDCalculator(traj, *spec.cells)
    mags, kv = calc.get_k_path([1, 1, 0], 2.0, 150)
    cases = [dict(summation_mode="coherent", k_chunk_size=40),                              # 4 chunks of <= 40
             dict(summation_mode="incoherent", basis_atom_types=[1, 2], k_chunk_size=64),   # two groups
             dict(summation_mode="coherent", basis_atom_indices=list(range(0, 1000, 3)))]   # gathered selection, 1 chunk
    monkeypatch.setenv("PSA_B200_STREAM_INGEST", "0")
    want = []
    for kw in cases:
        calc.release_device_memory()
        want.append(calc.calculate(mags, kv, **kw).sed)
    monkeypatch.setenv("PSA_B200_STREAM_INGEST", "1")
    monkeypatch.setattr(E, "_STREAM_MIN_BYTES", 1 << 20)
    monkeypatch.setattr(E, "_UPLOAD_CHUNK_BYTES", 1 << 20)               # several staging pieces per range (pageable)
    for rate in (2.7e13, 2.7e10):                                        # every chunk / only the first one streamed
        monkeypatch.setattr(E, "_STREAM_PROJECT_RATE", rate)
        for kw, w in zip(cases, want):
            calc.release_device_memory()
            launches = calc.engine.launches
            got = calc.calculate(mags, kv, **kw).sed
            np.testing.assert_array_equal(got, w)
            assert calc.device_trajectory._dev.get("vel") is not None    # the raw array ended up resident as usual
            assert calc.engine.launches - launches > 10                   # range-by-range launches really happened
            np.testing.assert_array_equal(calc.calculate(mags, kv, **kw).sed, w)      # cached planes afterwards


def test_project_extreme_digits_no_overflow(eng):
    """Worst-case digits (every product at its maximum) over a full 32768-atom pass stay exact."""
    from psa_b200 import _lib
    n_sel, rows, n_t = 32768 + 64, 2, 16
    worst = -128 * (1 + 256 + 65536) - 64 * 2 ** 24            # digits (-128, -128, -128, -64)
    xa = np.full((rows, n_sel), worst, np.int64)
    xb = np.full((3, n_t, n_sel), worst, np.int64)
    xa[1] = -worst - 2 * 128 * (1 + 256 + 65536)                 # digits (-128, -128, -128, +64)
    e = np.zeros((3, n_t), np.int32)
    want = M.project(xa, xb, e)
    for impl in (_lib.PROJECT_SIMT, _lib.PROJECT_TENSOR):
        np.testing.assert_array_equal(_run_project(eng, xa, xb, e, impl), want)


def test_project_two_pass_accumulation(eng):
    from psa_b200 import _lib
    rows, n_t, n_sel = 6, 32, 40000
    xa, xb, e = _proj_inputs(np.random.default_rng(99), rows, n_t, n_sel)
    want = M.project(xa, xb, e)
    for impl in (_lib.PROJECT_SIMT, _lib.PROJECT_TENSOR):
        np.testing.assert_array_equal(_run_project(eng, xa, xb, e, impl), want)


# ------------------------------------------------------------------ FFT + assembly
@pytest.mark.parametrize("n_t", [32, 64, 128, 1024, 2048, 8192, 16384, 32768, 65536])
def test_fft_coherent(eng, n_t):
    rng = np.random.default_rng(n_t)
    n_k, n_k_total, k_off = 3, 5, 1
    ldp = n_t
    P = rng.standard_normal((2 * n_k, 3, ldp)).astype(np.float32)
    out = torch.zeros((n_t, n_k_total, 3), dtype=torch.complex64, device=eng.device)
    eng.fft_sed(dev(eng, P), 1, P.size, n_k, n_t, ldp, 0, out, n_k_total, k_off)
    got = out.cpu().numpy()
    z = (P[0::2].astype(np.float64) + 1j * P[1::2].astype(np.float64))          # (n_k, 3, n_t)
    want = (np.fft.fft(z, axis=-1) / n_t).transpose(2, 0, 1)
    scale = np.abs(want).max()
    assert np.abs(got[:, k_off:k_off + n_k, :] - want).max() < 1.5e-7 * scale
    assert not got[:, 0, :].any() and not got[:, 4, :].any()                     # untouched k columns


@pytest.mark.parametrize("n_t", [64, 2048, 32768])
def test_fft_incoherent(eng, n_t):
    rng = np.random.default_rng(n_t + 1)
    n_k, groups = 4, 3
    P = rng.standard_normal((groups, 2 * n_k, 3, n_t)).astype(np.float32)
    out = torch.zeros((n_t, n_k), dtype=torch.float32, device=eng.device)
    eng.fft_sed(dev(eng, P), groups, P[0].size, n_k, n_t, n_t, 1, out, n_k, 0)
    z = P[:, 0::2].astype(np.float64) + 1j * P[:, 1::2].astype(np.float64)      # (g, n_k, 3, n_t)
    want = (np.abs(np.fft.fft(z, axis=-1) / n_t) ** 2).sum(axis=(0, 2)).T
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=2e-5, atol=1e-6 * want.max())


# 4 * 2^a 3^b 5^c -> mixed-radix core (final blocks of 16 / 8 / 4 points, load-time split by R = 3, 5, 25 ...);
# anything else (odd lengths, a large prime factor) -> Bluestein
@pytest.mark.parametrize("n_t", [1, 2, 3, 4, 7, 8, 12, 16, 20, 28, 48, 60, 80, 250, 360, 1000, 1200, 3000, 3001, 5000,
                                 8191, 10000, 12288, 20000, 50000])
def test_fft_any_length_bluestein(eng, n_t):
    """Frame counts that are not a power of two (the reference accepts any n_t): mixed radix or Bluestein."""
    rng = np.random.default_rng(n_t)
    n_k, ldp = 2, -(-n_t // 4) * 4
    P = np.zeros((2 * n_k, 3, ldp), np.float32)
    P[:, :, :n_t] = rng.standard_normal((2 * n_k, 3, n_t))
    out = torch.zeros((n_t, n_k, 3), dtype=torch.complex64, device=eng.device)
    eng.fft_sed(dev(eng, P), 1, P.size, n_k, n_t, ldp, 0, out, n_k, 0)
    z = P[0::2, :, :n_t].astype(np.float64) + 1j * P[1::2, :, :n_t].astype(np.float64)
    want = (np.fft.fft(z, axis=-1) / n_t).transpose(2, 0, 1)
    # the chirped spectrum stays in float64 between Bluestein's two legs: same tolerance as the direct lengths
    assert np.abs(out.cpu().numpy() - want).max() < 1.5e-7 * np.abs(want).max()
    # incoherent assembly over two groups through the same path
    Pg = np.stack([P, P[::-1].copy()])
    acc = torch.zeros((n_t, n_k), dtype=torch.float32, device=eng.device)
    eng.fft_sed(dev(eng, Pg), 2, P.size, n_k, n_t, ldp, 1, acc, n_k, 0)
    zg = Pg[:, 0::2, :, :n_t].astype(np.float64) + 1j * Pg[:, 1::2, :, :n_t].astype(np.float64)
    want_i = (np.abs(np.fft.fft(zg, axis=-1) / n_t) ** 2).sum(axis=(0, 2)).T
    np.testing.assert_allclose(acc.cpu().numpy(), want_i, rtol=5e-5, atol=2e-6 * want_i.max())


def test_fft_rejects_unsupported_length(eng):
    from psa_b200 import _lib
    assert _lib.load().psa_fft_plan_bytes(0) == -1 and _lib.load().psa_fft_plan_bytes(2 ** 19 + 1) == -1
    with pytest.raises(NotImplementedError):
        eng.fft_plan(2 ** 19 + 1)


@pytest.mark.parametrize("n_t,n_k,n_k_total,k_off", [(8192, 70, 70, 0), (16384, 54, 60, 5), (32768, 23, 23, 0),
                                                     (16384, 1, 1, 0), (8192, 16, 16, 0)])
def test_fft_four_step_many_columns(eng, n_t, n_k, n_k_total, k_off):
    """The four-step kernel (8192 / 16384 / 32768 frames) on enough columns to wrap its ring of L2-resident group
    slots several times, a ragged last group of columns, and a result wider than the call's k-range."""
    rng = np.random.default_rng(n_t + n_k)
    P = rng.standard_normal((2 * n_k, 3, n_t)).astype(np.float32)
    t = np.arange(n_t)
    P[0::2] += 40 * np.cos(2 * np.pi * 123 * t / n_t).astype(np.float32)       # a strong line over the noise floor
    P[1::2] += 40 * np.sin(2 * np.pi * 123 * t / n_t).astype(np.float32)
    out = torch.full((n_t, n_k_total, 3), float("nan"), dtype=torch.complex64, device=eng.device)
    for _ in range(2):                                                           # second call reuses plan + workspace
        eng.fft_sed(dev(eng, P), 1, P.size, n_k, n_t, n_t, 0, out, n_k_total, k_off)
    got = out.cpu().numpy()
    z = P[0::2].astype(np.float64) + 1j * P[1::2].astype(np.float64)
    want = (np.fft.fft(z, axis=-1) / n_t).transpose(2, 0, 1)
    # float64 transform, one float32 rounding: every bin is accurate relative to ITS OWN magnitude, like NumPy's
    # complex64 FFT (which is the float64 transform rounded once)
    sel = got[:, k_off:k_off + n_k, :]
    assert np.abs(sel - want).max() < 1.5e-7 * np.abs(want).max()
    weak = np.abs(want) > 1e-9 * np.abs(want).max()
    assert (np.abs(sel - want)[weak] / np.abs(want)[weak]).max() < 2e-7
    assert (sel == want.astype(np.complex64)).mean() > 0.999                     # = the float64 transform rounded once
    untouched = np.ones(n_k_total, bool)
    untouched[k_off:k_off + n_k] = False
    assert np.isnan(got[:, untouched, :].real).all()


@pytest.mark.parametrize("n_t", [4096, 8192])
def test_fft_window_is_applied_in_float64(eng, n_t):
    n_k = 2
    rng = np.random.default_rng(3)
    P = rng.standard_normal((2 * n_k, 3, n_t)).astype(np.float32)
    w = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n_t) / n_t)).astype(np.float32)
    out = torch.zeros((n_t, n_k, 3), dtype=torch.complex64, device=eng.device)
    eng.fft_sed(dev(eng, P), 1, P.size, n_k, n_t, n_t, 0, out, n_k, 0, window=dev(eng, w))
    z = (P[0::2].astype(np.float64) + 1j * P[1::2].astype(np.float64)) * w.astype(np.float64)
    want = (np.fft.fft(z, axis=-1) / n_t).transpose(2, 0, 1)
    assert np.abs(out.cpu().numpy() - want).max() < 1.5e-7 * np.abs(want).max()


def test_non_finite_samples_poison_their_frame(eng):
    """A NaN / Inf sample must not be digitised into a plausible finite value: its frame's exponent is the poison
    value and every projection of that frame comes out NaN (the reference propagates non-finite input as well)."""
    from psa_b200 import _lib
    rng = np.random.default_rng(8)
    n_t, n_a = 72, 128
    data = rng.standard_normal((n_t, n_a, 3)).astype(np.float32)
    data[5, 17, 1] = np.nan
    data[9, 3, 2] = np.inf
    data[11, 0, 0] = 3e38
    dig, expo, pitch = eng.digitize(dev(eng, data), None, None, n_a)
    e = expo.cpu().numpy()
    POISON = 0x40000000
    assert e[1, 5] == POISON and e[2, 9] == POISON and e[0, 11] == POISON
    assert (e == POISON).sum() == 3
    rows = 4
    ad = np.zeros((4, rows, pitch), np.int8)
    ad[3, :, :n_a] = 64                                       # phase value 1.0 everywhere
    P = torch.zeros((rows, 3, n_t), dtype=torch.float32, device=eng.device)
    for impl in (_lib.PROJECT_TENSOR, _lib.PROJECT_SIMT):
        eng.project(dev(eng, ad), rows, rows, dig, expo, n_t, n_a, pitch, P, n_t, impl=impl)
        got = P.cpu().numpy()
        assert np.isnan(got[:, 1, 5]).all() and np.isnan(got[:, 2, 9]).all() and np.isnan(got[:, 0, 11]).all()
        finite = np.ones_like(got, bool)
        finite[:, 1, 5] = finite[:, 2, 9] = finite[:, 0, 11] = False
        assert np.isfinite(got[finite]).all()
        np.testing.assert_allclose(got[0, 0, 0], data[0, :, 0].astype(np.float64).sum(), rtol=1e-6)


def test_digitize_weight_matches_float32_product(eng):
    rng = np.random.default_rng(12)
    n_t, n_a = 9, 204
    data = (rng.standard_normal((n_t, n_a, 3)) * 2.0).astype(np.float32)
    wgt = np.sqrt(rng.uniform(1.0, 200.0, n_a)).astype(np.float32)
    idx = np.arange(0, n_a, 3).astype(np.int32)
    for sel in (None, idx):
        want = (data * wgt[None, :, None]).astype(np.float32)
        want = want if sel is None else want[:, sel]
        n_sel = want.shape[1]
        dig, expo, pitch = eng.digitize(dev(eng, data), None, None if sel is None else dev(eng, sel), n_sel,
                                        weight=dev(eng, wgt))
        x_ref, e_ref = M.digitize(want)
        np.testing.assert_array_equal(expo.cpu().numpy(), e_ref)
        d = dig.cpu().numpy()
        for pol in range(3):
            np.testing.assert_array_equal(d[pol, :, :, :n_sel], M.balanced_digits(x_ref[pol]))


# ------------------------------------------------------------------ element-wise kernels
def test_chiral_and_intensity_kernels(eng, gold_gr):
    sed = gold_gr["sed_coh"]
    d = dev(eng, sed)
    inten = eng.intensity(d).cpu().numpy()
    np.testing.assert_allclose(inten, O.intensity(sed), rtol=1e-6)
    n = sed.shape[0] * sed.shape[1]
    flat = torch.view_as_real(d).view(-1, 2)
    strong = O.intensity(sed) > 1e-8 * O.intensity(sed).max()
    for axis, (i, j) in (("x", (1, 2)), ("y", (0, 2)), ("z", (0, 1))):
        out = eng.empty(sed.shape[:2], torch.float32)
        eng.chiral_phase(flat[i:], flat[j:], n, 3, 3, "C", out)
        got, want = out.cpu().numpy(), gold_gr[f"phase_C_{axis}"]
        assert np.abs(got - want)[strong].max() < 2e-5
        assert np.abs(got).max() <= np.pi / 2 + 1e-6
    for opt in "AB":
        out = eng.empty(sed.shape[:2], torch.float32)
        eng.chiral_phase(flat[0:], flat[1:], n, 3, 3, opt, out)
        assert np.median(np.abs(out.cpu().numpy() - gold_gr[f"phase_{opt}_z"])) < 1e-6


# ------------------------------------------------------------------ whole path vs the real reference's outputs
def _calc(g, **kw):
    from psa_b200 import SEDCalculator, Trajectory
    box = g["box_matrix"]
    traj = Trajectory(g["positions"], g["velocities"], g["types"], np.arange(g["positions"].shape[0]), box,
                      np.diag(box).copy(), np.zeros(3, np.float32), float(g["dt_ps"]))
    return SEDCalculator(traj, *[int(c) for c in g["cells"]], **kw)


GOLD_CASES = {
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
    "inc_indices": ("kpath_100_vecs", dict(basis_atom_indices=[[0, 1, 5, 9], [2, 3, 40]], summation_mode="incoherent")),
    "coh_indices_union": ("kpath_100_vecs", dict(basis_atom_indices=[[0, 1, 5, 9], [2, 3, 5]])),
    "coh_indices_flat_dup": ("kpath_100_vecs", dict(basis_atom_indices=[3, 1, 1, 20])),
    "coh_indices_ndarray": ("kpath_100_vecs", dict(basis_atom_indices=np.array([4, 8, 15, 16, 23, 42]))),
    "inc_all": ("kpath_100_vecs", dict(summation_mode="incoherent")),
    "kgrid_xy": ("kgrid_xy_vecs", dict(k_grid_shape=(4, 3))),
}


@pytest.mark.parametrize("name", sorted(GOLD_CASES))
def test_calculate_matches_reference_outputs(gold_si, name):
    kkey, kw = GOLD_CASES[name]
    calc = _calc(gold_si)
    kv = gold_si[kkey]
    res = calc.calculate(np.zeros(len(kv), np.float32), kv, **kw)
    ref = gold_si[f"sed_{name}"]
    assert res.sed.shape == ref.shape and res.sed.dtype == ref.dtype
    assert res.is_complex == bool(gold_si[f"cplx_{name}"])
    np.testing.assert_array_equal(res.freqs, gold_si["freqs"])
    # float32 noise of the reference itself is ~1e-6 of the peak at this size
    if res.is_complex:
        assert np.abs(res.sed - ref).max() < 2e-6 * np.abs(ref).max()
    else:
        np.testing.assert_allclose(res.sed, ref, rtol=0, atol=4e-6 * ref.max())
    assert res["sed"] is res.sed and res.k_grid_shape == kw.get("k_grid_shape")


def test_displacement_mode_matches_reference(gold_si):
    calc = _calc(gold_si, use_displacements=True)
    kv = gold_si["kpath_100_vecs"]
    ref = gold_si["sed_disp_coh_all_100"]
    res = calc.calculate(gold_si["kpath_100_mags"], kv)
    assert np.abs(res.sed - ref).max() < 2e-6 * np.abs(ref).max()
    ref2 = gold_si["sed_disp_inc_types12"]
    res2 = calc.calculate(gold_si["kpath_110_mags"], gold_si["kpath_110_vecs"], basis_atom_types=[1, 2],
                          summation_mode="incoherent")
    np.testing.assert_allclose(res2.sed, ref2, rtol=0, atol=4e-6 * ref2.max())


def test_api_errors_and_edge_cases(gold_si):
    calc = _calc(gold_si)
    kv = gold_si["kpath_100_vecs"]
    with pytest.raises(ValueError):
        calc.calculate(np.zeros(len(kv)), kv, summation_mode="both")
    with pytest.raises(ValueError):
        calc.calculate(np.zeros(len(kv)), kv, basis_atom_indices=[0, 10 ** 6])
    empty = calc.calculate(np.zeros(0, np.float32), np.zeros((0, 3), np.float32))
    assert empty.sed.shape == (256, 0, 3) and empty.sed.dtype == np.complex64
    one = calc.calculate(np.zeros(1, np.float32), kv[3:4])
    assert np.abs(one.sed[:, 0] - gold_si["sed_coh_all_100"][:, 3]).max() < 2e-6 * np.abs(gold_si["sed_coh_all_100"]).max()
    # intensity property of the result type (reference sed.py:22-24)
    np.testing.assert_allclose(calc.calculate(np.zeros(len(kv)), kv).intensity, gold_si["intensity_coh_all_100"],
                               rtol=1e-4, atol=1e-6 * gold_si["intensity_coh_all_100"].max())


def test_chunk_invariance_is_bitwise(gold_si):
    calc = _calc(gold_si)
    kv = gold_si["kpath_110_vecs"]
    a = calc.calculate(np.zeros(len(kv)), kv, k_chunk_size=500).sed
    b = calc.calculate(np.zeros(len(kv)), kv, k_chunk_size=5).sed
    np.testing.assert_array_equal(a, b)


def test_streamed_chunks_match_single_chunk(gold_si):
    """A k-set longer than one chunk is streamed chunk by chunk into the pinned host result on a side
    stream; ragged last chunk, both result kinds, and the device-resident path give the same bits."""
    calc = _calc(gold_si)
    kv = gold_si["kpath_110_vecs"]
    assert len(kv) % 7 != 0
    for kwargs in (dict(), dict(basis_atom_types=[1, 2], summation_mode="incoherent")):
        whole = calc.calculate(np.zeros(len(kv)), kv, k_chunk_size=500, **kwargs)
        streamed = calc.calculate(np.zeros(len(kv)), kv, k_chunk_size=7, **kwargs)
        assert streamed.sed.shape == whole.sed.shape and streamed.is_complex == whole.is_complex
        np.testing.assert_array_equal(streamed.sed, whole.sed)
        dev, _, _ = calc._calculate_device(kv, None, kwargs.get("basis_atom_types"),
                                           kwargs.get("summation_mode", "coherent"), 7)
        np.testing.assert_array_equal(dev.cpu().numpy(), whole.sed)


def test_intensity_map_on_device(gold_si):
    """N3: sum_pol |S|^2 straight from the FFT kernel's epilogue, cropped to 0 <= f <= max_freq, streamed or not."""
    calc = _calc(gold_si)
    kv = gold_si["kpath_110_vecs"]
    full = calc.calculate(np.zeros(len(kv)), kv)
    want = O.intensity(full.sed)
    got = calc.calculate_intensity(np.zeros(len(kv)), kv)
    assert got.sed.shape == want.shape and got.sed.dtype == np.float32 and not got.is_complex
    np.testing.assert_allclose(got.sed, want, rtol=2e-6, atol=1e-9 * want.max())
    np.testing.assert_array_equal(got.freqs, full.freqs)
    f_max = float(full.freqs[full.freqs > 0][len(full.freqs) // 5])
    keep = (full.freqs >= 0) & (full.freqs <= f_max)
    for chunk in (500, 7):                                  # single chunk / streamed ragged chunks
        crop = calc.calculate_intensity(np.zeros(len(kv)), kv, max_freq=f_max, k_chunk_size=chunk)
        np.testing.assert_array_equal(crop.freqs, full.freqs[keep])
        np.testing.assert_array_equal(crop.sed, got.sed[keep])
    inc = calc.calculate(np.zeros(len(kv)), kv, basis_atom_types=[1, 2], summation_mode="incoherent")
    inc_map = calc.calculate_intensity(np.zeros(len(kv)), kv, basis_atom_types=[1, 2], summation_mode="incoherent")
    np.testing.assert_array_equal(inc_map.sed, inc.sed)


def test_device_side_consumers_match_numpy(eng, gold_si):
    """N3: intensity scaling, nan-aware global range and np.percentile limits computed on the device."""
    from psa_b200 import consumers
    rng = np.random.default_rng(23)
    x = (rng.standard_normal(200_003) ** 2 * 10.0 ** rng.integers(-9, 3, 200_003)).astype(np.float32)
    x[::1000] = np.nan
    x[5::5000] = np.inf
    x[7] = -np.inf
    x[11:20] = 0.0
    x[30:33] = -2.5
    d = dev(eng, x)
    lo, hi, n = consumers.nan_range(eng, d)
    assert lo == np.nanmin(x) and hi == np.nanmax(x) and n == int(np.isfinite(x).sum())
    valid = x[np.isfinite(x)]
    for qs in ((1.0, 99.0), (0.0, 100.0), (50.0,), (33.3, 99.99)):
        got = consumers.percentiles(eng, d, qs)
        want = [float(np.percentile(valid, q)) for q in qs]
        assert got == want, (qs, got, want)                              # NumPy's float32 quantile rule, bit for bit
    ranks = [0, 1, len(valid) // 2, len(valid) - 1]
    np.testing.assert_array_equal(consumers.order_statistics(eng, d, ranks), np.sort(valid)[ranks])
    pos = np.abs(rng.standard_normal(5000)).astype(np.float32)
    pos[3] = 0.0
    for name, fn in (("log", lambda a: np.log10(np.maximum(a, 1e-12))), ("sqrt", lambda a: np.sqrt(np.maximum(a, 0))),
                     ("dsqrt", lambda a: np.sqrt(np.sqrt(np.maximum(a, 0)))), ("linear", lambda a: a)):
        t = dev(eng, pos)
        consumers.scale_intensity(eng, t, name)
        np.testing.assert_allclose(t.cpu().numpy(), fn(pos), rtol=3e-7, atol=1e-7)
    # through the calculator: the GUI's globally scaled k-path heat map
    calc = _calc(gold_si)
    kv = gold_si["kpath_110_vecs"]
    lin = calc.calculate_intensity(np.zeros(len(kv)), kv)
    res = calc.calculate_intensity(np.zeros(len(kv)), kv, intensity_scale="dsqrt", vmin_percentile=1.0,
                                   vmax_percentile=99.0, global_range=True)
    want = np.sqrt(np.sqrt(np.maximum(lin.sed, 0)))
    np.testing.assert_allclose(res.sed, want, rtol=3e-7)
    st = res.context["stats"]
    np.testing.assert_allclose([st["global_min"], st["global_max"]], [res.sed.min(), res.sed.max()], rtol=0)
    assert [st["vmin"], st["vmax"]] == [float(np.percentile(res.sed, q)) for q in (1.0, 99.0)]   # scalar q, like the plotter


def test_cli_batch_driver(gold_si, tmp_path):
    """N4: the YAML-driven batch run writes the reference's SED bundles, computes every direction once and reduces the
    global maximum on the device; a second run loads the bundles back instead of recomputing."""
    import json
    import yaml
    from psa_b200 import SED, cache, cli
    calc = _calc(gold_si)
    traj_file = tmp_path / "run.lammpstrj"
    cache.save_npy_cache(calc.traj, traj_file)
    cfg = {"md_system": {"dt": float(gold_si["dt_ps"]), "nx": 2, "ny": 2, "nz": 2, "lattice_parameter": synth.SI_A},
           "sed_calculation": {"directions": [[1, 0, 0], [1, 1, 0]], "n_kpoints": 12, "bz_coverage": 4.0},
           "ised": {"apply": True, "k_path": {"direction": "x", "n_points": 9, "bz_coverage": 1.0},
                    "target_point": {"k_value": float(gold_si["ised_k_target"]), "w_value_thz": float(gold_si["ised_w_target"])},
                    "reconstruction": {"rescaling_factor": 0.5, "num_animation_timesteps": 8}}}
    (tmp_path / "cfg.yaml").write_text(yaml.safe_dump(cfg))
    out = tmp_path / "out"
    argv = ["--trajectory", str(traj_file), "--config", str(tmp_path / "cfg.yaml"), "--output-dir", str(out)]
    assert cli.main(argv) == 0
    got = SED.load(out / "sed_data_regular_1.00_0.00_0.00")
    ref = gold_si["sed_coh_all_100"]
    assert got.sed.shape == ref.shape and np.abs(got.sed - ref).max() < 4e-6 * np.abs(ref).max()
    np.testing.assert_allclose(got.k_vectors, gold_si["kpath_100_vecs"], rtol=1e-6)     # 2 pi / a given vs derived from b1
    # [110] with an explicit lattice parameter (what the reference's CLI always passes, cli.py:93, 129): same bits as the
    # direct call with that k-path
    got110 = SED.load(out / "sed_data_regular_1.00_1.00_0.00")
    km, kv = calc.get_k_path([1, 1, 0], 4.0, 12, synth.SI_A)
    np.testing.assert_array_equal(got110.sed, calc.calculate(km, kv).sed)
    np.testing.assert_array_equal(got110.k_points, km)
    summary = json.loads((out / "summary.json").read_text())
    peaks = [float(np.max(SED.load(out / f"sed_data_regular_{lbl}").intensity)) for lbl in ("1.00_0.00_0.00", "1.00_1.00_0.00")]
    np.testing.assert_allclose([d["max_intensity"] for d in summary["directions"]], peaks, rtol=1e-5)
    np.testing.assert_allclose(summary["global_max_intensity"], max(peaks), rtol=1e-5)
    dump = (out / "ised_motion.dump").read_text().splitlines()
    assert dump[0] == "ITEM: TIMESTEP" and len(dump) == 8 * (9 + len(gold_si["types"]))
    assert cli.main(argv) == 0                                             # second run: loaded, not recomputed
    again = json.loads((out / "summary.json").read_text())
    assert all(d["loaded_from_cache"] for d in again["directions"])
    assert cli.main(argv + ["--chiral", "--output-dir", str(tmp_path / "chi")]) == 0
    chi = SED.load(tmp_path / "chi" / "sed_data_chiral_1.00_0.00_0.00")
    assert chi.phase is not None and chi.phase.shape == ref.shape[:2]


def test_chiral_sed_facade(gold_gr):
    calc = _calc(gold_gr)
    res = calc.calculate_chiral_sed([1, 0, 0], bz_coverage=4.0, n_k=10, chiral_axis="z")
    ref = gold_gr["sed_coh"]
    assert np.abs(res.sed - ref).max() < 2e-6 * np.abs(ref).max()
    strong = O.intensity(ref) > 1e-6 * O.intensity(ref).max()
    assert np.abs(res.phase - gold_gr["phase_C_z"])[strong].max() < 1e-3
    assert res.phase.dtype == np.float32 and res.phase.shape == ref.shape[:2]
    # the public method on host arrays
    ph = calc.calculate_chiral_phase(ref[:, :, 0], ref[:, :, 1], "C")
    assert np.abs(ph - gold_gr["phase_C_z"])[strong].max() < 2e-5
    with pytest.raises(ValueError):
        calc.calculate_chiral_phase(ref[:, :, 0], ref[:, :3, 1])


def test_kgrid_and_kpath_facades(gold_si):
    calc = _calc(gold_si)
    res = calc.calculate_kgrid_sed("xy", (-1.5, 2.0, -0.5, 1.0), 4, 3, 0.25)
    assert res.k_grid_shape == (4, 3) and res.k_points.size == 0
    ref = gold_si["sed_kgrid_xy"]
    assert np.abs(res.sed - ref).max() < 2e-6 * np.abs(ref).max()
    res = calc.calculate_kpath_sed([1, 1, 0], 4.0, 12, basis_atom_types=[1, 2], summation_mode="incoherent")
    np.testing.assert_allclose(res.sed, gold_si["sed_inc_types12"], rtol=0, atol=4e-6 * gold_si["sed_inc_types12"].max())
    assert not res.is_complex and res["is_complex"] is False


def test_ised_matches_reference(gold_si, tmp_path):
    calc = _calc(gold_si)
    common = dict(char_len_k_path=5.431, nk_on_path=9, bz_cov_ised=1.0, n_recon_frames=8)
    kt, wt = float(gold_si["ised_k_target"]), float(gold_si["ised_w_target"])
    out = calc.reconstruct([1, 0, 0], [(kt, wt)], basis_atom_types_ised=[1, 2], rescale_factor=0.5, **common)
    assert np.abs(out[0]["frames"] - gold_si["ised_types_float"]).max() < 2e-5
    out = calc.reconstruct([1, 0, 0], [(kt, wt)], rescale_factor="auto", **common)
    assert np.abs(out[0]["frames"] - gold_si["ised_all_auto"]).max() < 2e-5
    out = calc.reconstruct([1, 0, 0], [(0.3, wt * 0.5)], basis_atom_idx_ised=[[0, 1, 2, 3], [10, 11, 12]],
                           rescale_factor="auto", **common)
    assert np.abs(out[0]["frames"] - gold_si["ised_idx_auto"]).max() < 2e-5
    # batched call == single calls; dump file round trip
    both = calc.reconstruct([1, 0, 0], [(kt, wt), (0.3, wt * 0.5)], rescale_factor=0.5, **common)
    single = calc.reconstruct([1, 0, 0], [(0.3, wt * 0.5)], rescale_factor=0.5, **common)
    np.testing.assert_array_equal(both[1]["frames"], single[0]["frames"])
    dump = tmp_path / "m.dump"
    calc.ised([1, 0, 0], kt, wt, rescale_factor=0.5, dump_filepath=str(dump), **common)
    text = dump.read_text().splitlines()
    assert text[0] == "ITEM: TIMESTEP" and text[3] == str(len(gold_si["types"]))
    assert len(text) == 8 * (9 + len(gold_si["types"]))


def test_batched_ised_matches_oracle_groups_auto_and_overlap():
    """One batched launch over several (k, omega) points: two type groups, 'auto' and numeric rescale, and overlapping
    index groups (the reference's float32 running sum and running maximum, sed_calculator.py:494-524), against the
    oracle's restatement of the reference loop."""
    from psa_b200 import SEDCalculator
    from psa_b200 import groups as grp
    spec = synth.si_spec("ised", n_cells=3, n_frames=512, seed=41)
    traj = spec.trajectory(threads=1)
    calc = SEDCalculator(traj, *spec.cells)
    k_hat = np.array([1.0, 0.0, 0.0], np.float32)
    mags, kv = calc.get_k_path(k_hat, 1.0, 12, lat_param=synth.SI_A)
    freqs = np.fft.fftfreq(traj.n_frames, d=traj.dt_ps)
    targets = [(float(mags[3]), float(freqs[40])), (float(mags[7]), float(freqs[11])), (0.2, 3.3), (float(mags[3]), 9.0)]
    cases = [dict(basis_atom_types_ised=[1, 2], rescale_factor="auto"),
             dict(basis_atom_types_ised=[1, 2], rescale_factor=0.25),
             dict(basis_atom_idx_ised=[[0, 1, 2, 3, 4, 5], [4, 5, 6, 7, 7], [100, 101]], rescale_factor="auto"),
             dict(rescale_factor="auto")]
    for kw in cases:
        got = calc.reconstruct(k_hat, targets, synth.SI_A, nk_on_path=12, bz_cov_ised=1.0, n_recon_frames=16, **kw)
        groups = grp.resolve_ised_groups(traj.types, traj.n_atoms, kw.get("basis_atom_idx_ised"),
                                         kw.get("basis_atom_types_ised"))
        assert len(got) == len(targets)
        for (kt, wt), res in zip(targets, got):
            want = O.ised(traj.positions, traj.velocities, traj.types, traj.dt_ps, k_hat, mags, kv, kt, wt,
                          [np.asarray(g) for g in groups], rescale_factor=kw["rescale_factor"], n_frames=16)
            assert res["k_index"] == want["k_idx"] and res["w_index"] == want["w_idx"]
            scale = np.abs(want["frames"] - O.mean_positions(traj.positions)[None]).max()
            assert np.abs(res["frames"] - want["frames"]).max() < 2e-5 * max(scale, 1e-3) + 4e-6, kw
    # frames left on the device are the same bits as the host copies
    a = calc.reconstruct(k_hat, targets[:2], synth.SI_A, nk_on_path=12, n_recon_frames=16, keep_on_device=True)
    b = calc.reconstruct(k_hat, targets[:2], synth.SI_A, nk_on_path=12, n_recon_frames=16)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x["frames"].cpu().numpy(), y["frames"])


@pytest.mark.parametrize("n_t,n_a,gather", [(300, 64, False), (301, 37, False), (257, 50, True), (3, 5, False),
                                             (1200, 1000, False)])
def test_displacement_moments_every_load_path(eng, n_t, n_a, gather):
    """sum and sum of squares of the float32 displacements (the 'auto' rescale of iSED, sed_calculator.py:506-512):
    whole rows through 16-byte loads, whole rows of a length that is not a multiple of four, a gathered selection."""
    from psa_b200 import _lib
    rng = np.random.default_rng(n_t + n_a)
    pos = (rng.random((n_t, n_a, 3)) * 30 + rng.standard_normal((n_t, n_a, 3)) * 0.07).astype(np.float32)
    mean = np.mean(pos, axis=0, dtype=np.float32)
    idx = np.sort(rng.choice(n_a, size=n_a // 3, replace=False)).astype(np.int32) if gather else None
    sel = pos if idx is None else pos[:, idx]
    d = (sel - (mean if idx is None else mean[idx])[None]).astype(np.float64)
    out = torch.zeros(2, dtype=torch.float64, device=eng.device)
    idx_dev = None if idx is None else dev(eng, idx)
    pos_dev, mean_dev = dev(eng, pos), dev(eng, mean)              # named: the buffers must outlive the launch
    _lib.call("psa_disp_moments", pos_dev.data_ptr(), mean_dev.data_ptr(),
              None if idx is None else idx_dev.data_ptr(), n_t, n_a, n_a if idx is None else idx.size,
              out.data_ptr(), eng.stream())
    s1, s2 = out.cpu().tolist()
    assert abs(s1 - d.sum()) <= 1e-9 * np.abs(d).sum() + 1e-12
    assert abs(s2 - (d * d).sum()) <= 1e-12 * (d * d).sum()


def test_mass_weighting_and_window_extensions():
    """README-level keywords the shipped source lacks: defaults reproduce the unweighted, unwindowed
    reference bit for bit; when given they follow the oracle's definition (float32 sqrt(m) v, taper before the FFT)."""
    from psa_b200 import SEDCalculator
    spec = synth.si_spec("mw", n_cells=2, n_frames=1024, seed=5)
    traj = spec.trajectory(threads=1)
    plain = SEDCalculator(traj, *spec.cells)
    mags, kv = plain.get_k_path([1, 1, 0], 4.0, 20)
    base = plain.calculate(mags, kv)
    same = SEDCalculator(traj, *spec.cells, masses=None, window=None).calculate(mags, kv)
    np.testing.assert_array_equal(base.sed, same.sed)
    ones = SEDCalculator(traj, *spec.cells, masses=np.ones(traj.n_atoms), window="rectangular").calculate(mags, kv)
    np.testing.assert_array_equal(base.sed, ones.sed)                     # sqrt(1) = 1 is exact
    masses = {1: 28.0855, 2: 72.63}
    wgt = np.sqrt(np.where(traj.types == 1, masses[1], masses[2])).astype(np.float32)
    win = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(1024) / 1024)).astype(np.float32)
    for kw, okw in ((dict(masses=masses), dict(weight=wgt)), (dict(window="hann"), dict(window=win)),
                    (dict(masses=wgt.astype(np.float64) ** 2, window=win), dict(weight=wgt, window=win))):
        got = SEDCalculator(traj, *spec.cells, **kw).calculate(mags, kv, basis_atom_types=[1, 2],
                                                               summation_mode="incoherent")
        want = O.calculate_fp64(traj.positions, traj.velocities, traj.types, traj.dt_ps, kv, basis_atom_types=[1, 2],
                                summation_mode="incoherent", **okw)
        assert not got.is_complex
        np.testing.assert_allclose(got.sed, want["sed"], rtol=2e-5, atol=2e-6 * want["sed"].max())
    with pytest.raises(ValueError):
        SEDCalculator(traj, *spec.cells, masses=np.ones(3))
    with pytest.raises(ValueError):
        SEDCalculator(traj, *spec.cells, window="kaiser")


def test_single_frame_trajectory(gold_si):
    """n_t = 1: the reference accepts it (a one-point FFT); the spectrum is the projection itself."""
    from psa_b200 import SEDCalculator, Trajectory
    g = gold_si
    box = g["box_matrix"]
    traj = Trajectory(g["positions"][:1], g["velocities"][:1], g["types"], np.arange(1), box, np.diag(box).copy(),
                      np.zeros(3, np.float32), float(g["dt_ps"]))
    res = SEDCalculator(traj, 2, 2, 2).calculate(g["kpath_100_mags"], g["kpath_100_vecs"])
    ref = O.calculate(traj.positions, traj.velocities, traj.types, traj.dt_ps, g["kpath_100_vecs"])
    assert res.sed.shape == ref["sed"].shape == (1, len(g["kpath_100_vecs"]), 3)
    assert np.abs(res.sed - ref["sed"]).max() < 2e-6 * np.abs(ref["sed"]).max()


def test_ised_reconstructor_facade(gold_si, tmp_path):
    from psa_b200 import iSEDReconstructor
    calc = _calc(gold_si)
    res = calc.calculate_kpath_sed([1, 0, 0], 1.0, 9, lat_param=5.431)
    rec = iSEDReconstructor(res)
    motion = rec.reconstruct_motion(float(gold_si["ised_k_target"]), float(gold_si["ised_w_target"]),
                                    n_frames=8, rescale_factor="auto")
    assert np.abs(motion - gold_si["ised_all_auto"]).max() < 2e-5
    rec.save_trajectory(motion, str(tmp_path / "x.dump"))
    assert (tmp_path / "x.dump").exists()


def test_odd_frame_count_matches_reference(gold_si):
    """250 frames: not a power of two -> Bluestein path, checked against the real reference's output."""
    from psa_b200 import SEDCalculator, Trajectory
    g = gold_si
    box = g["box_matrix"]
    traj = Trajectory(g["positions"][:250], g["velocities"][:250], g["types"], np.arange(250), box,
                      np.diag(box).copy(), np.zeros(3, np.float32), float(g["dt_ps"]))
    calc = SEDCalculator(traj, 2, 2, 2)
    res = calc.calculate(g["kpath_100_mags"], g["kpath_100_vecs"])
    ref = g["sed_odd250_coh_all_100"]
    assert res.sed.shape == ref.shape
    assert np.abs(res.sed - ref).max() < 4e-6 * np.abs(ref).max()
    np.testing.assert_array_equal(res.freqs, np.fft.fftfreq(250, d=float(g["dt_ps"])))


# ------------------------------------------------------------------ medium size: three-distance parity report
@pytest.mark.parametrize("n_frames", [2048, 3000])      # power of two / mixed radix (8-point blocks, radix 5 and 3 passes)
def test_three_distance_parity_medium(n_frames):
    from psa_b200 import SEDCalculator
    spec = synth.si_spec("mid", n_cells=4, n_frames=n_frames, seed=21)
    traj = spec.trajectory()
    calc = SEDCalculator(traj, *spec.cells)
    for direction in ([1, 0, 0], [1, 1, 0]):
        mags, kv = calc.get_k_path(direction, 4.0, 48)
        new = calc.calculate(mags, kv)
        ref = O.calculate(traj.positions, traj.velocities, traj.types, traj.dt_ps, kv)
        o64 = O.calculate_fp64(traj.positions, traj.velocities, traj.types, traj.dt_ps, kv)
        i_new, i_ref = O.intensity(new.sed).astype(np.float64), O.intensity(ref["sed"]).astype(np.float64)
        i_64 = np.sum(np.abs(o64["sed"]) ** 2, axis=-1)
        rep = O.parity_report(i_new, i_ref, i_64)
        print(direction, {t: {k: f"{v['max']:.2e}" for k, v in rep[t].items()} for t in (1e-4, 1e-5, 1e-6)})
        assert rep["global_peak_equal"] and rep["per_k_peak_equal"]
        # against the float64 truth the CUDA path is well inside the north-star tolerance ...
        assert rep[1e-6]["new_o64"]["max"] < 1e-5
        # ... and against the reference it is limited by the reference's own float32 floor
        floor = rep[1e-6]["ref_o64"]["max"]
        assert rep[1e-6]["new_ref"]["max"] < max(1e-5, 1.5 * floor)
        assert rep[1e-4]["new_ref"]["max"] < 1e-5
        assert rep[1e-6]["new_o64"]["max"] <= floor                      # never worse than the reference


# ------------------------------------------------------------------ BASELINE config 1 at full size: properties
def test_full_size_config1_properties():
    from psa_b200 import SEDCalculator, _lib
    from psa_b200.engine import Engine
    cfg = synth.baseline_config("c1")
    spec = cfg["spec"]
    traj = spec.trajectory()
    calc = SEDCalculator(traj, *spec.cells)
    mags, kv = calc.get_k_path([1, 0, 0], 4.0, 100)
    res = calc.calculate(mags, kv)
    assert res.sed.shape == (8192, 100, 3)
    # (a) the Gamma-point column is the FFT of the total velocity: an O(n_t n_a) host computation
    tot = traj.velocities.astype(np.float64).sum(axis=1)
    want0 = np.fft.fft(tot, axis=0) / tot.shape[0]
    assert np.abs(res.sed[:, 0, :] - want0).max() < 5e-7 * np.abs(want0).max()
    # (b) dispersion peak of the strongest planted mode sits where it was planted (frequency bin)
    inten = res.intensity
    f_pk, k_pk = np.unravel_index(np.argmax(inten[: 4096]), inten[: 4096].shape)
    df = 1.0 / (8192 * spec.dt_ps)
    assert np.min(np.abs(spec.freq - f_pk * df)) <= df
    # (c) tensor-core and CUDA-core projection agree bit for bit at full size (k subset)
    calc_s = SEDCalculator(traj, *spec.cells)
    calc_s._engine = Engine(project_impl=_lib.PROJECT_SIMT)
    sub = calc_s.calculate(mags[40:48], kv[40:48])
    np.testing.assert_array_equal(sub.sed, res.sed[:, 40:48, :])
    # (d) linearity in the atom selection: S(all) = S(type 1) + S(type 2) up to float32 rounding
    s1 = calc.calculate(mags[:16], kv[:16], basis_atom_types=[1]).sed.astype(np.complex128)
    s2 = calc.calculate(mags[:16], kv[:16], basis_atom_types=[2]).sed.astype(np.complex128)
    assert np.abs(s1 + s2 - res.sed[:, :16, :]).max() < 1e-6 * np.abs(res.sed[:, :16, :]).max()


# ------------------------------------------------------------------ multi-GPU (needs >= 2 visible devices)
def test_k_sharded_two_gpus_equals_single_gpu():
    import subprocess
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = Path(__file__).resolve().parent / "multigpu_check.py"
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    import re as _re
    m = _re.search(r"MULTIGPU_CHECK world=2 results=\[([^\]]*)\]", out.stdout)
    # 3 broadcast-ingest cases + sliced-ingest cases: {frame-sharded, k-sharded} x {velocity, displacement} x 3 calls x
    # {first call, second call on the cached state} on an even frame split, the same on a ragged split (fallback path)
    got = m.group(1).split(", ") if m else []
    assert len(got) == 39 and got == ["True"] * 39, out.stdout[-2000:]


def test_npy_cache_streams_to_device(gold_si, tmp_path):
    """N1: a memory-mapped .npy cache goes through the chunked pinned-staging upload and gives the same bits."""
    from psa_b200 import SEDCalculator, cache
    from psa_b200 import engine as E
    calc = _calc(gold_si)
    cache.save_npy_cache(calc.traj, tmp_path / "run.lammpstrj")
    mapped = cache.load_npy_cache(tmp_path / "run.lammpstrj", dt=float(gold_si["dt_ps"]))
    old = E._UPLOAD_CHUNK_BYTES
    E._UPLOAD_CHUNK_BYTES = 40_000            # force several staging chunks on this small trajectory
    try:
        calc_m = SEDCalculator(mapped, *[int(c) for c in gold_si["cells"]])
        kv = gold_si["kpath_110_vecs"]
        a = calc_m.calculate(np.zeros(len(kv)), kv).sed
    finally:
        E._UPLOAD_CHUNK_BYTES = old
    b = calc.calculate(np.zeros(len(kv)), kv).sed
    np.testing.assert_array_equal(a, b)


# ------------------------------------------------------------------ BASELINE configs at FULL size, k-subset vs the oracle
def _subset_parity(calc, traj, kv_sub, **kw):
    """new / O-ref / O-64 on a handful of k-points of a full-size trajectory (the oracle's cost is linear in k)."""
    new = calc.calculate(np.zeros(len(kv_sub), np.float32), kv_sub, **kw)
    ref = O.calculate(traj.positions, traj.velocities, traj.types, traj.dt_ps, kv_sub, **kw)
    o64 = O.calculate_fp64(traj.positions, traj.velocities, traj.types, traj.dt_ps, kv_sub, **kw)
    assert new.sed.shape == ref["sed"].shape and new.sed.dtype == ref["sed"].dtype
    assert new.is_complex == ref["is_complex"]
    if new.is_complex:
        i_new, i_ref = O.intensity(new.sed).astype(np.float64), O.intensity(ref["sed"]).astype(np.float64)
        i_64 = np.sum(np.abs(o64["sed"]) ** 2, axis=-1)
    else:
        i_new, i_ref, i_64 = new.sed.astype(np.float64), ref["sed"].astype(np.float64), o64["sed"]
    rep = O.parity_report(i_new, i_ref, i_64)
    print({t: {k: f"{v['max']:.2e}" for k, v in rep[t].items()} for t in (1e-4, 1e-5, 1e-6)})
    assert rep["global_peak_equal"] and rep["per_k_peak_equal"]
    assert rep[1e-6]["new_o64"]["max"] < 1e-5                      # north-star tolerance against the truth
    # ... and at the level of the reference's own rounding floor (both are a few 1e-6 at these sizes)
    assert rep[1e-6]["new_o64"]["max"] <= max(1.5 * rep[1e-6]["ref_o64"]["max"], 4e-6)
    assert rep[1e-6]["new_ref"]["max"] < max(1e-5, 1.3 * rep[1e-6]["ref_o64"]["max"])
    assert rep[1e-4]["new_ref"]["max"] < 1e-5
    return new, rep


@pytest.mark.parametrize("name", ["c1", "c2"])
def test_full_size_baseline_config_parity_on_k_subset(name):
    from psa_b200 import SEDCalculator
    cfg = synth.baseline_config(name)
    spec = cfg["spec"]
    traj = spec.trajectory()
    calc = SEDCalculator(traj, *spec.cells)
    path = cfg["paths"][0]
    mags, kv = calc.get_k_path(path["direction"], cfg["bz_coverage"], path["n_k"])
    # the dispersion-peak column, its neighbours and a few others
    full = calc.calculate(mags, kv, basis_atom_types=cfg["basis_atom_types"], summation_mode=cfg["summation_mode"])
    inten = full.intensity if full.is_complex else full.sed
    k_pk = int(np.unravel_index(np.argmax(inten), inten.shape)[1])
    pick = sorted({0, 1, max(k_pk - 1, 0), k_pk, min(k_pk + 1, len(kv) - 1), len(kv) // 2, len(kv) - 1})
    new, _ = _subset_parity(calc, traj, kv[pick], basis_atom_types=cfg["basis_atom_types"],
                            summation_mode=cfg["summation_mode"])
    # the subset call and the full call agree bit for bit (k columns are independent and exact)
    np.testing.assert_array_equal(new.sed, full.sed[:, pick])


def test_full_size_kgrid_config4_parity_on_k_subset():
    """C4 at full size (13 824 atoms x 16 384 frames): grid points incl. the dispersion peak of a coarse scan,
    against the oracle; the k-grid result layout (k_grid_shape, empty k_points) on the way."""
    from psa_b200 import SEDCalculator
    cfg = synth.baseline_config("c4")
    spec = cfg["spec"]
    traj = spec.trajectory()
    calc = SEDCalculator(traj, *spec.cells)
    kr = cfg["k_ranges"]
    mags, kv, shape = calc.get_k_grid(cfg["plane"], kr[:2], kr[2:], cfg["n_kx"], cfg["n_ky"], cfg["k_fixed"])
    assert shape == (100, 100) and mags.size == 0 and kv.shape == (10000, 3)
    coarse = np.arange(0, 10000, 97)                                     # 104 grid points across the plane
    scan = calc.calculate_intensity(mags, kv[coarse], k_grid_shape=None)
    k_pk = int(coarse[np.unravel_index(np.argmax(scan.sed), scan.sed.shape)[1]])
    pick = sorted({0, 50, k_pk, min(k_pk + 1, 9999), 5050, 9999})
    new, _ = _subset_parity(calc, traj, kv[pick], summation_mode="coherent")
    # the same columns out of a streamed multi-chunk call (ragged chunks of 5) are the same bits
    again = calc.calculate(np.zeros(len(pick), np.float32), kv[pick], k_chunk_size=5)
    np.testing.assert_array_equal(again.sed, new.sed)


@pytest.mark.slow
@pytest.mark.skipif(not os.environ.get("PSA_TEST_C5"), reason="50 GB of host trajectory + minutes of CPU: set PSA_TEST_C5=1")
def test_full_size_config5_parity_on_k_subset():
    """C5 at full size (64 000 atoms x 32 768 frames): the two-pass (> 32 768 atoms) float32 accumulation of the
    projection and the 32 768-point transform in the whole path, three k-points against the float64 truth fed the
    reference's float32 mean positions and complex64 phase table (O-64, evaluated frame block by frame block so that
    it fits in memory; the float32 reference arithmetic itself needs ~3 copies of the 25 GB series and is not run)."""
    from psa_b200 import SEDCalculator
    cfg = synth.baseline_config("c5")
    spec = cfg["spec"]
    traj = spec.trajectory(threads=16)
    calc = SEDCalculator(traj, *spec.cells)
    mags, kv = calc.get_k_path([1, 0, 0], cfg["bz_coverage"], 256)
    pick = [0, 64, 255]
    new = calc.calculate(mags[pick], kv[pick])
    assert new.sed.shape == (32768, 3, 3) and new.is_complex
    mean = O.mean_positions(traj.positions)
    np.testing.assert_array_equal(calc.device_trajectory.mean.cpu().numpy(), mean)
    ph = O.phase_table(kv[pick], mean).astype(np.complex128)                 # (3, n_a), float32 inputs widened
    proj = np.empty((32768, 3, 3), np.complex128)
    for t0 in range(0, 32768, 512):
        blk = traj.velocities[t0:t0 + 512].astype(np.float64)                # (512, n_a, 3)
        proj[t0:t0 + 512] = np.einsum("tap,ka->tkp", blk, ph, optimize=True)
    want = np.fft.fft(proj, axis=0) / 32768
    i_new, i_64 = O.intensity(new.sed).astype(np.float64), np.sum(np.abs(want) ** 2, axis=-1)
    for thr in (1e-4, 1e-5, 1e-6):
        r = O.rel_err_above(i_new, i_64, thr)
        print(f"C5 new<->fp64 at I > {thr:g} peak: max {r['max']:.2e} p99 {r['p99']:.2e} (n={r['n']})")
    assert O.rel_err_above(i_new, i_64, 1e-6)["max"] < 1e-5
    assert O.peak_indices(i_new)[0] == O.peak_indices(i_64)[0]
    assert np.array_equal(O.peak_indices(i_new)[1], O.peak_indices(i_64)[1])


def test_full_size_graphene_chirality():
    """C3 at full size: parity on a k-subset, and the handedness of the planted circular modes."""
    from psa_b200 import SEDCalculator
    cfg = synth.baseline_config("c3")
    spec = cfg["spec"]
    traj = spec.trajectory()
    calc = SEDCalculator(traj, *spec.cells)
    res = calc.calculate_chiral_sed([1, 0, 0], cfg["bz_coverage"], 80, chiral_axis="z")
    assert res.phase.shape == res.sed.shape[:2] and np.abs(res.phase).max() <= np.pi / 2 + 1e-6
    _subset_parity(calc, traj, res.k_vectors[[0, 7, 20, 41, 79]], summation_mode="coherent")
    # A planted mode u = A Re[(x + i h y)/sqrt2 e^{i(q.r - w t)}] shows up at (k, f) = (+q, +w) through its
    # conjugate term, S_y / S_x = -i h, and at (-q, -w) with S_y / S_x = +i h: the folded phase
    # arg(S_x) - arg(S_y) must be +h pi/2 resp. -h pi/2 there.  Modes whose q lies on the sampled [100] path
    # are checked at the nearest k sample, at the exact frequency bin.
    df = 1.0 / (spec.n_frames * spec.dt_ps)
    on_bin = [(q, f, np.sign(h)) for q, f, h in zip(spec.q, spec.freq, spec.pol_im[:, 1])
              if abs(q[1]) < 1e-9 and abs(q[2]) < 1e-9 and q[0] > 0 and abs(f / df - round(f / df)) < 1e-6]
    assert on_bin, "the synthetic graphene spec must plant at least one on-bin mode along [100]"
    kq = np.array([[q[0], 0.0, 0.0] for q, _, _ in on_bin], np.float32)          # sample exactly at the planted q
    at_q = calc.calculate(np.zeros(len(kq), np.float32), kq)
    phase = calc.calculate_chiral_phase(at_q.sed[:, :, 0], at_q.sed[:, :, 1], "C")
    inten = O.intensity(at_q.sed)
    for col, (q, f, h) in enumerate(on_bin):
        f_bin = int(round(f / df))
        assert inten[f_bin, col] > 0.2 * inten[:, col].max()                         # the planted line is there
        assert abs(phase[f_bin, col] - h * np.pi / 2) < 0.1, (q, f, h, phase[f_bin, col])
