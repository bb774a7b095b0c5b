"""Run under torchrun (one rank per GPU): the k-sharded multi-GPU SED must equal the single-GPU
result bit for bit (every k column is independent and the projection is exact).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/multigpu_check.py
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from psa_b200 import SEDCalculator  # noqa: E402
import synthetic as synth  # noqa: E402
from psa_b200 import dist as pdist  # noqa: E402


def main():
    rank, world, local = pdist.init_from_env()
    spec = synth.si_spec("mg", n_cells=3, n_frames=1024, seed=31)
    traj = spec.trajectory(threads=1)
    if rank != 0:   # only rank 0 owns real data; the others get zero-stride placeholders of the same shape
        zero = np.broadcast_to(np.zeros(1, np.float32), traj.positions.shape)
        traj = spec.wrap(zero, zero)
    calc = SEDCalculator(traj, *spec.cells)
    mags, kv = calc.get_k_path([1, 1, 0], 4.0, 37)
    checks = []
    for kw in (dict(summation_mode="coherent"),
               dict(basis_atom_types=[1, 2], summation_mode="incoherent"),
               dict(basis_atom_indices=[3, 5, 8, 13, 21, 34], summation_mode="coherent")):
        res = pdist.calculate_sharded(calc, mags, kv, **kw)
        if rank == 0:
            single = SEDCalculator(spec.trajectory(threads=1), *spec.cells).calculate(mags, kv, **kw)
            checks.append(bool(np.array_equal(res.sed, single.sed)) and res.is_complex == single.is_complex)
        else:
            assert res is None
    # sliced ingest: every rank uploads 1/N of the frames (here: handed over as its only data), the running
    # mean sums travel down the ranks and the digit planes are all-gathered - still the same bits.
    # 1000 frames over 3 ranks would be ragged; 1024 over 2/4/8 is even: cover both with a ragged 1000-frame run.
    full = spec.trajectory(threads=1)
    for n_t in (1024, 1000):
        t0, t1 = pdist.shard_range(n_t, rank, world)
        shape_only = np.broadcast_to(np.zeros(1, np.float32), (n_t,) + full.positions.shape[1:])
        from psa_b200 import Trajectory
        ph = Trajectory(shape_only, shape_only, full.types, np.arange(n_t), full.box_matrix, full.box_lengths,
                        full.box_tilts, full.dt_ps)
        for use_disp in (False, True):
            calc_s = SEDCalculator(ph, *spec.cells, use_displacements=use_disp)
            rows = (np.ascontiguousarray(full.positions[t0:t1]), np.ascontiguousarray(full.velocities[t0:t1]))
            for kw in (dict(summation_mode="coherent"), dict(basis_atom_types=[1, 2], summation_mode="incoherent")):
                res = pdist.calculate_sharded(calc_s, mags, kv, ingest="sliced", local_rows=rows, **kw)
                if rank == 0:
                    whole = Trajectory(full.positions[:n_t], full.velocities[:n_t], full.types, np.arange(n_t),
                                       full.box_matrix, full.box_lengths, full.box_tilts, full.dt_ps)
                    single = SEDCalculator(whole, *spec.cells, use_displacements=use_disp).calculate(mags, kv, **kw)
                    checks.append(bool(np.array_equal(res.sed, single.sed)))
    if rank == 0:
        print(f"MULTIGPU_CHECK world={world} results={checks}", flush=True)
        assert all(checks)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
