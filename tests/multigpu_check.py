"""Run under torchrun (one rank per GPU): the k-sharded multi-GPU SED must equal the single-GPU
result bit for bit (every k column is independent and the projection is exact).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/multigpu_check.py
"""
import os
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
    # 1024 frames over 2/4/8 ranks is even; 1002 gives frame ranges that are not multiples of 4 (the frame-sharded
    # path then hands over to the k-sharded one, whose exchange falls back from the pipelined ring as well).
    full = spec.trajectory(threads=1)
    for n_t in (1024, 1002):
        t0, t1 = pdist.shard_range(n_t, rank, world)
        shape_only = np.broadcast_to(np.zeros(1, np.float32), (n_t,) + full.positions.shape[1:])
        from psa_b200 import Trajectory
        ph = Trajectory(shape_only, shape_only, full.types, np.arange(n_t), full.box_matrix, full.box_lengths,
                        full.box_tilts, full.dt_ps)
        even = all(pdist.shard_range(n_t, r, world)[0] % 4 == 0 for r in range(world))
        whole = Trajectory(full.positions[:n_t], full.velocities[:n_t], full.types, np.arange(n_t),
                           full.box_matrix, full.box_lengths, full.box_tilts, full.dt_ps)
        # "frames" (default): nothing but projections crosses NVLink (needs frame ranges that are multiples of 4, else
        # it falls back by itself); "k": digit planes all-gathered by the digitise kernel, k-sharded compute
        for shard in (("frames", "k") if even else ("frames",)):
            os.environ["PSA_B200_SHARD"] = shard
            for use_disp in (False, True):
                calc_s = SEDCalculator(ph, *spec.cells, use_displacements=use_disp)
                rows = (np.ascontiguousarray(full.positions[t0:t1]), np.ascontiguousarray(full.velocities[t0:t1]))
                for kw in (dict(summation_mode="coherent"), dict(basis_atom_types=[1, 2], summation_mode="incoherent"),
                           dict(summation_mode="coherent", k_chunk_size=3)):          # several chunks per owner
                    res = pdist.calculate_sharded(calc_s, mags, kv, ingest="sliced", local_rows=rows, **kw)
                    want_path = "frames" if (shard == "frames" and even) else "k"
                    if rank == 0:
                        single = SEDCalculator(whole, *spec.cells, use_displacements=use_disp).calculate(mags, kv, **kw)
                        checks.append(bool(np.array_equal(res.sed, single.sed)) and pdist.last_path == want_path)
                    # a second call on the cached state (other k-points), still the same bits
                    res2 = pdist.calculate_sharded(calc_s, mags[::2], kv[::2], ingest="sliced", local_rows=rows, **kw)
                    if rank == 0:
                        checks.append(bool(np.array_equal(res2.sed, single.sed[:, ::2])))
        os.environ.pop("PSA_B200_SHARD", None)
    if rank == 0:
        print(f"MULTIGPU_CHECK world={world} results={checks}", flush=True)
        assert all(checks)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
