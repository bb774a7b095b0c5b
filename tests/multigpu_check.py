"""Run under torchrun (one rank per GPU): the k-sharded multi-GPU SED must equal the single-GPU
result bit for bit (every k column is independent and the projection is exact).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/multigpu_check.py
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from psa_b200 import SEDCalculator, synth  # noqa: E402
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
    if rank == 0:
        print(f"MULTIGPU_CHECK world={world} results={checks}", flush=True)
        assert all(checks)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
