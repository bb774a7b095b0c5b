"""Host-side logic of the k-sharded multi-GPU path, on CPU with the gloo backend (world_size 2 and 3)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from psa_b200 import dist as pdist


@pytest.mark.parametrize("n,world", [(200, 1), (200, 2), (768, 8), (10, 4), (3, 8), (0, 2), (10000, 8)])
def test_shard_range_partitions_exactly(n, world):
    ranges = [pdist.shard_range(n, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
        assert a1 == b0 and a1 >= a0
    widths = [b - a for a, b in ranges]
    assert max(widths) - min(widths) <= 1 and sum(widths) == n


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_k, result_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the "true" full result every rank can reconstruct deterministically
        g = torch.Generator().manual_seed(1234)
        full_c = torch.view_as_complex(torch.randn(16, n_k, 3, 2, generator=g))
        full_r = torch.randn(16, n_k, generator=g)
        k0, k1 = pdist.shard_range(n_k, rank, world)
        got_c = pdist.gather_k_slices(full_c[:, k0:k1].contiguous(), n_k, dst=0)
        got_r = pdist.gather_k_slices(full_r[:, k0:k1].contiguous(), n_k, dst=0)
        # k-independent state known only to rank 0 (shapes included)
        tensors, metas = [], None
        if rank == 0:
            tensors = [torch.arange(12, dtype=torch.float32).reshape(4, 3),
                       torch.arange(3 * 4 * 5 * 64, dtype=torch.int32).reshape(3, 4, 5, 64).to(torch.int8),
                       torch.arange(15, dtype=torch.int32).reshape(3, 5)]
            metas = [(tuple(t.shape), t.dtype) for t in tensors]
        state = pdist.broadcast_tensors(tensors, metas, 0, torch.device("cpu"))
        ok_state = (state[0].shape == (4, 3) and state[1].dtype == torch.int8 and state[1].shape == (3, 4, 5, 64)
                    and int(state[2].sum()) == sum(range(15)) and float(state[0][3, 2]) == 11.0)
        if rank == 0:
            result_q.put((bool(torch.equal(got_c, full_c)), bool(torch.equal(got_r, full_r)), ok_state))
        else:
            assert got_c is None and got_r is None
            result_q.put((True, True, ok_state))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_k", [(2, 200), (2, 7), (3, 10)])
def test_gather_and_broadcast_over_gloo(world, n_k):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_k, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(all(r) for r in results), results
