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


@pytest.mark.parametrize("n_k,world,cap", [(10000, 8, 2048), (10000, 2, 2048), (1000, 8, 2048), (37, 2, 3), (5, 8, 2048),
                                           (200, 4, 64)])
def test_frame_shard_plan_covers_every_k_once(n_k, world, cap):
    """Frame-sharded multi-GPU plan: the owners' slices are the k-sharded layout, their chunks tile them exactly, no
    chunk exceeds the cap, and at every (chunk, step) the ranks write to pairwise different owners (one stream per
    NVLink port: rank r works for owner r + s at step s)."""
    slices, n_chunks, chunk = pdist.frame_shard_plan(n_k, world, cap)
    assert slices == [pdist.shard_range(n_k, q, world) for q in range(world)]
    seen = np.zeros(n_k, np.int64)
    for q in range(world):
        pieces = [chunk(q, j) for j in range(n_chunks)]
        assert pieces[0][0] == slices[q][0] and pieces[-1][1] == slices[q][1]
        for (a, b), (c, d) in zip(pieces[:-1], pieces[1:]):
            assert b == c
        for a, b in pieces:
            assert 0 <= b - a <= cap
            seen[a:b] += 1
    assert np.all(seen == 1)
    for s_ in range(world):
        assert sorted((r + s_) % world for r in range(world)) == list(range(world))
    # routed launches (one per chunk for all owners): the launch order lists every k once, and the row table hands
    # owner q exactly the rows of its piece, two per k-point
    every = []
    for j in range(n_chunks):
        order, row_begin = pdist.routed_chunk_rows(chunk, world, j)
        assert len(row_begin) == world + 1 and row_begin[0] == 0 and row_begin[-1] == 2 * order.size
        for q in range(world):
            a, b = chunk(q, j)
            assert row_begin[q + 1] - row_begin[q] == 2 * (b - a)
            assert np.array_equal(order[row_begin[q] // 2:row_begin[q + 1] // 2], np.arange(a, b))
        every.append(order)
    assert np.array_equal(np.sort(np.concatenate(every)), np.arange(n_k))


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


# ------------------------------------------------------------------ sliced ingest: chain + block exchange (host logic)
def _seq_sum_f32(acc: np.ndarray, rows: np.ndarray) -> np.ndarray:
    """acc (+) rows[0] (+) rows[1] ... in float32, in order - the stand-in for psa_mean_accumulate."""
    if rows.shape[0] == 0:
        return acc
    return np.cumsum(np.concatenate([acc[None], rows], axis=0), axis=0, dtype=np.float32)[-1]


def _sliced_worker(rank, world, port, n_t, result_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        pos = (rng.random((n_t, 5, 3)) * 30 + rng.standard_normal((n_t, 5, 3)) * 0.01).astype(np.float32)
        bounds = [pdist.shard_range(n_t, r, world) for r in range(world)]
        t0, t1 = bounds[rank]

        def accumulate(acc, last):
            out = _seq_sum_f32(acc.numpy().copy(), pos[t0:t1])
            if last:
                out = out / np.float32(n_t)
            acc.copy_(torch.from_numpy(out))

        mean = pdist.chain_running_sum(torch.zeros((5, 3), dtype=torch.float32), accumulate)
        ok_mean = bool(np.array_equal(mean.numpy(), np.mean(pos, axis=0, dtype=np.float32)))
        # row blocks: every rank fills its own rows of two "planes", afterwards all ranks hold everything
        want = [torch.arange(n_t * 4, dtype=torch.int32).reshape(n_t, 4), torch.arange(n_t, dtype=torch.int32) * 3]
        planes = [torch.full_like(w, -1) for w in want]
        for p, w in zip(planes, want):
            p[t0:t1] = w[t0:t1]
        pdist.exchange_row_blocks(planes, bounds)
        ok_planes = all(bool(torch.equal(p, w)) for p, w in zip(planes, want))
        result_q.put((ok_mean, ok_planes))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_t", [(2, 64), (2, 7), (3, 100), (3, 2)])
def test_sliced_ingest_chain_and_exchange_over_gloo(world, n_t):
    """The ordered float32 running sum handed from rank to rank equals np.mean(dtype=float32) bit for bit, and
    the row-block exchange (all-gather for equal blocks, broadcasts for ragged or empty ones) fills every rank."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sliced_worker, args=(r, world, port, n_t, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(all(r) for r in results), results


# ------------------------------------------------------------------ shared host result (no funnel through one rank)
def _shared_worker(rank, world, port, result_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_f, n_k = 6, 11
        shared = pdist.SharedHostArray((n_f, n_k, 3), np.complex64, src=0, register=False)
        k0, k1 = pdist.shard_range(n_k, rank, world)
        full = (np.arange(n_f * n_k * 3).reshape(n_f, n_k, 3) * (1 + 2j)).astype(np.complex64)
        shared.array[:, k0:k1] = full[:, k0:k1]                       # every rank fills its own column slice
        dist.barrier()
        ok = bool(np.array_equal(shared.array, full)) if rank == 0 else True
        again = pdist.SharedHostArray((4,), np.float32, src=0, register=False)   # a second array, different shape
        again.array[rank::world] = rank + 1
        dist.barrier()
        ok = ok and bool(np.array_equal(again.array, [(i % world) + 1 for i in range(4)]))
        dist.barrier()
        shared.close()
        again.close()
        result_q.put((ok,))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_shared_host_array_over_gloo(world):
    """Every rank maps the same POSIX shared-memory array and writes its k-slice; the source rank sees all of it."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shared_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(all(r) for r in results), results


# ------------------------------------------------------------------ frame-sharded SED: the whole data flow, CPU stand-ins
def _frame_sharded_worker(rank, world, port, n_t, n_k, cap, result_q):
    """What dist.frame_sharded_sed does, with NumPy standing in for the kernels and a SharedHostArray per owner for
    its IPC-mapped projection buffer: ordered mean chain, every rank projects ITS frames for all k-points in the
    routed launch order, rows [row_begin[q], row_begin[q + 1]) go into owner q's buffer at this rank's frame offset,
    a barrier plays the fence, owners transform their slice."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import psa_oracle as O
        rng = np.random.default_rng(11)
        n_a = 9
        pos = (rng.random((n_t, n_a, 3)) * 12 + rng.standard_normal((n_t, n_a, 3)) * 0.02).astype(np.float32)
        vel = rng.standard_normal((n_t, n_a, 3)).astype(np.float32)
        kv = (rng.standard_normal((n_k, 3)) * 0.7).astype(np.float32)
        bounds = [pdist.shard_range(n_t, r, world) for r in range(world)]
        f0, f1 = bounds[rank]

        def accumulate(acc, last):
            out = _seq_sum_f32(acc.numpy().copy(), pos[f0:f1])
            acc.copy_(torch.from_numpy(out / np.float32(n_t) if last else out))

        mean = pdist.chain_running_sum(torch.zeros((n_a, 3), dtype=torch.float32), accumulate).numpy()
        slices, n_chunks, chunk = pdist.frame_shard_plan(n_k, world, cap)
        kc = max(chunk(q, j)[1] - chunk(q, j)[0] for q in range(world) for j in range(n_chunks))
        k0, k1 = slices[rank]
        out = np.zeros((n_t, k1 - k0, 3), np.complex128)
        # one projection buffer per owner: [2 kc rows][3][n_t], row 2 i = Re, 2 i + 1 = Im of the owner's i-th k-point
        bufs = [pdist.SharedHostArray((2 * kc, 3, n_t), np.float64, src=q, register=False) for q in range(world)]
        for j in range(n_chunks):
            order, row_begin = pdist.routed_chunk_rows(chunk, world, j)
            if order.size:
                ph = O.phase_table(kv[order], mean).astype(np.complex128)                  # (n_all, n_a)
                proj = np.einsum("tap,ka->kpt", vel[f0:f1].astype(np.float64), ph)         # this rank's frames, all k
                rows = np.empty((2 * order.size, 3, f1 - f0))
                rows[0::2], rows[1::2] = proj.real, proj.imag
                for q in range(world):
                    a, b = row_begin[q], row_begin[q + 1]
                    bufs[q].array[:b - a, :, f0:f1] = rows[a:b]                            # the epilogue's peer stores
            dist.barrier()                                                                 # the fence of chunk j
            ka, kb = chunk(rank, j)
            if kb > ka:
                mine = bufs[rank].array[:2 * (kb - ka)]
                z = mine[0::2] + 1j * mine[1::2]                                           # (nk, 3, n_t)
                out[:, ka - k0:kb - k0] = (np.fft.fft(z, axis=2) / n_t).transpose(2, 0, 1)
            dist.barrier()                                                                 # buffers free for chunk j + 1
        ph_all = O.phase_table(kv[k0:k1], mean).astype(np.complex128)
        want = np.fft.fft(np.einsum("tap,ka->tkp", vel.astype(np.float64), ph_all), axis=0) / n_t
        ok_mean = bool(np.array_equal(mean, np.mean(pos, axis=0, dtype=np.float32)))
        scale = max(float(np.abs(want).max()), 1e-30) if want.size else 1.0
        ok_sed = bool(np.abs(out - want).max() <= 1e-12 * scale) if want.size else True
        for b in bufs:
            b.close()
        result_q.put((ok_mean, ok_sed))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_t,n_k,cap", [(2, 64, 21, 2048), (2, 40, 10, 3), (3, 96, 7, 2), (3, 12, 2, 2048)])
def test_frame_sharded_data_flow_over_gloo(world, n_t, n_k, cap):
    """Frames spread over the ranks, k-points owned by the ranks: plan, routed row table, frame offsets and owner layout
    of dist.frame_sharded_sed reproduce the one-process result (float64 stand-in arithmetic; several chunks per owner,
    ragged slices, an owner without k-points)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_frame_sharded_worker, args=(r, world, port, n_t, n_k, cap, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(all(r) for r in results), results
