"""Multi-GPU SED: k-point sharding over the GPUs of one box (one process per GPU).

The SED path shards naturally by k-point - every output column depends on the whole trajectory
but on no other k (reference: src/psa/core/sed_calculator.py:287-327 already loops over independent
k-chunks).  So there is exactly one exchange step: the k-independent device state (float32 mean
positions and the int8 digit planes of the projected series) is produced once on the source rank
and broadcast with NCCL over NVLink; every rank then projects + transforms its own contiguous
k-slice with no further communication, and the slices are gathered on the destination rank.

``torch.distributed`` is plumbing here (process group, broadcast, gather); the arithmetic is the
same C-ABI path as on one GPU.  The host-side pieces (slice arithmetic, gather re-assembly) work
on any backend and are covered by world_size-2 ``gloo`` tests on CPU.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of ``range(n)``: the first ``n % world`` ranks get one extra item."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """``(rank, world, local_rank)`` from torchrun's environment; initialises the process group when
    WORLD_SIZE > 1 (NCCL when CUDA is available, gloo otherwise)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if torch.cuda.is_available():
        torch.cuda.set_device(local % torch.cuda.device_count())
    if world > 1 and not dist.is_initialized():
        # NCCL writes its version/debug lines to stdout by default; callers print machine-readable
        # results there, so send NCCL's own chatter to stderr unless the user chose a file
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", torch.cuda.current_device())
        dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local


def broadcast_tensors(tensors: Sequence[Optional[torch.Tensor]], metas: Optional[List[Tuple]], src: int,
                      device: torch.device, group=None) -> List[torch.Tensor]:
    """Broadcast a list of tensors whose shapes/dtypes only ``src`` knows.  Returns the tensors on every rank."""
    rank = dist.get_rank(group)
    box = [metas if rank == src else None]
    dist.broadcast_object_list(box, src=src, group=group)
    out: List[torch.Tensor] = []
    for i, (shape, dtype) in enumerate(box[0]):
        t = tensors[i] if rank == src else torch.empty(shape, dtype=dtype, device=device)
        dist.broadcast(t, src=src, group=group)
        out.append(t)
    return out


def gather_k_slices(local: torch.Tensor, n_k_total: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Assemble per-rank results ``(n_f, n_k_local, ...)`` (contiguous k-slices in rank order, as produced
    with :func:`shard_range`) into ``(n_f, n_k_total, ...)`` on ``dst``; other ranks get ``None``."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    widths = [b - a for a, b in (shard_range(n_k_total, r, world) for r in range(world))]
    w_max = max(widths)
    shape = list(local.shape)
    padded = local
    if shape[1] != w_max:
        shape[1] = w_max
        padded = torch.zeros(shape, dtype=local.dtype, device=local.device)
        padded[:, :local.shape[1]] = local
    padded = padded.contiguous()
    if padded.is_complex():
        padded = torch.view_as_real(padded)
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    full_shape = list(local.shape)
    full_shape[1] = n_k_total
    full = torch.empty(full_shape, dtype=local.dtype, device=local.device)
    for r, buf in enumerate(bufs):
        a, b = shard_range(n_k_total, r, world)
        part = torch.view_as_complex(buf) if local.is_complex() else buf
        full[:, a:b] = part[:, : b - a]
    return full


def sliced_ingest(calc, proj_groups, local_rows=None, group=None) -> None:
    """k-independent state from a trajectory whose FRAMES are spread over the ranks.

    Rank r uploads only frames ``shard_range(n_t, r, world)`` - 1/N of the bytes over its own PCIe link -
    and the ranks then build the shared state together:

    * mean positions: the float32 sum of a column must run in frame order to stay bit-identical with
      NumPy, so the running sums travel down the ranks (one (n_atoms, 3) message per hop); the last rank
      divides and broadcasts the mean;
    * digit planes: every frame row is digitised independently (its own exponent), each rank handles its
      rows and the row blocks are exchanged with one all-gather (or broadcast) per plane.

    ``local_rows = (positions[t0:t1], velocities[t0:t1])`` when this process holds nothing but its range;
    otherwise the range is sliced out of ``calc.traj``.  On return every rank's device trajectory has the
    mean and the digit planes of ``proj_groups`` installed, exactly as after the broadcast ingest.
    """
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    eng, dtraj = calc.engine, calc.device_trajectory
    # skip only if EVERY rank already holds the state (a rank that computed something on its own must not
    # leave the others waiting in the chain)
    ready = torch.tensor([1 if dtraj.has_state(proj_groups, calc.use_displacements) else 0], device=eng.device)
    dist.all_reduce(ready, op=dist.ReduceOp.MIN, group=group)
    if int(ready.item()) == 1:
        return
    n_t, n_a = dtraj.n_t, dtraj.n_a
    bounds = [shard_range(n_t, r, world) for r in range(world)]
    t0, t1 = bounds[rank]
    disp = calc.use_displacements
    pos_rows = dtraj.upload_rows("pos", t0, t1, None if local_rows is None else local_rows[0])
    data_rows = pos_rows if disp else dtraj.upload_rows("vel", t0, t1, None if local_rows is None else local_rows[1])

    acc = chain_running_sum(torch.zeros((n_a, 3), dtype=torch.float32, device=eng.device),
                            lambda a, last: eng.mean_accumulate(pos_rows, a, n_t if last else 0), group)
    dtraj.install_mean(acc)

    for g in proj_groups:
        idx, idx_dev, n_sel = dtraj.selection(g, disp)
        pitch = int(eng_pitch(n_sel))
        dig = eng.empty((3, 4, n_t, pitch), torch.int8)
        expo = eng.empty((3, n_t), torch.int32)
        eng.digitize_rows(data_rows, acc if disp else None, idx_dev, n_sel, pitch, dig, expo, n_t, t0)
        exchange_row_blocks(list(dig.view(12, n_t, pitch).unbind(0)) + list(expo.unbind(0)), bounds, group)
        dtraj.install_group(idx, disp, dig, expo)


def chain_running_sum(acc: torch.Tensor, accumulate, group=None) -> torch.Tensor:
    """Ordered reduction over the ranks: rank 0 starts from ``acc``, every rank continues the running value
    with ``accumulate(acc, is_last_rank)`` (in place) and hands it to the next one; the last rank's result
    is broadcast to all.  The order of operations is that of a single process walking the ranks' data in
    rank order - which is what keeps a float32 sum bit-identical."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if rank > 0:
        dist.recv(acc, src=rank - 1, group=group)
    accumulate(acc, rank == world - 1)
    if rank < world - 1:
        dist.send(acc, dst=rank + 1, group=group)
    dist.broadcast(acc, src=world - 1, group=group)
    return acc


def exchange_row_blocks(planes, bounds, group=None) -> None:
    """Every tensor in ``planes`` has its leading axis split over the ranks as ``bounds[r] = (a, b)``; each
    rank has filled its own block.  Afterwards every rank holds every block (in place): one all-gather
    per plane when the blocks have equal length, one broadcast per (plane, rank) otherwise."""
    rank = dist.get_rank(group)
    t0, t1 = bounds[rank]
    equal = all(b - a == t1 - t0 for a, b in bounds)
    for plane in planes:
        if equal:
            dist.all_gather_into_tensor(plane, plane[t0:t1], group=group)
        else:
            for r, (a, b) in enumerate(bounds):
                if b > a:
                    dist.broadcast(plane[a:b], src=r, group=group)


def eng_pitch(n_sel: int) -> int:
    from . import _lib
    return int(_lib.load().psa_pitch(n_sel))


def calculate_sharded(calc, k_points_mags: np.ndarray, k_vectors_3d: np.ndarray, basis_atom_indices=None,
                      basis_atom_types=None, summation_mode: str = "coherent", k_grid_shape=None, src: int = 0,
                      group=None, ingest: str = "broadcast", local_rows=None):
    """``SEDCalculator.calculate`` over all ranks of the process group.

    Every rank calls this with a calculator built on a trajectory of the right *shape*.  With
    ``ingest="broadcast"`` only ``src`` needs real positions/velocities (the others may hold zero-stride
    placeholders): it uploads and ingests everything and broadcasts the result.  With ``ingest="sliced"``
    every rank holds (at least) its own range of frames and uploads just that, see :func:`sliced_ingest`.
    Returns the ``SED`` on ``src`` and ``None`` elsewhere.
    """
    from . import groups as grp
    from .engine import sed_on_device
    from .sed import SED

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return calc.calculate(k_points_mags, k_vectors_3d, basis_atom_indices, basis_atom_types,
                              summation_mode, k_grid_shape)
    if summation_mode not in ("coherent", "incoherent"):
        raise ValueError(f"summation_mode must be 'coherent' or 'incoherent', got {summation_mode}")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    traj = calc.traj
    eng, dtraj = calc.engine, calc.device_trajectory
    groups = grp.resolve_sed_groups(traj.types, traj.n_atoms, basis_atom_indices, basis_atom_types, summation_mode)
    complex_out, proj_groups = grp.plan_sed_groups(groups, summation_mode)

    if ingest not in ("broadcast", "sliced"):
        raise ValueError(f"ingest must be 'broadcast' or 'sliced', got {ingest}")
    with torch.cuda.device(eng.device):
        if ingest == "sliced":
            sliced_ingest(calc, proj_groups, local_rows, group)
        # 1. k-independent state: built on src, broadcast once
        tensors: List[Optional[torch.Tensor]] = []
        metas: List[Tuple] = []
        if ingest == "sliced":
            got = []
        elif rank == src:
            tensors.append(dtraj.mean)
            for g in proj_groups:
                _, _, _, dig, expo = dtraj.group(g, calc.use_displacements)
                tensors += [dig, expo]
            metas = [(tuple(t.shape), t.dtype) for t in tensors]
        if ingest == "broadcast":
            got = broadcast_tensors(tensors, metas, src, eng.device, group)
        if ingest == "broadcast" and rank != src:
            dtraj.install_mean(got[0])
            for i, g in enumerate(proj_groups):
                dtraj.install_group(g, calc.use_displacements, got[1 + 2 * i], got[2 + 2 * i])

        # 2. every rank: its contiguous k-slice, no communication
        k_vecs = np.ascontiguousarray(np.asarray(k_vectors_3d, dtype=np.float32).reshape(-1, 3))
        n_k = k_vecs.shape[0]
        k0, k1 = shard_range(n_k, rank, world)
        local = sed_on_device(dtraj, k_vecs[k0:k1], proj_groups, complex_out, calc.use_displacements)

        # 3. gather on src
        full = gather_k_slices(local, n_k, dst=src, group=group)
        if rank != src:
            return None
        sed_host = calc._to_host(full)
    freqs = np.fft.fftfreq(traj.n_frames, d=calc.dt_ps)
    return SED(sed_host, freqs, k_points_mags, k_vectors_3d, k_grid_shape=k_grid_shape,
               is_complex=complex_out, phase=None, context=calc._context(groups))
