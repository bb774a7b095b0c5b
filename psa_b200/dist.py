"""Multi-GPU SED: k-point sharding over the GPUs of one box (one process per GPU).

The SED path shards naturally by k-point - every output column depends on the whole trajectory
but on no other k (reference: src/psa/core/sed_calculator.py:287-327 already loops over independent
k-chunks).  So there is exactly one exchange step: the k-independent device state (float32 mean
positions and the int8 digit planes of the projected series) has to reach every rank.  Two ways:

* ``ingest="broadcast"``: the source rank uploads and ingests everything, one NCCL broadcast.
* ``ingest="sliced"``: every rank uploads 1/N of the frames over its own PCIe link.  The float32 mean is an
  ordered running sum handed from rank to rank; the digit planes are produced by ONE kernel per rank that stores
  every digitised row into all ranks' planes through NVLink (CUDA-IPC mapped peer memory,
  ``psa_digitize_rows_peers``) - the all-gather is fused into the producing kernel.  When peer mapping is not
  available the rows are exchanged with NCCL all-gathers instead (same bits).

Every rank then projects + transforms its own contiguous k-slice with no further communication and copies its
spectra straight into its column slice of ONE result array in shared, page-locked host memory
(:class:`SharedHostArray`) - every PCIe link carries 1/N of the result, nothing funnels through the source rank.

``torch.distributed`` is plumbing here (process group, small collectives); the arithmetic is the same C-ABI
path as on one GPU.  The host-side pieces (slice arithmetic, the ordered chain, shared result array) work on
any backend and are covered by world_size-2 ``gloo`` tests on CPU.
"""
from __future__ import annotations

import ctypes
import os
from multiprocessing import resource_tracker, shared_memory
from typing import Dict, List, Optional, Sequence, Tuple

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


# ---------------------------------------------------------------------------------------------- shared host result
class SharedHostArray:
    """One host array mapped by every rank of the box (POSIX shared memory), page-locked for CUDA in each process.

    Collective: every rank of ``group`` constructs it with the same shape/dtype.  ``array`` is the NumPy view,
    ``ptr`` its address.  The segment is unlinked as soon as every rank has mapped it, so nothing is left behind
    in /dev/shm if a process dies; the memory lives until the last mapping closes."""

    def __init__(self, shape: Sequence[int], dtype, src: int = 0, group=None, register: Optional[bool] = None):
        self.shape, self.dtype = tuple(int(s) for s in shape), np.dtype(dtype)
        nbytes = max(1, int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize)
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        rank = dist.get_rank(group) if distributed else src
        self._shm, self.array, self.ptr, self.nbytes, self._registered = None, None, 0, nbytes, False
        if rank == src:
            try:
                self._shm = shared_memory.SharedMemory(create=True, size=nbytes)
            except OSError as exc:                       # /dev/shm too small (containers often cap it at 64 MB)
                self.error = f"cannot create {nbytes} bytes of POSIX shared memory: {exc}"
        if distributed:
            box = [self._shm.name if self._shm is not None else None]
            dist.broadcast_object_list(box, src=src, group=group)
            if box[0] is None:
                self.error = getattr(self, "error", "the source rank could not create the shared segment")
                return
            if rank != src:
                self._shm = shared_memory.SharedMemory(name=box[0])
                try:        # the creator owns the name; keep Python's tracker from unlinking it a second time at exit
                    resource_tracker.unregister(self._shm._name, "shared_memory")
                except Exception:
                    pass
            dist.barrier(group=group)                    # everybody has mapped the segment
        if self._shm is None:
            return
        if rank == src:
            self._shm.unlink()
        self.array = np.ndarray(self.shape, self.dtype, buffer=self._shm.buf)
        self.ptr = ctypes.addressof(ctypes.c_char.from_buffer(self._shm.buf))
        if register is None:
            register = torch.cuda.is_available()
        if register:
            from . import _lib
            _lib.call("psa_host_register", self.ptr, nbytes)
            self._registered = True

    @property
    def available(self) -> bool:
        return self.array is not None

    def close(self) -> None:
        if getattr(self, "_shm", None) is None:
            return
        if self._registered:
            from . import _lib
            try:
                _lib.call("psa_host_unregister", self.ptr)
            except Exception:
                pass
            self._registered = False
        self.array = None
        try:
            self._shm.close()
        except BufferError:       # a NumPy view handed to the caller is still alive: the mapping stays until it goes
            pass
        self._shm = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------- small collectives
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
    with :func:`shard_range`) into ``(n_f, n_k_total, ...)`` on ``dst``; other ranks get ``None``.  (Device-resident
    gather for callers that want the whole result on one GPU; ``calculate_sharded`` does not funnel through it.)"""
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
    per plane when the blocks have equal length, one broadcast per (plane, rank) otherwise.  (The NCCL fallback
    of the fused peer-store exchange.)"""
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


# ---------------------------------------------------------------------------------------------- CUDA IPC mappings
# A peer allocation can be mapped once per process at a time, and the caching allocator may carve several of our
# buffers out of one allocation: mappings are shared process-wide and reference counted.
_IPC_MAPPED: Dict[bytes, List[int]] = {}          # handle -> [base address, references]


def _ipc_map(handle: bytes) -> int:
    from . import _lib
    hit = _IPC_MAPPED.get(handle)
    if hit is None:
        out = ctypes.c_void_p(0)
        buf = ctypes.create_string_buffer(handle, 64)
        _lib.call("psa_ipc_open", ctypes.addressof(buf), ctypes.addressof(out))
        hit = _IPC_MAPPED[handle] = [int(out.value), 0]
    hit[1] += 1
    return hit[0]


def _ipc_unmap(handle: bytes) -> None:
    from . import _lib
    hit = _IPC_MAPPED.get(handle)
    if hit is None:
        return
    hit[1] -= 1
    if hit[1] <= 0:
        del _IPC_MAPPED[handle]
        try:
            _lib.call("psa_ipc_close", hit[0])
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------- peer-mapped planes
class PeerPlanes:
    """Digit planes + exponents of one atom selection on this rank, plus the same buffers of every peer rank mapped
    into this process through CUDA IPC.  Built once per (calculator, selection) and reused by every ingest: the
    buffers must outlive the peers' mappings, so they are owned here and never returned to the allocator."""

    def __init__(self, eng, n_t: int, pitch: int, group=None):
        from . import _lib
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.dig = eng.empty((3, 4, n_t, pitch), torch.int8)
        self.expo = eng.empty((3, n_t), torch.int32)
        mine, failed = [], None
        try:
            for t in (self.dig, self.expo):
                handle = (ctypes.c_char * 64)()
                off = ctypes.c_int64(0)
                _lib.call("psa_ipc_export", t.data_ptr(), ctypes.addressof(handle), ctypes.addressof(off))
                mine.append((bytes(handle), int(off.value)))
        except (RuntimeError, ValueError, NotImplementedError) as exc:
            mine, failed = None, exc
        everyone: List = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)           # every rank reaches this, failed or not
        self._opened: Dict[bytes, int] = {}
        if any(e is None for e in everyone):
            raise RuntimeError(f"CUDA IPC export failed on a rank ({failed})")
        ptrs = [[0] * self.world, [0] * self.world]
        for r, pair in enumerate(everyone):
            for which, (handle, off) in enumerate(pair):
                if r == self.rank:
                    ptrs[which][r] = (self.dig, self.expo)[which].data_ptr()
                    continue
                base = self._opened.get(handle)
                if base is None:
                    base = self._opened[handle] = _ipc_map(handle)
                ptrs[which][r] = base + off
        self.dig_ptrs = (ctypes.c_void_p * self.world)(*ptrs[0])
        self.expo_ptrs = (ctypes.c_void_p * self.world)(*ptrs[1])

    def close(self) -> None:
        for handle in self._opened:
            _ipc_unmap(handle)
        self._opened = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _peer_planes(calc, key, n_t: int, pitch: int, group=None) -> Optional[PeerPlanes]:
    """The cached :class:`PeerPlanes` of a selection, or ``None`` when peer mapping is not possible on every rank
    (not NCCL/CUDA, more than 8 ranks, IPC refused - e.g. an allocator using virtual-memory segments)."""
    cache = calc.__dict__.setdefault("_peer_planes", {})
    if key in cache:
        return cache[key]
    world = dist.get_world_size(group)
    planes, ok = None, 1
    if os.environ.get("PSA_B200_PEER_STORES", "1") == "0" or world > 8 or dist.get_backend(group) != "nccl":
        ok = 0
    flag = torch.tensor([ok], device=calc.engine.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 1:
        try:
            planes = PeerPlanes(calc.engine, n_t, pitch, group)
        except (RuntimeError, ValueError, NotImplementedError):
            planes = None
        flag = torch.tensor([1 if planes is not None else 0], device=calc.engine.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            planes = None
    cache[key] = planes
    return planes


def _stream_fence(device, group=None) -> None:
    """Order the ranks' streams: nothing queued after this call on any rank starts before everything queued before
    it on every rank has finished (a one-element all-reduce; stream-ordered, the host does not wait)."""
    dist.all_reduce(torch.zeros(1, device=device), group=group)


def _upload_and_mean(calc, local_rows, group, mark):
    """This rank's frame range on the device + the float32 mean positions of the WHOLE trajectory on every rank (the
    ordered chain of running sums).  Returns ``(rows to project, mean, frame bounds of every rank)``."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    eng, dtraj = calc.engine, calc.device_trajectory
    n_t, n_a = dtraj.n_t, dtraj.n_a
    bounds = [shard_range(n_t, r, world) for r in range(world)]
    t0, t1 = bounds[rank]
    disp = calc.use_displacements
    pos_rows = dtraj.upload_rows("pos", t0, t1, None if local_rows is None else local_rows[0])
    data_rows = pos_rows if disp else dtraj.upload_rows("vel", t0, t1, None if local_rows is None else local_rows[1])
    mark("upload")
    acc = chain_running_sum(torch.zeros((n_a, 3), dtype=torch.float32, device=eng.device),
                            lambda a, last: eng.mean_accumulate(pos_rows, a, n_t if last else 0), group)
    dtraj.install_mean(acc)
    mark("mean_chain")
    return data_rows, acc, bounds


def sliced_ingest(calc, proj_groups, local_rows=None, group=None, marks=None, pipeline: Optional[bool] = None) -> None:
    """k-independent state from a trajectory whose FRAMES are spread over the ranks.

    Rank r uploads only frames ``shard_range(n_t, r, world)`` - 1/N of the bytes over its own PCIe link -
    and the ranks then build the shared state together:

    * mean positions: the float32 sum of a column must run in frame order to stay bit-identical with
      NumPy, so the running sums travel down the ranks (one (n_atoms, 3) message per hop); the last rank
      divides and broadcasts the mean;
    * digit planes: every frame row is digitised independently (its own exponent); each rank's kernel stores its
      rows into every rank's planes through NVLink (fallback: NCCL all-gather per plane).

    ``local_rows = (positions[t0:t1], velocities[t0:t1])`` (host arrays or CUDA tensors) when this process holds
    nothing but its range; otherwise the range is sliced out of ``calc.traj``.  On return every rank's device
    trajectory has the mean and the digit planes of ``proj_groups`` installed, exactly as after a one-GPU ingest.
    ``marks``: optional callable ``marks(name)`` recording a stage boundary on the stream (bench breakdown).

    ``pipeline`` (default on; ``PSA_B200_PIPELINE_EXCHANGE=0`` turns it off): the peer stores run as a ring on a side
    stream - step s sends this rank's rows to rank ``r + s`` while rank ``r - s``'s rows arrive - and the digit planes
    are installed with that arrival schedule, so the first k-chunk's projection starts on the frames already present
    (its own, then one peer's range per step) instead of waiting for the whole exchange (``engine.sed_on_device``).
    """
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    eng, dtraj = calc.engine, calc.device_trajectory
    mark = marks or (lambda name: None)
    if pipeline is None:
        pipeline = os.environ.get("PSA_B200_PIPELINE_EXCHANGE", "1") != "0"
    main = torch.cuda.current_stream(eng.device) if eng.device.type == "cuda" else None
    if main is not None and eng._comm_stream is not None:
        main.wait_stream(eng._comm_stream)              # an exchange nobody consumed must not run into this one
    # skip only if EVERY rank already holds the state (a rank that computed something on its own must not
    # leave the others waiting in the chain)
    ready = torch.tensor([1 if dtraj.has_state(proj_groups, calc.use_displacements) else 0], device=eng.device)
    dist.all_reduce(ready, op=dist.ReduceOp.MIN, group=group)
    if int(ready.item()) == 1:
        return
    n_t, n_a = dtraj.n_t, dtraj.n_a
    disp = calc.use_displacements
    data_rows, acc, bounds = _upload_and_mean(calc, local_rows, group, mark)
    t0, t1 = bounds[rank]

    for g in proj_groups:
        idx, idx_dev, n_sel = dtraj.selection(g, disp)
        pitch = int(eng_pitch(n_sel))
        key = dtraj._group_key(g, disp)[0]
        peers = _peer_planes(calc, (key, n_t, pitch), n_t, pitch, group)
        if peers is not None and pipeline and world > 1 and all(a % 4 == 0 for a, _ in bounds):
            _stream_fence(eng.device, group)            # nobody still projects from the planes of a previous ingest
            own = ((ctypes.c_void_p * 1)(peers.dig_ptrs[rank]), (ctypes.c_void_p * 1)(peers.expo_ptrs[rank]))
            mean_ptr = acc.data_ptr() if disp else None
            w_ptr = None if dtraj.weight is None else dtraj.weight.data_ptr()
            i_ptr = None if idx_dev is None else idx_dev.data_ptr()
            eng._run("psa_digitize_rows_peers", 1, data_rows.data_ptr(), mean_ptr, w_ptr, i_ptr, t1 - t0, n_a, n_sel, pitch,
                     ctypes.addressof(own[0]), ctypes.addressof(own[1]), 1, n_t, t0, 0, eng.stream())
            arrivals = [(t0, t1, None)]
            comm = eng.comm_stream
            comm.wait_stream(main)                      # rows uploaded, mean final, fence passed
            data_rows.record_stream(comm)
            with torch.cuda.stream(comm):
                for step in range(1, world):
                    dst, src_rank = (rank + step) % world, (rank - step) % world
                    one = ((ctypes.c_void_p * 1)(peers.dig_ptrs[dst]), (ctypes.c_void_p * 1)(peers.expo_ptrs[dst]))
                    eng._run("psa_digitize_rows_peers", 1, data_rows.data_ptr(), mean_ptr, w_ptr, i_ptr, t1 - t0, n_a, n_sel,
                             pitch, ctypes.addressof(one[0]), ctypes.addressof(one[1]), 1, n_t, t0, 1, eng.stream())
                    _stream_fence(eng.device, group)    # every rank's step has landed: rank r - step's rows are here
                    ev = torch.cuda.Event()
                    ev.record(comm)
                    arrivals.append((bounds[src_rank][0], bounds[src_rank][1], ev))
            dtraj.install_group(idx, disp, peers.dig, peers.expo, arrivals=arrivals)
            continue
        if peers is not None:
            _stream_fence(eng.device, group)            # nobody still projects from the planes of a previous ingest
            eng._run("psa_digitize_rows_peers", 1, data_rows.data_ptr(), acc.data_ptr() if disp else None,
                     None if dtraj.weight is None else dtraj.weight.data_ptr(),
                     None if idx_dev is None else idx_dev.data_ptr(), t1 - t0, n_a, n_sel, pitch,
                     ctypes.addressof(peers.dig_ptrs), ctypes.addressof(peers.expo_ptrs), world, n_t, t0, 0, eng.stream())
            _stream_fence(eng.device, group)            # every rank's rows have landed everywhere
            dig, expo = peers.dig, peers.expo
        else:
            dig = eng.empty((3, 4, n_t, pitch), torch.int8)
            expo = eng.empty((3, n_t), torch.int32)
            eng.digitize_rows(data_rows, acc if disp else None, idx_dev, n_sel, pitch, dig, expo, n_t, t0, dtraj.weight)
            exchange_row_blocks(list(dig.view(12, n_t, pitch).unbind(0)) + list(expo.unbind(0)), bounds, group)
        dtraj.install_group(idx, disp, dig, expo)
    mark("digitize_exchange")


# ---------------------------------------------------------------------------------------------- frame-sharded path
FRAME_K_CAP = 2048        # k-points per projection launch of the frame-sharded path (one owner's chunk)
FRAME_K_CAP_STREAMED = 1024   # ... when the chunks stream out to host memory while the next one is projected


class PeerBuffers:
    """``count`` equal device buffers of this rank plus the same buffers of every peer rank mapped into this process
    through CUDA IPC (``ptrs[i][rank]`` = device address of buffer i on that rank).  Collective; the buffers must
    outlive the peers' mappings, so they are owned here and reused by every call that fits into them."""

    def __init__(self, eng, nbytes: int, count: int, group=None):
        from . import _lib
        self.world, self.rank, self.nbytes = dist.get_world_size(group), dist.get_rank(group), int(nbytes)
        self.local = [eng.empty((self.nbytes,), torch.uint8) for _ in range(count)]
        mine, failed = [], None
        try:
            for t in self.local:
                handle = (ctypes.c_char * 64)()
                off = ctypes.c_int64(0)
                _lib.call("psa_ipc_export", t.data_ptr(), ctypes.addressof(handle), ctypes.addressof(off))
                mine.append((bytes(handle), int(off.value)))
        except (RuntimeError, ValueError, NotImplementedError) as exc:
            mine, failed = None, exc
        everyone: List = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)           # every rank reaches this, failed or not
        self._opened: Dict[bytes, int] = {}
        if any(e is None for e in everyone):
            raise RuntimeError(f"CUDA IPC export failed on a rank ({failed})")
        self.ptrs = [[0] * self.world for _ in range(count)]
        for r, handles in enumerate(everyone):
            for i, (handle, off) in enumerate(handles):
                if r == self.rank:
                    self.ptrs[i][r] = self.local[i].data_ptr()
                    continue
                base = self._opened.get(handle)
                if base is None:
                    base = self._opened[handle] = _ipc_map(handle)
                self.ptrs[i][r] = base + off

    def close(self) -> None:
        for handle in self._opened:
            _ipc_unmap(handle)
        self._opened = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def frame_shard_plan(n_k: int, world: int, k_cap: int = FRAME_K_CAP):
    """Who owns which k-points in a frame-sharded run and how the owners' slices are cut into projection chunks.
    Returns ``(slices, n_chunks, chunk)``: ``slices[q]`` = rank q's contiguous k-range (the layout of the k-sharded
    path and of the shared result), ``chunk(q, j)`` = the j-th of ``n_chunks`` balanced pieces of it (possibly empty).
    Every rank evaluates the same plan."""
    slices = [shard_range(n_k, q, world) for q in range(world)]
    longest = max(b - a for a, b in slices)
    n_chunks = max(1, -(-longest // max(1, int(k_cap))))

    def chunk(q: int, j: int) -> Tuple[int, int]:
        a, b = slices[q]
        c0, c1 = shard_range(b - a, j, n_chunks)
        return a + c0, a + c1

    return slices, n_chunks, chunk


def routed_chunk_rows(chunk, world: int, j: int) -> Tuple[np.ndarray, List[int]]:
    """One routed projection launch covers chunk j of EVERY owner: returns the k indices in launch order (owner 0's
    piece, owner 1's, ...) and ``row_begin`` (``world + 1`` entries): projection rows ``[row_begin[q], row_begin[q + 1])``
    (two per k-point: cos and sin) belong to owner q and become rows 0.. of its buffer."""
    pieces = [chunk(q, j) for q in range(world)]
    order = np.concatenate([np.arange(a, b, dtype=np.int64) for a, b in pieces]) if pieces else np.zeros(0, np.int64)
    row_begin = [0]
    for a, b in pieces:
        row_begin.append(row_begin[-1] + 2 * (b - a))
    return order, row_begin


def _all_agree(ok: bool, device, group=None) -> bool:
    flag = torch.tensor([1 if ok else 0], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return int(flag.item()) == 1


def _frame_buffers(calc, nbytes: int, group=None) -> Optional[PeerBuffers]:
    """Three peer-mapped projection buffers of at least ``nbytes`` (cached on the calculator, grown collectively), or
    ``None`` when peer mapping is not possible on every rank."""
    eng = calc.engine
    world = dist.get_world_size(group)
    if os.environ.get("PSA_B200_PEER_STORES", "1") == "0" or world > 8 or dist.get_backend(group) != "nccl":
        return None
    have = calc.__dict__.get("_frame_buffers")
    if have is False:
        return None
    if have is not None and have.nbytes >= nbytes:
        return have
    if have is not None:                     # grow: every rank unmaps its peers before anybody frees
        have.close()
        torch.cuda.synchronize(eng.device)
        dist.barrier(group=group)
        calc.__dict__["_frame_buffers"] = have = None
    try:
        have = PeerBuffers(eng, nbytes, 3, group)
    except (RuntimeError, ValueError, NotImplementedError):
        have = None
    if not _all_agree(have is not None, eng.device, group):
        have = None
    calc.__dict__["_frame_buffers"] = have if have is not None else False
    return have


def frame_sharded_sed(calc, k_vecs: np.ndarray, proj_groups, complex_out: bool, local_rows=None, group=None,
                      host_out=None, marks=None, k_chunk_size: int = 500) -> Tuple[bool, Optional[torch.Tensor]]:
    """The SED of ``k_vecs`` with the FRAMES of the trajectory spread over the ranks - no digit plane ever leaves the
    GPU that produced it.

    Rank r uploads and digitises frames ``shard_range(n_t, r, N)`` only (the float32 mean still comes from the ordered
    chain of running sums).  Every rank then projects ITS frames for ALL k-points, owner by owner: the k-points are
    owned as in the k-sharded layout (``shard_range(n_k, q, N)``), and the projection kernel's epilogue stores each
    tile straight into the owner's projection buffer through NVLink (CUDA-IPC mapped, column offset = this rank's first
    frame) - the frames->k transpose, an all-to-all of 24 bytes per (k, frame), is fused into the tensor-core kernel;
    rank r works on owner ``r + s`` at step s, so every NVLink port carries one stream at a time.  One stream-ordered
    one-element all-reduce per chunk (on a side stream, under the next chunk's projection; three rotating buffers)
    tells the owners that all frames of a chunk have landed; they run the time FFT + assembly on their k-slice.

    Compared with all-gathering the digit planes (k-sharded path) this moves ``2 n_k / (N n_atoms)`` of the bytes
    (C4 on 8 GPUs: 0.43 GB instead of 2.4 GB per rank, C5: 0.09 GB instead of 22 GB) and needs 1/N of the plane memory.
    Bit-identical to one GPU: every (k, frame) of the projection is an independent exact sum.

    Returns ``(True, result)`` - ``result`` = this rank's k-slice ``(n_t, n_k_local[, 3])`` on the device, or ``None``
    when ``host_out`` (a :class:`engine.HostTarget` whose ``k_offset`` is this rank's first k) received it - or
    ``(False, None)`` when this path cannot run (no peer mapping, frame ranges not multiples of 4, a rank without
    frames): the caller falls back to the k-sharded path.  The decision is collective."""
    from . import _lib
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    eng, dtraj = calc.engine, calc.device_trajectory
    mark = marks or (lambda name: None)
    n_t, n_a, n_k = dtraj.n_t, dtraj.n_a, int(k_vecs.shape[0])
    bounds = [shard_range(n_t, r, world) for r in range(world)]
    if any(a % 4 or b <= a for a, b in bounds) or n_k == 0:
        return False, None
    # device-resident result: big launches (whole waves of projection tiles) and the FFT of a chunk issued one chunk
    # late, so that nobody ever waits for the fence; streamed to the host: smaller chunks, transformed as soon as
    # their frames have landed - every copy starts one chunk earlier and the tail after the last projection is short
    lag = 1 if host_out is None else 0
    k_cap = (FRAME_K_CAP if lag else FRAME_K_CAP_STREAMED) if k_chunk_size >= 500 else max(1, int(k_chunk_size))
    slices, n_chunks, chunk = frame_shard_plan(n_k, world, k_cap)
    kc = max(chunk(q, j)[1] - chunk(q, j)[0] for q in range(world) for j in range(n_chunks))
    n_groups = len(proj_groups)
    ldp = (n_t + 3) // 4 * 4
    p_rows = 2 * kc
    group_stride = p_rows * 3 * ldp
    bufs = _frame_buffers(calc, n_groups * group_stride * 4, group)
    if bufs is None:
        return False, None
    f0, f1 = bounds[rank]
    n_loc = f1 - f0
    disp = calc.use_displacements
    main = torch.cuda.current_stream(eng.device)
    comm = eng.comm_stream
    main.wait_stream(comm)
    _stream_fence(eng.device, group)          # no owner still transforms out of the buffers this call will overwrite

    # ---- k-independent state: mean positions (all ranks) + digit planes of this rank's frames (cached)
    local = dtraj._frame_local
    keys = [dtraj._group_key(g, disp)[0] for g in proj_groups]
    cached = dtraj._mean is not None and local.get("bounds") == (f0, f1) and all(k in local.get("groups", {}) for k in keys)
    if not _all_agree(cached, eng.device, group):
        data_rows, mean, _ = _upload_and_mean(calc, local_rows, group, mark)
        local.clear()
        local.update(bounds=(f0, f1), groups={})
        for key, g in zip(keys, proj_groups):
            _, idx_dev, n_sel = dtraj.selection(g, disp)
            dig, expo, pitch = eng.digitize(data_rows, mean if disp else None, idx_dev, n_sel, dtraj.weight)
            local["groups"][key] = (idx_dev, n_sel, pitch, dig, expo)
        mark("digitize_exchange")
    mean = dtraj.mean
    entries = [local["groups"][k] for k in keys]

    # ---- projection of the local frames, owner by owner, stored into the owners' buffers
    k0_own, k1_own = slices[rank]
    n_k_own = k1_own - k0_own
    shape = (n_t, n_k_own, 3) if complex_out else (n_t, n_k_own)
    dtype = torch.complex64 if complex_out else torch.float32
    mode = _lib.MODE_COHERENT if complex_out else _lib.MODE_INCOHERENT
    window = dtraj.window
    out = None
    if host_out is None:
        out = torch.empty(shape, dtype=dtype, device=eng.device)
    else:
        elem = 24 if complex_out else 4
        n_rows = max(0, min(int(host_out.n_rows), n_t))
        chunk_bufs = [torch.empty((n_t, kc) + shape[2:], dtype=dtype, device=eng.device) for _ in range(2)]
        drained = [None, None]
        copy = eng.copy_stream
    kv_dev = eng.upload_small(np.ascontiguousarray(k_vecs, np.float32))
    adig_bufs: Dict[int, torch.Tensor] = {}
    p_local = [b.view(torch.float32) for b in bufs.local]
    fences: List[Optional[torch.cuda.Event]] = [None] * n_chunks
    n_done = 0

    def transform(j: int) -> None:
        """FFT + assembly of this rank's chunk j (all frames have landed: fence j has passed)."""
        nonlocal n_done
        main.wait_event(fences[j])
        ka, kb = chunk(rank, j)
        nk = kb - ka
        if nk == 0:
            return
        P = p_local[j % 3]
        if host_out is None:
            eng.fft_sed(P, n_groups, group_stride, nk, n_t, ldp, mode, out, n_k_own, ka - k0_own, window)
            return
        b = n_done & 1
        n_done += 1
        if drained[b] is not None:
            main.wait_event(drained[b])
        eng.fft_sed(P, n_groups, group_stride, nk, n_t, ldp, mode, chunk_bufs[b], kc, 0, window)
        ready = torch.cuda.Event()
        ready.record(main)
        copy.wait_event(ready)
        _lib.call("psa_copy_rows", host_out.ptr + (host_out.k_offset + ka - k0_own) * elem, host_out.n_k_total * elem,
                  chunk_bufs[b].data_ptr(), kc * elem, nk * elem, n_rows, copy.cuda_stream)
        drained[b] = torch.cuda.Event()
        drained[b].record(copy)

    # One launch per chunk for the k-points of ALL owners (the kernel routes every projection row to the buffer of the
    # rank that owns it): whole waves of tiles and full-width tiles across owner boundaries.  PSA_B200_FRAMES_ROUTED=0
    # (or the CUDA-core cross-check kernel) falls back to one launch per owner, rank r starting with owner r.
    routed = eng.project_impl == _lib.PROJECT_TENSOR and os.environ.get("PSA_B200_FRAMES_ROUTED", "1") != "0"
    kv_routed: List[Optional[torch.Tensor]] = []
    if routed:
        routes = [routed_chunk_rows(chunk, world, j) for j in range(n_chunks)]
        for order, _ in routes:
            kv_routed.append(eng.upload_small(np.ascontiguousarray(k_vecs[order], np.float32)) if order.size else None)
        rows_all = 2 * world * kc
    for j in range(n_chunks):
        if routed:
            order, row_begin = routes[j]
            begin = (ctypes.c_int64 * (world + 1))(*row_begin)
            n_all = int(order.size)
            for g, (idx_dev, n_sel, pitch, dig, expo) in enumerate(entries if n_all else []):
                adig = adig_bufs.get(pitch)
                if adig is None:
                    adig = adig_bufs[pitch] = eng.empty((4, rows_all, pitch), torch.int8)
                eng.phase_digits(kv_routed[j], mean, idx_dev, n_sel, pitch, rows_all, out=adig)
                dests = (ctypes.c_void_p * world)(*[bufs.ptrs[j % 3][q] + 4 * (g * group_stride + f0)
                                                    for q in range(world)])
                eng._run("psa_project_routed", -(-n_sel // 32768), adig.data_ptr(), 2 * n_all, rows_all, dig.data_ptr(),
                         expo.data_ptr(), n_loc, n_sel, pitch, ctypes.addressof(dests), ctypes.addressof(begin), world,
                         ldp, eng.stream(), label="psa_project")
        for s in range(world if not routed else 0):
            q = (rank + s) % world
            ka, kb = chunk(q, j)
            nk = kb - ka
            if nk == 0:
                continue
            for g, (idx_dev, n_sel, pitch, dig, expo) in enumerate(entries):
                adig = adig_bufs.get(pitch)
                if adig is None:
                    adig = adig_bufs[pitch] = eng.empty((4, p_rows, pitch), torch.int8)
                eng.phase_digits(kv_dev[ka:kb], mean, idx_dev, n_sel, pitch, p_rows, out=adig)
                dst = bufs.ptrs[j % 3][q] + 4 * (g * group_stride + f0)
                eng._run("psa_project", -(-n_sel // 32768), adig.data_ptr(), 2 * nk, p_rows, dig.data_ptr(),
                         expo.data_ptr(), n_loc, n_sel, pitch, dst, ldp, eng.project_impl, eng.stream())
        if lag and j >= 1:
            transform(j - 1)
        done = torch.cuda.Event()
        done.record(main)
        comm.wait_event(done)
        with torch.cuda.stream(comm):
            _stream_fence(eng.device, group)          # every rank's stores of chunk j (and its FFT of j - 1) are done
            fences[j] = torch.cuda.Event()
            fences[j].record(comm)
        if not lag:
            transform(j)
    if lag:
        transform(n_chunks - 1)
    if host_out is not None:
        for b in chunk_bufs:
            b.record_stream(eng.copy_stream)
    return True, out


last_path: Optional[str] = None      # which path the last sharded_sed_on_device / calculate_sharded call took


def frames_mode() -> bool:
    """Frame-sharded multi-GPU path on (default) or off (``PSA_B200_SHARD=k``: all-gather the digit planes, shard k)."""
    return os.environ.get("PSA_B200_SHARD", "frames").lower() != "k"


def sharded_sed_on_device(calc, k_vecs: np.ndarray, proj_groups, complex_out: bool, local_rows=None, group=None,
                          k_chunk_size: int = 500) -> torch.Tensor:
    """This rank's k-slice ``shard_range(n_k, rank, N)`` of the SED, device-resident, from a trajectory whose frames
    are spread over the ranks: :func:`frame_sharded_sed` when it can run, else :func:`sliced_ingest` + the one-GPU
    pipeline on the slice.  (What ``bench.py`` times as ``value`` on several GPUs.)"""
    from .engine import sed_on_device
    global last_path
    if frames_mode():
        ok, out = frame_sharded_sed(calc, k_vecs, proj_groups, complex_out, local_rows, group, k_chunk_size=k_chunk_size)
        if ok:
            last_path = "frames"
            return out
    last_path = "k"
    sliced_ingest(calc, proj_groups, local_rows, group)
    k0, k1 = shard_range(int(k_vecs.shape[0]), dist.get_rank(group), dist.get_world_size(group))
    return sed_on_device(calc.device_trajectory, k_vecs[k0:k1], proj_groups, complex_out, calc.use_displacements,
                         k_chunk=k_chunk_size)


def shared_result(calc, shape, np_dtype, src: int = 0, group=None) -> SharedHostArray:
    """The shared page-locked result array of this shape, cached on the calculator (mapping + pinning a multi-GB
    array is setup, like allocating the pinned input buffers)."""
    cache = calc.__dict__.setdefault("_shared_results", {})
    key = (tuple(shape), np.dtype(np_dtype).str)
    shared = cache.get(key)
    if shared is None:
        shared = cache[key] = SharedHostArray(shape, np_dtype, src=src, group=group)
    return shared


def calculate_sharded(calc, k_points_mags: np.ndarray, k_vectors_3d: np.ndarray, basis_atom_indices=None,
                      basis_atom_types=None, summation_mode: str = "coherent", k_grid_shape=None, src: int = 0,
                      group=None, ingest: str = "broadcast", local_rows=None, k_chunk_size: int = 500,
                      timings: Optional[Dict[str, float]] = None):
    """``SEDCalculator.calculate`` over all ranks of the process group.

    Every rank calls this with a calculator built on a trajectory of the right *shape*.  With
    ``ingest="broadcast"`` only ``src`` needs real positions/velocities (the others may hold zero-stride
    placeholders): it uploads and ingests everything and broadcasts the result.  With ``ingest="sliced"``
    every rank holds (at least) its own range of frames and uploads just that, see :func:`sliced_ingest`.
    Every rank streams its k-slice of the spectra into ONE result array in shared page-locked host memory.
    Returns the ``SED`` on ``src`` (its ``sed`` is a view of that shared array, valid until the next call with
    the same result shape on this calculator) and ``None`` elsewhere.
    ``timings`` (a dict) receives a per-stage breakdown in milliseconds of this rank's stream.
    """
    from . import groups as grp
    from .engine import HostTarget, sed_on_device
    from .sed import SED

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return calc.calculate(k_points_mags, k_vectors_3d, basis_atom_indices, basis_atom_types,
                              summation_mode, k_grid_shape, k_chunk_size)
    if summation_mode not in ("coherent", "incoherent"):
        raise ValueError(f"summation_mode must be 'coherent' or 'incoherent', got {summation_mode}")
    if ingest not in ("broadcast", "sliced"):
        raise ValueError(f"ingest must be 'broadcast' or 'sliced', got {ingest}")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    traj = calc.traj
    eng, dtraj = calc.engine, calc.device_trajectory
    groups = grp.resolve_sed_groups(traj.types, traj.n_atoms, basis_atom_indices, basis_atom_types, summation_mode)
    complex_out, proj_groups = grp.plan_sed_groups(groups, summation_mode)
    k_vecs = np.ascontiguousarray(np.asarray(k_vectors_3d, dtype=np.float32).reshape(-1, 3))
    n_k, n_t = k_vecs.shape[0], traj.n_frames

    events: List[Tuple[str, torch.cuda.Event]] = []

    def mark(name: str) -> None:
        if timings is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(eng.device))
            events.append((name, ev))

    with torch.cuda.device(eng.device):
        # the result array, shared by the ranks and page-locked (cached per shape)
        shape = (n_t, n_k, 3) if complex_out else (n_t, n_k)
        shared = shared_result(calc, shape, np.complex64 if complex_out else np.float32, src, group)
        mark("start")
        k0, k1 = shard_range(n_k, rank, world)
        # 0. frames spread over the ranks and a shared result array: nothing but projections crosses NVLink
        streamed = False
        if ingest == "sliced" and frames_mode() and shared.available and n_t > 0:
            streamed, _ = frame_sharded_sed(calc, k_vecs, proj_groups, complex_out, local_rows, group,
                                            host_out=HostTarget(shared.ptr, n_k, k0, n_t, shared), marks=mark,
                                            k_chunk_size=k_chunk_size)
        # 1. k-independent state on every rank
        global last_path
        last_path = "frames" if streamed else "k"
        if streamed:
            pass
        elif ingest == "sliced":
            sliced_ingest(calc, proj_groups, local_rows, group, marks=mark)
        else:
            tensors: List[Optional[torch.Tensor]] = []
            metas: List[Tuple] = []
            if rank == src:
                tensors.append(dtraj.mean)
                for g in proj_groups:
                    _, _, _, dig, expo = dtraj.group(g, calc.use_displacements)
                    tensors += [dig, expo]
                metas = [(tuple(t.shape), t.dtype) for t in tensors]
            got = broadcast_tensors(tensors, metas, src, eng.device, group)
            if rank != src:
                dtraj.install_mean(got[0])
                for i, g in enumerate(proj_groups):
                    dtraj.install_group(g, calc.use_displacements, got[1 + 2 * i], got[2 + 2 * i])
            mark("ingest_broadcast")

        # 2. every rank: its contiguous k-slice, no communication; spectra stream out over this rank's own PCIe link
        if not shared.available:
            # no shared segment on this box: the slices are gathered on the source GPU and leave through its link
            local = sed_on_device(dtraj, k_vecs[k0:k1], proj_groups, complex_out, calc.use_displacements,
                                  k_chunk=k_chunk_size)
            full = gather_k_slices(local, n_k, dst=src, group=group)
            if rank != src:
                return None
            return SED(calc._to_host(full), np.fft.fftfreq(n_t, d=calc.dt_ps), k_points_mags, k_vectors_3d,
                       k_grid_shape=k_grid_shape, is_complex=complex_out, phase=None, context=calc._context(groups))
        if k1 > k0 and n_t > 0 and not streamed:
            sed_on_device(dtraj, k_vecs[k0:k1], proj_groups, complex_out, calc.use_displacements, k_chunk=k_chunk_size,
                          host_out=HostTarget(shared.ptr, n_k, k0, n_t, shared))
        mark("compute")
        eng.copy_stream.synchronize()
        torch.cuda.current_stream(eng.device).synchronize()
        if timings is not None:
            mark("result_drain")
            torch.cuda.current_stream(eng.device).synchronize()
            for (_, a), (name, b) in zip(events[:-1], events[1:]):
                timings[name] = timings.get(name, 0.0) + a.elapsed_time(b)
        dist.barrier(group=group)                          # every slice is in the shared array
    if rank != src:
        return None
    freqs = np.fft.fftfreq(n_t, d=calc.dt_ps)
    return SED(shared.array, freqs, k_points_mags, k_vectors_3d, k_grid_shape=k_grid_shape,
               is_complex=complex_out, phase=None, context=calc._context(groups))
