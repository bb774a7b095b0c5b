"""Device-side orchestration of the SED hot path.

PyTorch is used for exactly three things here: owning device buffers, streams,
and host<->device copies.  Every arithmetic step is a call through the C ABI
(``psa_b200._lib``) into hand-written sm_100a kernels; there is no tensor
arithmetic in this file and no fallback if the library is missing.

Pipeline for one ``calculate`` (reference: src/psa/core/sed_calculator.py:182-336)::

    upload trajectory ─► mean positions ─► digit planes of the projected series   (once per group)
    for each k-chunk:  phase digit planes ─► tensor-core projection ─► FFT + assembly into the result

A k-set longer than one chunk can be streamed: each chunk's spectra leave for pinned host memory on a side
stream while the next chunk is projected (``sed_on_device(host_out=...)``).  On several GPUs the ingest can
be sliced by frames (``DeviceTrajectory.upload_rows`` + ``Engine.mean_accumulate`` / ``digitize_rows``, driven
by ``psa_b200.dist.sliced_ingest``).
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import threading
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

# k-points per chunk (at most 1024 = 2048 projection rows = 16 row tiles); PSA_B200_K_CHUNK lowers it for A/B runs
K_CHUNK_CAP = max(8, min(1024, int(os.environ.get("PSA_B200_K_CHUNK", "1024"))))
K_CHUNK_REFERENCE_DEFAULT = 500   # the reference's default k_chunk_size (sed_calculator.py:185)
_UPLOAD_CHUNK_BYTES = 256 << 20   # pinned staging buffers for uploads from pageable / memory-mapped arrays


# Streamed ingest (one GPU, velocities still in host memory): planning constants only - how many leading k-chunks are
# projected frame range by frame range while the rest of the array is still crossing PCIe.  Results do not depend on them.
_STREAM_MIN_BYTES = int(os.environ.get("PSA_B200_STREAM_MIN_MB", "256")) << 20
_STREAM_H2D_RATE = {True: 50e9, False: 10e9}       # bytes/s from page-locked / pageable host memory
_STREAM_PROJECT_RATE = 2.7e13                      # (k, frame, atom) triples per second of the projection kernel
_STREAM_P_BYTES = 8 << 30                          # ceiling for the projection buffer of the streamed chunks


def effective_k_chunk(k_chunk_size: int, n_k: int) -> int:
    """k-points per projection launch.  The reference's ``k_chunk_size`` bounds the size of its temporaries; here
    the results do not depend on the chunking at all (the contraction is exact), so the knob is only honoured as a
    memory limit when the caller LOWERS it below the reference's default; at the default or above, chunks are as
    large as the projection scratch allows (fewer passes over the digit planes, fewer launch tails)."""
    k_chunk_size = max(1, int(k_chunk_size))
    cap = K_CHUNK_CAP if k_chunk_size >= K_CHUNK_REFERENCE_DEFAULT else k_chunk_size
    cap = max(1, min(cap, n_k, K_CHUNK_CAP))
    if cap < K_CHUNK_REFERENCE_DEFAULT or n_k <= cap:
        return cap
    # equal chunks instead of full ones plus a remainder (10 000 k: 10 x 1000, not 9 x 1024 + 784): every launch fills
    # whole waves of projection tiles, and a streamed result ends with a chunk-sized copy instead of a long tail
    n_chunks = -(-n_k // cap)
    return -(-n_k // n_chunks)


def plan_k_chunks(n_k: int, k_chunk_size: int, first_rows_hint: Optional[Tuple[int, int]] = None) -> List[Tuple[int, int]]:
    """``[(k0, n_k_chunk), ...]``.  Normally equal chunks (:func:`effective_k_chunk`).  ``first_rows_hint =
    (frames_per_range, n_clusters)``: the first chunk will be projected range by range while a multi-GPU exchange is
    in flight, so it gets a whole number of 128-row tiles chosen such that ONE range launch fills whole waves of the
    projection's CTA pairs (r tiles x ceil(frames / 256) x 3 polarisations ~ a multiple of n_clusters)."""
    if n_k <= 0:
        return []
    kc = effective_k_chunk(k_chunk_size, n_k)
    chunks: List[Tuple[int, int]] = []
    k0 = 0
    if first_rows_hint is not None and n_k >= 64 and kc >= 64:
        frames, clusters = first_rows_hint
        per_row_tile = -(-max(1, frames) // 256) * 3
        best_r, best_eff = 1, 0.0
        for r in range(1, min(16, min(n_k, kc) // 64) + 1):
            waves = r * per_row_tile / clusters
            eff = waves / -(-r * per_row_tile // clusters)
            if eff >= best_eff - 0.02 or eff >= 0.95:           # prefer more tiles unless clearly less efficient
                if eff >= 0.95 or eff > best_eff:
                    best_r, best_eff = r, max(eff, best_eff if eff >= 0.95 else eff)
        first = min(n_k, best_r * 64)
        if n_k - first < 128 and n_k <= kc:                     # not worth a second, tiny launch
            first = n_k
        chunks.append((0, first))
        k0 = first
    rest = n_k - k0
    if rest > 0:
        kc_rest = effective_k_chunk(k_chunk_size, rest)
        while k0 < n_k:
            nk = min(kc_rest, n_k - k0)
            chunks.append((k0, nk))
            k0 += nk
    return chunks


@dataclass
class HostTarget:
    """Where a streamed result goes: a page-locked host array of shape ``(n_rows, n_k_total[, 3])`` of which one call
    fills the columns ``[k_offset, k_offset + n_k)`` (reference: ``full_sed_data[:, k0:k1]``, sed_calculator.py:310, 325).
    On several GPUs every rank targets its own column slice of ONE array in shared host memory."""
    ptr: int                 # host address of element [0][0]
    n_k_total: int
    k_offset: int
    n_rows: int
    keepalive: Any = None    # whatever owns the memory


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Engine:
    """Thin, typed wrappers of the C entry points operating on torch CUDA tensors."""

    def __init__(self, device: Optional[int] = None, project_impl: Optional[int] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("psa_b200 needs a CUDA device (B200, sm_100a); there is no CPU path.")
        if device is None:
            device = int(os.environ.get("PSA_B200_DEVICE", torch.cuda.current_device()))
        self.device = torch.device("cuda", device)
        _lib.check(_lib.load().psa_device_check(device))
        if project_impl is None:
            project_impl = int(os.environ.get("PSA_B200_PROJECT_IMPL", _lib.PROJECT_TENSOR))
        self.project_impl = project_impl
        self._plans: Dict[int, torch.Tensor] = {}
        self._fft_ws: Optional[torch.Tensor] = None
        self._copy_stream: Optional[torch.cuda.Stream] = None
        self._comm_stream: Optional[torch.cuda.Stream] = None
        self.launches = 0          # kernels launched through this engine (bench reports it)
        self.profile: Optional[Dict[str, List[Tuple[torch.cuda.Event, torch.cuda.Event, int]]]] = None
        self.timeline: Optional[List[Tuple[str, "torch.cuda.Event"]]] = None   # stage marks of one call (bench breakdown)

    # -- helpers
    def stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    @property
    def copy_stream(self) -> "torch.cuda.Stream":
        """Side stream for result copies that overlap with compute (created on first use)."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        return self._copy_stream

    @property
    def comm_stream(self) -> "torch.cuda.Stream":
        """Side stream of the multi-GPU exchange (peer stores + fences run under the first chunk's projection)."""
        if self._comm_stream is None:          # high priority: its small kernels take the first SM that frees up
            self._comm_stream = torch.cuda.Stream(device=self.device, priority=-1)
        return self._comm_stream

    @property
    def n_clusters(self) -> int:
        """CTA pairs the projection kernel runs (one per two SMs)."""
        return max(1, torch.cuda.get_device_properties(self.device).multi_processor_count // 2)

    def empty(self, shape, dtype) -> torch.Tensor:
        return torch.empty(shape, dtype=dtype, device=self.device)

    def upload_small(self, arr: np.ndarray) -> torch.Tensor:
        """Stream-ordered upload of a small host array.  A plain ``tensor.to(device)`` from pageable memory
        blocks the host until everything queued on the stream has finished (measured: 0.35 ms per call
        inside a 1 ms step); staging through pinned memory keeps the host running ahead of the GPU."""
        staged = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
        return staged.to(self.device, non_blocking=True)

    def _run(self, kernel: str, n_launches: int, *args, label: Optional[str] = None) -> None:
        """One C-ABI call; with ``self.profile`` set, bracketed by CUDA events on the launching stream (and booked
        under ``label`` when a variant entry point should count as the kernel it is a variant of)."""
        self.launches += n_launches
        if self.profile is None:
            _lib.call(kernel, *args)
            return
        stream = torch.cuda.current_stream(self.device)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(stream)
        _lib.call(kernel, *args)
        end.record(stream)
        self.profile.setdefault(label or kernel, []).append((start, end, n_launches))

    def mark(self, label: str, stream: Optional["torch.cuda.Stream"] = None) -> None:
        """With ``self.timeline`` set: a timing event on ``stream`` (default: the current one) under ``label``."""
        if self.timeline is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream if stream is not None else torch.cuda.current_stream(self.device))
            self.timeline.append((label, ev))

    def timeline_ms(self) -> List[Tuple[str, float]]:
        """``[(label, ms since the first mark)]`` of the marks collected so far (synchronises the device)."""
        torch.cuda.synchronize(self.device)
        marks = self.timeline or []
        return [(label, marks[0][1].elapsed_time(ev)) for label, ev in marks]

    def profile_summary(self) -> Dict[str, Dict[str, float]]:
        """{kernel: {ms, calls, launches}} from the events collected while ``self.profile`` was set."""
        torch.cuda.synchronize(self.device)
        out: Dict[str, Dict[str, float]] = {}
        for name, evs in (self.profile or {}).items():
            out[name] = dict(ms=sum(a.elapsed_time(b) for a, b, _ in evs), calls=len(evs),
                             launches=sum(n for _, _, n in evs))
        return out

    # -- kernels
    def mean_positions(self, pos: torch.Tensor) -> torch.Tensor:
        n_t, n_a, _ = pos.shape
        mean = self.empty((n_a, 3), torch.float32)
        self._run("psa_mean_positions", 1, pos.data_ptr(), n_t, n_a, mean.data_ptr(), self.stream())
        return mean

    def digitize(self, data: torch.Tensor, mean: Optional[torch.Tensor], idx: Optional[torch.Tensor],
                 n_sel: int, weight: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, int]:
        n_t, n_a, _ = data.shape
        pitch = int(_lib.load().psa_pitch(n_sel))
        dig = self.empty((3, 4, n_t, pitch), torch.int8)
        expo = self.empty((3, n_t), torch.int32)
        self._run("psa_digitize", 1, data.data_ptr(), _ptr(mean), _ptr(weight), _ptr(idx), n_t, n_a, n_sel, pitch,
                  dig.data_ptr(), expo.data_ptr(), self.stream())
        return dig, expo, pitch

    def mean_accumulate(self, rows: torch.Tensor, acc: Optional[torch.Tensor], divide_by: int) -> torch.Tensor:
        """Running float32 column sums continued over ``rows`` (frames in order); ``/ divide_by`` when > 0."""
        n_rows, n_a, _ = rows.shape
        out = acc if acc is not None else self.empty((n_a, 3), torch.float32)
        self._run("psa_mean_accumulate", 1, rows.data_ptr(), n_rows, n_a, _ptr(acc), divide_by, out.data_ptr(),
                  self.stream())
        return out

    def digitize_rows(self, rows: torch.Tensor, mean: Optional[torch.Tensor], idx: Optional[torch.Tensor], n_sel: int,
                      pitch: int, dig: torch.Tensor, expo: torch.Tensor, n_t_total: int, t0: int,
                      weight: Optional[torch.Tensor] = None) -> None:
        n_rows, n_a, _ = rows.shape
        self._run("psa_digitize_rows", 1, rows.data_ptr(), _ptr(mean), _ptr(weight), _ptr(idx), n_rows, n_a, n_sel, pitch,
                  dig.data_ptr(), expo.data_ptr(), n_t_total, t0, self.stream())

    def phase_digits(self, kvecs: torch.Tensor, mean: torch.Tensor, idx: Optional[torch.Tensor], n_sel: int,
                     pitch: int, rows_alloc: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        n_k = kvecs.shape[0]
        if out is None:
            out = self.empty((4, rows_alloc, pitch), torch.int8)
        self._run("psa_phase_digits", 1, kvecs.data_ptr(), n_k, mean.data_ptr(), _ptr(idx), n_sel, pitch,
                  rows_alloc, out.data_ptr(), self.stream())
        return out

    def project(self, adig: torch.Tensor, rows: int, rows_alloc: int, bdig: torch.Tensor, expo: torch.Tensor,
                n_t: int, n_sel: int, pitch: int, P: torch.Tensor, ldp: int, impl: Optional[int] = None,
                t_range: Optional[Tuple[int, int]] = None) -> None:
        """``t_range = (t0, t1)`` projects those frames only (same bits as the full call, frame by frame)."""
        impl = self.project_impl if impl is None else impl
        if t_range is None:
            self._run("psa_project", -(-n_sel // 32768), adig.data_ptr(), rows, rows_alloc, bdig.data_ptr(),
                      expo.data_ptr(), n_t, n_sel, pitch, P.data_ptr(), ldp, impl, self.stream())
        else:
            t0, t1 = t_range
            self._run("psa_project_rows", -(-n_sel // 32768), adig.data_ptr(), rows, rows_alloc, bdig.data_ptr(),
                      expo.data_ptr(), n_t, t0, t1 - t0, n_sel, pitch, P.data_ptr(), ldp, impl, self.stream())

    def fft_plan(self, n_t: int) -> torch.Tensor:
        """Device-resident FFT plan for ``n_t`` frames (twiddles; chirp tables when n_t is not 2^s), cached."""
        plan = self._plans.get(n_t)
        if plan is None:
            nbytes = int(_lib.load().psa_fft_plan_bytes(n_t))
            if nbytes < 0:
                _lib.check(_lib.ERR_UNSUPPORTED)
            plan = self.empty((nbytes,), torch.uint8)
            self._run("psa_fft_plan_init", 3, n_t, plan.data_ptr(), self.stream())
            self._plans[n_t] = plan
        return plan

    def fft_sed(self, P: torch.Tensor, n_groups: int, group_stride: int, n_k: int, n_t: int, ldp: int,
                mode: int, out: torch.Tensor, n_k_total: int, k_offset: int,
                window: Optional[torch.Tensor] = None) -> None:
        plan = self.fft_plan(n_t)
        ws_bytes = int(_lib.load().psa_fft_workspace_bytes(n_t, n_k, n_groups))
        ws = None
        if ws_bytes > 0:
            if self._fft_ws is None or self._fft_ws.numel() < ws_bytes:
                self._fft_ws = self.empty((ws_bytes,), torch.uint8)
            ws = self._fft_ws
        self._run("psa_fft_sed", 1 if ws is None else 2, P.data_ptr(), n_groups, group_stride, n_k, n_t, ldp,
                  plan.data_ptr(), _ptr(ws), ws_bytes, _ptr(window), mode, out.data_ptr(), n_k_total, k_offset,
                  self.stream())

    def chiral_phase(self, z1: torch.Tensor, z2: torch.Tensor, n: int, stride1: int, stride2: int, opt: str,
                     out: torch.Tensor) -> None:
        self._run("psa_chiral_phase", 1, z1.data_ptr(), z2.data_ptr(), n, stride1, stride2, ord(opt),
                  out.data_ptr(), self.stream())

    def intensity(self, sed: torch.Tensor) -> torch.Tensor:
        n_pol = sed.shape[-1]
        n_rows = sed.numel() // n_pol
        out = self.empty(sed.shape[:-1], torch.float32)
        self._run("psa_intensity", 1, sed.data_ptr(), n_rows, n_pol, out.data_ptr(), self.stream())
        return out


class IngestStream:
    """Frame ranges ``(t0, t1, event)`` of digit planes that are produced while they are consumed: iterating performs
    the next host->device copy + digitise launch on a side stream and yields the event that marks the range valid.
    ``seconds``: planning estimate of the whole transfer."""

    def __init__(self, ranges, seconds: float):
        self._ranges, self.seconds = ranges, float(seconds)

    def __iter__(self):
        return self._ranges


class DeviceTrajectory:
    """A trajectory resident in HBM plus everything derived from it that is k-independent."""

    def __init__(self, engine: Engine, positions: np.ndarray, velocities: np.ndarray,
                 weight: Optional[np.ndarray] = None, window: Optional[np.ndarray] = None):
        """``weight``: per-atom float32 factor multiplied into the projected series (sqrt(mass) for the README's mass
        weighting; ``None`` = the reference's unweighted SED).  ``window``: float32 taper over the frames applied
        before the time FFT (``None`` = rectangular, the reference)."""
        self.engine = engine
        self.n_t, self.n_a = positions.shape[0], positions.shape[1]
        self._host = {"pos": positions, "vel": velocities}
        self._weight_host, self._window_host = weight, window
        self._weight_dev: Optional[torch.Tensor] = None
        self._window_dev: Optional[torch.Tensor] = None
        self._dev: Dict[str, torch.Tensor] = {}
        self._mean: Optional[torch.Tensor] = None
        self._groups: Dict[Tuple, Tuple] = {}
        self._arrivals: Dict[int, List[Tuple[int, int, Optional["torch.cuda.Event"]]]] = {}
        self._frame_local: Dict[str, Any] = {}      # frame-sharded multi-GPU: digit planes of this rank's frames only
        self._indices: Dict[Optional[str], torch.Tensor] = {}
        self._lock = threading.RLock()
        self.h2d_bytes = 0

    # -- uploads (pinned source -> one async copy; pageable source -> chunked pinned staging)
    def _upload(self, which: str) -> torch.Tensor:
        dev = self._dev.get(which)
        if dev is not None:
            return dev
        host = self._host[which]
        if isinstance(host, torch.Tensor) and host.is_cuda:
            dev = host.to(self.engine.device, torch.float32).contiguous()
        else:
            arr = host.numpy() if isinstance(host, torch.Tensor) else np.asarray(host)
            if arr.dtype != np.float32 or not arr.flags.c_contiguous:
                arr = np.ascontiguousarray(arr, dtype=np.float32)
            src = torch.from_numpy(arr)
            dev = torch.empty(src.shape, dtype=torch.float32, device=self.engine.device)
            if src.is_pinned():
                dev.copy_(src, non_blocking=True)
            else:
                flat_src, flat_dst = arr.reshape(-1), dev.view(-1)
                step = max(1, _UPLOAD_CHUNK_BYTES // 4)
                bufs = [torch.empty(min(step, flat_src.size), dtype=torch.float32, pin_memory=True) for _ in range(2)]
                views = [b.numpy() for b in bufs]
                events: List[Optional[torch.cuda.Event]] = [None, None]
                for i, off in enumerate(range(0, flat_src.size, step)):
                    n = min(step, flat_src.size - off)
                    b = i & 1
                    if events[b] is not None:
                        events[b].synchronize()
                    np.copyto(views[b][:n], flat_src[off:off + n])     # plain memcpy: 2x torch's copy_ on the host
                    flat_dst[off:off + n].copy_(bufs[b][:n], non_blocking=True)
                    events[b] = torch.cuda.Event()
                    events[b].record(torch.cuda.current_stream(self.engine.device))
            self.h2d_bytes += src.numel() * 4
        self._dev[which] = dev
        return dev

    def upload_rows(self, which: str, t0: int, t1: int, rows=None) -> torch.Tensor:
        """Frames [t0, t1) of one array on the device (not cached): the per-rank share of a sliced multi-GPU
        ingest.  ``rows`` overrides the host source when this process only holds that range."""
        src = rows if rows is not None else self._host[which][t0:t1]
        if isinstance(src, torch.Tensor) and src.is_cuda:
            return src.to(self.engine.device, torch.float32).contiguous()
        arr = src.numpy() if isinstance(src, torch.Tensor) else np.asarray(src)
        if arr.dtype != np.float32 or not arr.flags.c_contiguous:
            arr = np.ascontiguousarray(arr, dtype=np.float32)
        host = torch.from_numpy(arr)
        dev = torch.empty(host.shape, dtype=torch.float32, device=self.engine.device)
        dev.copy_(host, non_blocking=host.is_pinned())
        self.h2d_bytes += host.numel() * 4
        return dev

    def has_state(self, groups: Sequence[Optional[np.ndarray]], use_displacements: bool) -> bool:
        """True when the mean and the digit planes of every group are already on the device."""
        with self._lock:
            return self._mean is not None and all(self._group_key(g, use_displacements)[0] in self._groups for g in groups)

    def selection(self, idx: Optional[np.ndarray], use_displacements: bool):
        """``(key-normalised idx, idx_dev|None, n_sel)`` of an atom selection."""
        key, idx = self._group_key(idx, use_displacements)
        if idx is None:
            return idx, None, self.n_a
        return idx, self._index_tensor(key, idx), int(idx.size)

    @property
    def weight(self) -> Optional[torch.Tensor]:
        if self._weight_host is not None and self._weight_dev is None:
            self._weight_dev = self.engine.upload_small(np.ascontiguousarray(self._weight_host, np.float32))
        return self._weight_dev

    @property
    def window(self) -> Optional[torch.Tensor]:
        if self._window_host is not None and self._window_dev is None:
            self._window_dev = self.engine.upload_small(np.ascontiguousarray(self._window_host, np.float32))
        return self._window_dev

    @property
    def positions(self) -> torch.Tensor:
        return self._upload("pos")

    @property
    def velocities(self) -> torch.Tensor:
        return self._upload("vel")

    def release(self, which: str) -> None:
        self._dev.pop(which, None)

    def reset_derived(self) -> None:
        """Forget mean positions and digit planes (keeps the raw trajectory resident)."""
        with self._lock:
            self._mean = None
            self._groups.clear()
            self._arrivals.clear()
            self._frame_local.clear()

    @property
    def mean(self) -> torch.Tensor:
        with self._lock:
            if self._mean is None:
                self._mean = self.engine.mean_positions(self.positions)
            return self._mean

    def _group_key(self, idx: Optional[np.ndarray], use_displacements: bool) -> Tuple[Tuple, Optional[np.ndarray]]:
        if idx is not None and idx.size == self.n_a and np.array_equal(idx, np.arange(self.n_a)):
            idx = None
        digest = None if idx is None else hashlib.blake2b(np.ascontiguousarray(idx, np.int64).tobytes(),
                                                          digest_size=16).hexdigest()
        return (use_displacements, digest), idx

    def _index_tensor(self, key: Tuple, idx: np.ndarray) -> torch.Tensor:
        """Device copy of an atom index list; survives ``reset_derived`` (it does not depend on the data)."""
        hit = self._indices.get(key[1])
        if hit is None:
            hit = self._indices[key[1]] = self.engine.upload_small(np.ascontiguousarray(idx, np.int32))
        return hit

    def install_mean(self, mean: torch.Tensor) -> None:
        """Adopt mean positions computed elsewhere (multi-GPU: broadcast from the source rank)."""
        with self._lock:
            self._mean = mean

    def install_group(self, idx: Optional[np.ndarray], use_displacements: bool, dig: torch.Tensor,
                      expo: torch.Tensor, arrivals=None) -> None:
        """Adopt digit planes computed elsewhere for the atom selection ``idx``.  ``arrivals``: frame ranges
        ``(t0, t1, event|None)`` in the order in which they become valid (a multi-GPU exchange still in flight on a
        side stream): the first projection over these planes goes range by range, waiting for each event."""
        key, idx = self._group_key(idx, use_displacements)
        idx_dev = None if idx is None else self._index_tensor(key, idx)
        n_sel = self.n_a if idx is None else int(idx.size)
        with self._lock:
            self._groups[key] = (idx_dev, n_sel, int(dig.shape[-1]), dig, expo)
            if isinstance(arrivals, IngestStream):
                self._arrivals[dig.data_ptr()] = arrivals
            elif arrivals:
                self._arrivals[dig.data_ptr()] = list(arrivals)
            else:
                self._arrivals.pop(dig.data_ptr(), None)

    def pop_arrivals(self, dig: torch.Tensor):
        with self._lock:
            return self._arrivals.pop(dig.data_ptr(), None)

    def prepare(self, groups: Sequence[Optional[np.ndarray]], use_displacements: bool,
                stream: bool = False) -> Tuple[torch.Tensor, List[Tuple]]:
        """Mean positions and the digit planes of every group (both cached).  Running the two passes on
        separate streams was measured (profiles/, round 1): both are HBM-bound and the pair gained < 3 %,
        so they stay on one stream, which also keeps per-kernel timings clean.  ``stream``: the caller consumes
        frame ranges in arrival order (:func:`sed_on_device`), see :meth:`_start_stream`."""
        mean = self.mean
        if stream and not use_displacements:
            self._start_stream(groups)
        return mean, [self.group(g, use_displacements) for g in groups]

    def _start_stream(self, groups: Sequence[Optional[np.ndarray]]) -> None:
        """Velocity mode, raw velocities still in host memory: instead of one upload followed by one digitise pass,
        hand the groups an :class:`IngestStream` - frame ranges are copied and digitised on a side stream one by one
        and the first projection consumes them in that order, so the tensor cores work while PCIe is still busy.
        (Displacements need the final mean, i.e. every position, before the first row can be digitised.)"""
        if os.environ.get("PSA_B200_STREAM_INGEST", "1") == "0" or "vel" in self._dev:
            return
        host = self._host["vel"]
        if isinstance(host, torch.Tensor):
            if host.is_cuda or host.dtype != torch.float32 or not host.is_contiguous():
                return
            src = host
        else:
            arr = np.asarray(host)
            if arr.dtype != np.float32 or not arr.flags.c_contiguous:
                return
            src = torch.from_numpy(arr)
        n_t, n_a = self.n_t, self.n_a
        if src.numel() * 4 < _STREAM_MIN_BYTES or n_t < 1024:
            return
        with self._lock:
            keyed = [self._group_key(g, False) for g in groups]
            if any(key in self._groups for key, _ in keyed):
                return
            eng = self.engine
            pinned = src.is_pinned()
            main, ingest = torch.cuda.current_stream(eng.device), eng.comm_stream
            dev = torch.empty(src.shape, dtype=torch.float32, device=eng.device)
            weight = self.weight
            planes = []
            for key, idx in keyed:
                idx_dev = None if idx is None else self._index_tensor(key, idx)
                n_sel = n_a if idx is None else int(idx.size)
                pitch = int(_lib.load().psa_pitch(n_sel))
                dig = eng.empty((3, 4, n_t, pitch), torch.int8)
                expo = eng.empty((3, n_t), torch.int32)
                planes.append((key, idx_dev, n_sel, pitch, dig, expo))
            ingest.wait_stream(main)            # buffers handed over by the allocator may still be in use on `main`
            for t in [dev] + [p[4] for p in planes] + [p[5] for p in planes]:
                t.record_stream(ingest)
            step = max(256, -(-n_t // 16) // 256 * 256)           # ~16 ranges, whole 256-frame projection tiles

            def ranges():
                flat_src, flat_dst = src.view(-1), dev.view(-1)
                row = n_a * 3
                sub = max(1, _UPLOAD_CHUNK_BYTES // 4)
                bufs, events = None, [None, None]
                if not pinned:
                    bufs = [torch.empty(min(sub, flat_src.numel()), dtype=torch.float32, pin_memory=True) for _ in range(2)]
                n_sub = 0
                for t0 in range(0, n_t, step):
                    t1 = min(n_t, t0 + step)
                    with torch.cuda.stream(ingest):
                        if pinned:
                            dev[t0:t1].copy_(src[t0:t1], non_blocking=True)
                        else:
                            for off in range(t0 * row, t1 * row, sub):
                                n = min(sub, t1 * row - off)
                                b = n_sub & 1
                                n_sub += 1
                                if events[b] is not None:
                                    events[b].synchronize()
                                np.copyto(bufs[b].numpy()[:n], flat_src[off:off + n].numpy())
                                flat_dst[off:off + n].copy_(bufs[b][:n], non_blocking=True)
                                events[b] = torch.cuda.Event()
                                events[b].record(ingest)
                        rows = dev[t0:t1]
                        for _, idx_dev, n_sel, pitch, dig, expo in planes:
                            own = ((ctypes.c_void_p * 1)(dig.data_ptr()), (ctypes.c_void_p * 1)(expo.data_ptr()))
                            # the 128-thread launch shares the SMs with the projection kernel of the previous range
                            eng._run("psa_digitize_rows_peers", 1, rows.data_ptr(), None, _ptr(weight), _ptr(idx_dev),
                                     t1 - t0, n_a, n_sel, pitch, ctypes.addressof(own[0]), ctypes.addressof(own[1]), 1,
                                     n_t, t0, 1, eng.stream())
                        ev = torch.cuda.Event()
                        ev.record(ingest)
                        eng.mark("range_on_device", ingest)
                    yield t0, t1, ev
                self._dev["vel"] = dev
                self.h2d_bytes += src.numel() * 4

            sched = IngestStream(ranges(), src.numel() * 4 / _STREAM_H2D_RATE[pinned])
            for key, idx_dev, n_sel, pitch, dig, expo in planes:
                self._groups[key] = (idx_dev, n_sel, pitch, dig, expo)
                self._arrivals[dig.data_ptr()] = sched

    def group(self, idx: Optional[np.ndarray], use_displacements: bool) -> Tuple:
        """``(idx_dev|None, n_sel, pitch, digits, exponents)`` for an atom selection (cached)."""
        key, idx = self._group_key(idx, use_displacements)
        with self._lock:
            hit = self._groups.get(key)
            if hit is not None:
                return hit
            eng = self.engine
            if idx is None:
                idx_dev, n_sel = None, self.n_a
            else:
                idx_dev = self._index_tensor(key, idx)
                n_sel = int(idx.size)
            if use_displacements:
                dig, expo, pitch = eng.digitize(self.positions, self.mean, idx_dev, n_sel, self.weight)
            else:
                dig, expo, pitch = eng.digitize(self.velocities, None, idx_dev, n_sel, self.weight)
            entry = (idx_dev, n_sel, pitch, dig, expo)
            self._groups[key] = entry
            return entry


def sed_on_device(traj: DeviceTrajectory, k_vecs: np.ndarray, groups: Sequence[Optional[np.ndarray]],
                  complex_out: bool, use_displacements: bool, k_chunk: int = K_CHUNK_CAP,
                  host_out: Optional[HostTarget] = None) -> Optional[torch.Tensor]:
    """Run the projection + FFT pipeline; returns the device-resident result.

    ``groups``: atom index arrays (``None`` = all atoms).  ``complex_out`` -> complex64
    ``(n_t, n_k, 3)`` from the single group; otherwise float32 ``(n_t, n_k)`` summed over groups.

    With ``host_out`` (a :class:`HostTarget`: page-locked host memory) nothing result-sized is kept on the
    device: every k-chunk is transformed into one of two chunk buffers and copied into its column slice
    of the host array on a side stream while the next chunk is projected (the reference fills
    ``full_sed_data[:, k0:k1]`` chunk by chunk as well, sed_calculator.py:287-327).  Returns ``None``;
    the caller synchronises ``traj.engine.copy_stream`` before reading the array.  ``host_out.n_rows`` limits the
    streamed copy to the first rows: fftfreq order puts 0 <= f <= f_max first, so a frequency crop never
    leaves the device.
    """
    eng = traj.engine
    n_t, n_k = traj.n_t, int(k_vecs.shape[0])
    shape = (n_t, n_k, 3) if complex_out else (n_t, n_k)
    dtype = torch.complex64 if complex_out else torch.float32
    if complex_out:
        assert len(groups) == 1
    n_rows = n_t if host_out is None else max(0, min(int(host_out.n_rows), n_t))
    out = None if host_out is not None else torch.empty(shape, dtype=dtype, device=eng.device)
    if n_k == 0:
        return out
    eng.mark("start")
    mean, entries = traj.prepare(groups, use_displacements, stream=True)
    eng.mark("ingest")                      # positions + mean (+ velocities + digit planes unless they are streamed)
    arrivals = [traj.pop_arrivals(e[3]) for e in entries]
    ldp = (n_t + 3) // 4 * 4
    # Chunks whose projection runs frame range by frame range, in the order in which the digit planes become valid:
    # * one GPU, velocities still crossing PCIe (IngestStream shared by all groups): as many leading chunks as the
    #   tensor cores can project during the transfer, frames outermost;
    # * a multi-GPU exchange in flight on a side stream (one list of ranges per group): the first chunk.
    shared = arrivals[0] if arrivals and isinstance(arrivals[0], IngestStream) else None
    n_prefix = 0
    if shared is not None:
        assert all(a is shared for a in arrivals)
        chunks = plan_k_chunks(n_k, k_chunk)
        per_k = n_t * sum(e[1] for e in entries) / _STREAM_PROJECT_RATE
        k_budget, k_sum = 0.85 * shared.seconds / per_k, 0
        for _, nk in chunks:
            p_bytes = len(entries) * 2 * (k_sum + nk) * 3 * ldp * 4
            if n_prefix and (k_sum + nk / 2 > k_budget or p_bytes > _STREAM_P_BYTES):
                break
            n_prefix, k_sum = n_prefix + 1, k_sum + nk
    elif any(arrivals):
        longest = max(t1 - t0 for arr in arrivals if arr for t0, t1, _ in arr)
        chunks = plan_k_chunks(n_k, k_chunk, (longest, eng.n_clusters))
        n_prefix = 1
    else:
        chunks = plan_k_chunks(n_k, k_chunk)
    kc = max(nk for _, nk in chunks)
    k_prefix = sum(nk for _, nk in chunks[:n_prefix])
    window = traj.window
    rows_alloc = 2 * kc
    p_rows = 2 * max(kc, k_prefix)
    kv_dev = eng.upload_small(np.ascontiguousarray(k_vecs, np.float32))
    P = eng.empty((len(entries), p_rows, 3, ldp), torch.float32)
    group_stride = p_rows * 3 * ldp
    adig_bufs: Dict[int, torch.Tensor] = {}
    mode = _lib.MODE_COHERENT if complex_out else _lib.MODE_INCOHERENT
    if host_out is not None:
        elem = 24 if complex_out else 4                                  # bytes per (f, k)
        chunk_bufs = [torch.empty((n_t, kc) + shape[2:], dtype=dtype, device=eng.device) for _ in range(2)]
        drained = [None, None]                                           # copy-stream events per buffer
        compute, copy = torch.cuda.current_stream(eng.device), eng.copy_stream
    main = torch.cuda.current_stream(eng.device)
    if n_prefix:
        try:
            adig_pre = []
            for idx_dev, n_sel, pitch, dig, expo in entries:
                adig_pre.append(eng.phase_digits(kv_dev[:k_prefix], mean, idx_dev, n_sel, pitch, 2 * k_prefix))
            if shared is not None:
                for t0, t1, ev in shared:                     # iterating issues the next copy + digitise launch
                    main.wait_event(ev)
                    for g, (idx_dev, n_sel, pitch, dig, expo) in enumerate(entries):
                        eng.project(adig_pre[g], 2 * k_prefix, 2 * k_prefix, dig, expo, n_t, n_sel, pitch, P[g], ldp,
                                    t_range=(t0, t1))
            else:
                for g, (idx_dev, n_sel, pitch, dig, expo) in enumerate(entries):
                    for t0, t1, ev in arrivals[g] or [(0, n_t, None)]:   # frames in the order the peers' rows land
                        if ev is not None:
                            main.wait_event(ev)
                        if t1 > t0:
                            eng.project(adig_pre[g], 2 * k_prefix, 2 * k_prefix, dig, expo, n_t, n_sel, pitch, P[g], ldp,
                                        t_range=(t0, t1))
            del adig_pre
            eng.mark("prefix_projected")
        except BaseException:
            traj.reset_derived()                              # half-filled digit planes must not be reused
            raise
    for ci, (k0, nk) in enumerate(chunks):
        Pc = P
        if ci < n_prefix:
            Pc = P[:, 2 * k0:]                                # projected above; same group stride
        else:
            for g, (idx_dev, n_sel, pitch, dig, expo) in enumerate(entries):
                adig = adig_bufs.get(pitch)
                if adig is None:
                    adig = adig_bufs[pitch] = eng.empty((4, rows_alloc, pitch), torch.int8)
                eng.phase_digits(kv_dev[k0:k0 + nk], mean, idx_dev, n_sel, pitch, rows_alloc, out=adig)
                eng.project(adig, 2 * nk, rows_alloc, dig, expo, n_t, n_sel, pitch, P[g], ldp)
        if host_out is None:
            eng.fft_sed(Pc, len(entries), group_stride, nk, n_t, ldp, mode, out, n_k, k0, window)
            continue
        buf = chunk_bufs[ci & 1]
        if drained[ci & 1] is not None:
            compute.wait_event(drained[ci & 1])                          # its previous contents are on the host
        eng.fft_sed(Pc, len(entries), group_stride, nk, n_t, ldp, mode, buf, kc, 0, window)
        eng.mark("chunk_transformed")
        ready = torch.cuda.Event()
        ready.record(compute)
        copy.wait_event(ready)
        _lib.call("psa_copy_rows", host_out.ptr + (host_out.k_offset + k0) * elem, host_out.n_k_total * elem,
                  buf.data_ptr(), kc * elem, nk * elem, n_rows, copy.cuda_stream)
        drained[ci & 1] = torch.cuda.Event()
        drained[ci & 1].record(copy)
        eng.mark("chunk_on_host", copy)
    if host_out is not None:
        for b in chunk_bufs:                     # the copy stream still reads them: keep the allocator from reusing
            b.record_stream(copy)
    return out
