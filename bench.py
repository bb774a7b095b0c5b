#!/usr/bin/env python
"""Headline benchmark of the SED hot path (contract: see the task brief / DESIGN.md section 5).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1..c5] [--impl reference]

One *step* = one full pass of the hot path over one synthetic trajectory of the named BASELINE.json
config: mean positions -> digit planes -> (per k-chunk) phase table -> tensor-core projection -> FFT +
assembly.  The default workload is ``configs[3]`` (C4: the 100 x 100 k-grid on Si 12x12x12, 16384 frames) -
the largest configuration BASELINE.json names for 1/2/4/8 GPUs.  Units are (k-point, timestep, atom) triples/s.

* ``value``: inputs already resident in HBM.  N = 1: the raw float32 trajectory is on the device, the result stays
  there.  N > 1 (torchrun, one rank per GPU): STRONG scaling of the same job - the frames are resident in HBM spread
  over the ranks (rank r holds frames ``shard_range(n_t, r, N)``).  Every step: ordered float32 mean chain rank to
  rank; each rank digitises ITS frames and projects them for ALL k-points, the projection kernel's epilogue storing
  every tile through NVLink into the buffer of the rank that owns those k-points (the frames->k all-to-all fused into
  the tensor-core kernel, ``dist.frame_sharded_sed``); one one-element all-reduce per k-chunk as the fence; each
  rank then transforms its 1/N of the k-points.  The collectives inside the timed region are exactly those.
  (``PSA_B200_SHARD=k``: the previous scheme - digit planes all-gathered by the digitise kernel's peer stores,
  k-sharded projection.)
* ``e2e``: the public call on HOST (pinned) arrays - ``SEDCalculator.calculate`` (N = 1) or
  ``psa_b200.dist.calculate_sharded(ingest="sliced")`` (N > 1): H2D of positions + velocities (1/N per rank over its
  own PCIe link), the same kernels and exchange, D2H of every rank's spectra into one shared pinned host array.
* ``--impl reference`` times the reference's CPU algorithm (the NumPy oracle port, pinned bit for bit to the
  reference; the reference itself is pure Python and does not travel to the GPU box) with all host threads, on a
  bounded k-subset of the same workload per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path


def _host_cpus() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm must use all host cores, and BLAS reads these
# variables when NumPy is first imported - so this has to happen before `import numpy`.
if "reference" in sys.argv:
    for _var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_var] = str(_host_cpus())

import numpy as np  # noqa: E402

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "k-points·timesteps·atoms/sec"      # BASELINE.json's metric, verbatim
UNIT = "k*t*atom/s"
FLOP_PER_UNIT = 12.0          # 3 pol x (2 mul + 2 add): real series x complex phase
DEFAULT_WORKLOAD = "c4"
CPU_K = {"c1": 100, "c2": 200, "c3": 80, "c4": 48, "c5": 6}     # k-points per CPU-arm step (C1-C3: the full k-set)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner there) must not
# pollute it: keep a private handle on the real stdout and point fd 1 at stderr for everyone else.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line: dict) -> None:
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def config_of(args, cfg) -> dict:
    """The ``config`` object - identical in both arms (the driver compares them)."""
    return {"workload": args.workload, "desc": cfg["desc"]}


# ------------------------------------------------------------------------------------------------ workload
def build_jobs(cfg, calc):
    """[(k_mags, k_vecs, kwargs, chiral_pair|None)] for one step of the config."""
    jobs = []
    common = dict(basis_atom_types=cfg.get("basis_atom_types"), summation_mode=cfg["summation_mode"])
    if cfg["kind"] in ("kpath", "chiral"):
        for path in cfg["paths"]:
            mags, vecs = calc.get_k_path(path["direction"], cfg["bz_coverage"], path["n_k"])
            pair = (0, 1) if cfg["kind"] == "chiral" else None
            jobs.append((mags, vecs, dict(common), pair))
    else:
        kr = cfg["k_ranges"]
        mags, vecs, shape = calc.get_k_grid(cfg["plane"], kr[:2], kr[2:], cfg["n_kx"], cfg["n_ky"], cfg["k_fixed"])
        jobs.append((mags, vecs, dict(common), None))
    return jobs


def units_of(types, n_t, jobs_slices):
    """(k, t, atom) triples of a step: sum over jobs and projected groups of n_k * n_t * n_atoms_group."""
    from psa_b200 import groups as grp
    total = 0
    for (mags, vecs, kw, _), (k0, k1) in jobs_slices:
        g = grp.resolve_sed_groups(types, len(types), None, kw["basis_atom_types"], kw["summation_mode"])
        _, proj = grp.plan_sed_groups(g, kw["summation_mode"])
        total += (k1 - k0) * n_t * sum(int(p.size) for p in proj)
    return total


def pinned_frames(spec, f0, f1):
    """Frames [f0, f1) of the synthetic trajectory in pinned host memory."""
    import torch
    shape = (f1 - f0, spec.n_atoms, 3)
    pos = torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy()
    vel = torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy()
    spec.frames(f0, f1, out_pos=pos, out_vel=vel, threads=min(16, _host_cpus()))
    return pos, vel


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception as exc:  # nvidia-smi missing: report that instead of failing the bench
            log(f"[bench] clock sampling unavailable: {exc}")

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        inside = [r for ts, r in self.rows if t0 - 0.05 <= ts <= t1 + 0.15] or [r for _, r in self.rows]
        sm, reasons, sm_max, power = [], set(), None, []
        for r in inside:
            try:
                sm.append(float(r[0])); sm_max = float(r[1]); power.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": sm_max,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_sample(cfg, traj, max_k):
    """Oracle (NumPy port of the reference, all host cores) on a bounded k-subset; returns (units/s, desc, s)."""
    from oracle import psa_oracle as O
    from psa_b200 import kspace
    lat = kspace.Lattice.from_box(traj.box_matrix, *cfg["spec"].cells)
    if cfg["kind"] == "kgrid":
        kr = cfg["k_ranges"]
        _, vecs, _ = kspace.k_grid(cfg["plane"], kr[:2], kr[2:], cfg["n_kx"], cfg["n_ky"], cfg["k_fixed"])
    else:
        p = cfg["paths"][0]
        _, vecs = kspace.k_path(lat, p["direction"], cfg["bz_coverage"], p["n_k"])
    sel = np.linspace(0, len(vecs) - 1, min(max_k, len(vecs))).round().astype(int)
    kv = np.ascontiguousarray(vecs[sel])
    t0 = time.perf_counter()
    res = O.calculate(traj.positions, traj.velocities, traj.types, traj.dt_ps, kv,
                      basis_atom_types=cfg.get("basis_atom_types"), summation_mode=cfg["summation_mode"])
    sec = time.perf_counter() - t0
    n_atoms = sum(int(g.size) for g in res["groups"]) if not res["is_complex"] else \
        int(np.unique(np.concatenate(res["groups"])).size)
    units = len(kv) * traj.n_frames * n_atoms
    desc = (f"{len(kv)} of {len(vecs)} k-points of {cfg['spec'].name} ({traj.n_frames} frames x {n_atoms} atoms), "
            f"full trajectory, one calculate() incl. mean/phase/einsum/FFT")
    return units / sec, desc, sec


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [lib["num_threads"] for lib in threadpool_info() if lib.get("user_api") == "blas"]
        if n:
            return max(n)
    except Exception:
        pass
    return _host_cpus()


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    traj = cfg["spec"].trajectory(threads=min(16, _host_cpus()))
    max_k = args.cpu_k or CPU_K.get(args.workload, 48)
    for _ in range(args.warmup):
        cpu_sample(cfg, traj, max_k=min(8, max_k))
    rates, desc = [], ""
    t_begin = time.perf_counter()
    for _ in range(args.steps):
        val, desc, _sec = cpu_sample(cfg, traj, max_k=max_k)
        rates.append(val)
    ms = 1e3 * (time.perf_counter() - t_begin) / args.steps
    units_per_s = float(np.median(rates))
    cores = host_threads()
    line = {"impl": "reference", "metric": METRIC, "value": units_per_s, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(args, cfg),
            "cpu_baseline": {"value": units_per_s, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": units_per_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "details": {"inputs": "host memory", "omp_num_threads": os.environ.get("OMP_NUM_THREADS")},
            "gpu_launches": 0}
    emit(line)


def timeline_stages(marks, d2h_bytes):
    """Stage breakdown of one single-GPU public call from ``Engine.timeline`` marks (ms on the device's clock)."""
    def at(label, which):
        hits = [t for name, t in marks if name == label]
        return None if not hits else (hits[0] if which == "first" else hits[-1])
    out = {}
    ingest, last_range = at("ingest", "first"), at("range_on_device", "last")
    prefix, last_fft, first_fft = at("prefix_projected", "last"), at("chunk_transformed", "last"), at("chunk_transformed", "first")
    last_host = at("chunk_on_host", "last")
    if ingest is not None:
        out["positions_upload_mean" + ("" if last_range is not None else "_velocities_digitize")] = ingest
    if last_range is not None:
        out["velocity_ranges_upload_digitize"] = last_range - ingest
        if prefix is not None:
            out["streamed_chunks_projection_tail"] = prefix - last_range
    if last_fft is not None:
        out["compute_rest"] = last_fft - (prefix if prefix is not None else ingest)
    if last_host is not None and last_fft is not None:
        out["drain_tail"] = last_host - last_fft
        if first_fft is not None and last_host > first_fft:
            out["d2h_GBps_while_draining"] = d2h_bytes / (last_host - first_fft) / 1e6
    return out


# ------------------------------------------------------------------------------------------------ roofline
def peaks():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        p = json.loads(path.read_text())
        return (p["hbm_gbs"], p["bf16_tflops"], p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "measured (MEASURED_PEAKS.json)")
    return 6650.0, 1590.0, 1400.0, "fallback (B200_PROFILING.md)"


def measure_int8_peak(dev):
    """Dense int8 tensor throughput of this GPU with the driver's own methodology for bf16 (torch, 8192^3: best of 10
    = burst, back to back for 4 s = sustained), through cuBLASLt's int8 GEMM.  None when the library path is missing."""
    import torch
    try:
        n = 8192
        a = torch.randint(-100, 100, (n, n), dtype=torch.int8, device=dev)
        b = torch.randint(-100, 100, (n, n), dtype=torch.int8, device=dev)
        for _ in range(3):
            torch._int_mm(a, b)
        torch.cuda.synchronize(dev)
        best = float("inf")
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._int_mm(a, b)
            e1.record()
            torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        ops = 2.0 * n ** 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_end, reps = time.time() + 4.0, 0
        e0.record()
        while time.time() < t_end:
            for _ in range(20):
                torch._int_mm(a, b)
            reps += 20
            torch.cuda.synchronize(dev)
        e1.record()
        torch.cuda.synchronize(dev)
        return {"int8_tops": ops / (best / 1e3) / 1e12,
                "int8_tops_sustained": ops * reps / (e0.elapsed_time(e1) / 1e3) / 1e12,
                "how": "torch._int_mm int8 8192^3 (2*N^3): best of 10 (burst) and back to back for 4 s (sustained)"}
    except Exception as exc:
        log(f"[bench] int8 peak not measured: {exc}")
        return None


def roofline_for(kernel, ms_total, calls, work, traffic, int8_peak=None, long_step=False):
    """``work`` = algorithmic bytes or flops per step for that kernel."""
    hbm, tf_burst, tf_sust, which = peaks()
    per_launch_s = ms_total / 1e3 / max(calls, 1)
    if kernel == "psa_project":
        tf = tf_sust if long_step else tf_burst
        ach = work / max(calls, 1) / per_launch_s / 1e12
        if int8_peak:
            ipeak = int8_peak["int8_tops_sustained" if long_step else "int8_tops"]
            isrc = "measured in this run: " + int8_peak["how"]
        else:
            ipeak, isrc = 2.0 * tf, "ASSUMED 2 x the measured dense bf16 peak (int8 GEMM not available to measure)"
        return {"kernel": kernel, "bound": "tensor", "achieved": ach, "peak": tf, "unit": "TFLOP/s",
                "frac": ach / tf, "traffic": traffic,
                "peak_source": which + (" sustained" if long_step else " burst"),
                # the same launch counted in the int8 operations the tensor cores actually execute
                "executed": {"achieved": 10.0 * ach, "peak": ipeak, "unit": "int8 TOP/s", "frac": 10.0 * ach / ipeak,
                             "peak_source": isrc},
                "note": "algorithmic 12 flop/unit; executed: 10 exact int8 digit products per MAC (120 int8-op/unit), "
                        "so the algorithmic frac is <= int8_peak/(10*bf16_peak) by construction; ncu tensor-pipe "
                        "utilisation in profiles/"}
    ach = work / max(calls, 1) / per_launch_s / 1e9
    return {"kernel": kernel, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
            "traffic": traffic, "peak_source": which}


def ncu_traffic(kernel, workload):
    path = ROOT / "profiles" / "ncu_traffic.json"
    if path.exists():
        try:
            return json.loads(path.read_text()).get(workload, {}).get(kernel)
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="psa_b200", choices=["psa_b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("PSA_BENCH_WORKLOAD", DEFAULT_WORKLOAD))
    ap.add_argument("--frames", type=int, default=None, help="override n_frames (smoke runs only)")
    ap.add_argument("--cells", type=int, default=None, help="override the supercell size (smoke runs only)")
    ap.add_argument("--cpu-k", type=int, default=None, help="k-points in one CPU-arm step (default: per workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ised", action="store_true")
    ap.add_argument("--no-int8-peak", action="store_true")
    args = ap.parse_args()
    args.workload = args.workload.lower()

    import synthetic as synth
    cfg = synth.baseline_config(args.workload, n_frames=args.frames, n_cells=args.cells)
    spec = cfg["spec"]
    cfg["desc"] = (f"{spec.name}: {spec.n_atoms} atoms x {spec.n_frames} frames, {cfg['kind']}, "
                   f"{cfg['summation_mode']}, basis_atom_types={cfg.get('basis_atom_types')}")
    if args.impl == "reference":
        return run_reference(args, cfg)

    import torch
    import torch.distributed as dist
    from psa_b200 import SEDCalculator
    from psa_b200 import dist as pdist
    from psa_b200 import groups as grp

    rank, world, local = pdist.init_from_env()
    if world != args.gpus:
        log(f"[bench] --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    dev = torch.device("cuda", torch.cuda.current_device())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- inputs.  N = 1: the whole trajectory (pinned host + device).  N > 1: this rank's frames only.
    t_gen = time.time()
    n_t, n_a = spec.n_frames, spec.n_atoms
    f0, f1 = pdist.shard_range(n_t, rank, world)
    if world > 1 and f0 % synth.BLOCK:
        raise SystemExit(f"[bench] {n_t} frames do not split over {world} ranks on {synth.BLOCK}-frame blocks")
    pos_rows, vel_rows = pinned_frames(spec, f0, f1)
    if world == 1:
        traj = spec.wrap(pos_rows, vel_rows)
    else:           # shape-only placeholder: the sharded calls take this rank's rows explicitly
        zero = np.broadcast_to(np.zeros(1, np.float32), (n_t, n_a, 3))
        traj = spec.wrap(zero, zero)
    log(f"[bench r{rank}] generated frames [{f0}, {f1}) of {cfg['desc']} in {time.time() - t_gen:.1f}s")
    calc = SEDCalculator(traj, *spec.cells, device=dev.index)
    eng = calc.engine
    jobs = build_jobs(cfg, calc)
    slices = [pdist.shard_range(len(j[1]), rank, world) for j in jobs]
    rows_dev = None
    if world > 1:
        rows_dev = (torch.from_numpy(pos_rows).to(dev), torch.from_numpy(vel_rows).to(dev))     # resident, untimed

    def proj_groups_of(kw):
        g = grp.resolve_sed_groups(traj.types, n_a, None, kw["basis_atom_types"], kw["summation_mode"])
        return grp.plan_sed_groups(g, kw["summation_mode"])

    def step_resident():
        """Hot path from the device-resident raw frames to the device-resident result (this rank's k-slice)."""
        calc.device_trajectory.reset_derived()
        outs = []
        for (mags, vecs, kw, pair), (k0, k1) in zip(jobs, slices):
            if world > 1:
                cplx_, groups_ = proj_groups_of(kw)
                out = pdist.sharded_sed_on_device(calc, np.ascontiguousarray(vecs, np.float32).reshape(-1, 3), groups_,
                                                  cplx_, rows_dev)
            else:
                out, _, _ = calc._calculate_device(vecs[k0:k1], None, kw["basis_atom_types"], kw["summation_mode"])
            if pair is not None:
                outs.append(calc._chiral_phase_of_result(out, pair))
            outs.append(out)
        return outs

    # ---------------- value: inputs resident in HBM
    if world == 1:
        _ = calc.device_trajectory.positions, calc.device_trajectory.velocities     # upload once, untimed
    for _ in range(args.warmup):
        step_resident()
    barrier()
    launches0 = eng.launches
    sampler = ClockSampler(dev.index)
    time.sleep(0.25)
    t0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    marks = []
    for _ in range(args.steps):
        step_resident()
        marks.append(torch.cuda.Event(enable_timing=True))
        marks[-1].record()
    ev1.record()
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    ms_per_step = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    per_step = [a.elapsed_time(b) for a, b in zip([ev0] + marks[:-1], marks)]     # this rank; shows the power-cap drift
    launches = eng.launches - launches0                         # this rank's kernels inside the timed region

    jobs_full = [(j, (0, len(j[1]))) for j in jobs]
    units_total = float(units_of(traj.types, n_t, jobs_full))   # the whole job, whatever N
    units_local = float(units_of(traj.types, n_t, list(zip(jobs, slices))))
    value = units_total / (ms_per_step / 1e3)

    # ---------------- per-kernel breakdown (one extra step, events around every C-ABI call)
    eng.profile = {}
    step_resident()
    prof = eng.profile_summary()
    eng.profile = None
    total_prof = sum(v["ms"] for v in prof.values()) or 1.0
    n_k_local = sum(k1 - k0 for k0, k1 in slices)
    cplx, proj_groups = proj_groups_of(jobs[0][2])
    n_sel_sum = sum(int(p.size) for p in proj_groups)
    rows_local = (f1 - f0) if world > 1 else n_t
    frames_path = world > 1 and pdist.last_path == "frames"
    work = {
        "psa_project": FLOP_PER_UNIT * units_local,                                           # flops
        "psa_mean_positions": 12.0 * n_t * n_a,                                               # bytes
        "psa_mean_accumulate": 12.0 * rows_local * n_a,
        "psa_digitize": 24.0 * rows_local * n_sel_sum,
        "psa_digitize_rows": 24.0 * rows_local * n_sel_sum,
        "psa_digitize_rows_peers": (12.0 + 12.0 * world) * rows_local * n_sel_sum,           # read once, stored on N ranks
        # frame-sharded: every rank builds the phase table of ALL k-points (for its own frames)
        "psa_phase_digits": 8.0 * (units_total if frames_path else units_local) / max(n_t, 1),
        "psa_fft_sed": (48.0 if cplx else 24.0 * len(proj_groups) + 4.0) * n_k_local * n_t,
        "psa_chiral_phase": 20.0 * n_k_local * n_t,
    }
    long_step = ms_per_step > 20.0          # a step of tens of ms runs under the power cap: sustained peaks apply
    int8_peak = None if (args.no_int8_peak or rank != 0) else measure_int8_peak(dev)
    kernels = {k: {"ms": v["ms"], "share": v["ms"] / total_prof, "calls": v["calls"]} for k, v in prof.items()}
    dominant = max(prof, key=lambda k: prof[k]["ms"])
    rooflines = {k: roofline_for(k, prof[k]["ms"], prof[k]["launches"], work[k], ncu_traffic(k, args.workload),
                                 int8_peak, long_step) for k in prof if k in work}
    roof = rooflines.get(dominant) or roofline_for(dominant, prof[dominant]["ms"], prof[dominant]["launches"], 0.0, None)

    # ---------------- e2e: public API on host (pinned) arrays, H2D + compute + D2H inside the timed region
    e2e, parity = None, None
    if not args.no_e2e:
        d2h = [0]
        stage_ms = {}
        last = {}

        def step_e2e(timings=None):
            calc.release_device_memory()
            res_bytes = 0
            for (mags, vecs, kw, pair) in jobs:
                if world == 1:
                    if pair is not None:
                        res = calc.calculate_chiral_sed(cfg["paths"][0]["direction"], cfg["bz_coverage"],
                                                        len(mags), chiral_axis=cfg.get("chiral_axis", "z"))
                    else:
                        res = calc.calculate(mags, vecs, **kw)
                else:
                    res = pdist.calculate_sharded(calc, mags, vecs, ingest="sliced", local_rows=(pos_rows, vel_rows),
                                                  timings=timings, **kw)
                if res is not None:
                    res_bytes += res.sed.nbytes + (res.phase.nbytes if res.phase is not None else 0)
                    last["res"] = res
            d2h[0] = res_bytes

        n_e2e = max(1, min(args.steps, 5))
        for _ in range(min(args.warmup, 2)):
            step_e2e()
        barrier()
        w0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_e2e):
            step_e2e()
        e1.record()
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - w0) / n_e2e
        e2e_ms = max_over_ranks(max(e0.elapsed_time(e1) / n_e2e, wall_ms))
        if world > 1:                                           # one extra, untimed call for the stage breakdown
            step_e2e(timings=stage_ms)
            barrier()
        else:
            eng.timeline = []
            step_e2e()
            marks = eng.timeline_ms()
            eng.timeline = None
            stage_ms.update(timeline_stages(marks, d2h[0]))
        e2e = {"value": units_total / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms,
               "steps": n_e2e, "h2d_bytes_per_step": int(2 * n_t * n_a * 12), "d2h_bytes_per_step": int(d2h[0]),
               "stage_ms_rank0": {k: round(v, 3) for k, v in stage_ms.items()} or None,
               "path": "SEDCalculator.calculate on pinned host arrays (velocity ranges copied + digitised on a side stream "
                       "while the leading k-chunks are projected range by range; k-chunks leave for the pinned result "
                       "while the next one is computed)" if world == 1 else
                       "psa_b200.dist.calculate_sharded(ingest='sliced'): every rank uploads 1/N of the frames over its own "
                       "PCIe link, float32 mean chain, " +
                       ("each rank projects its frames for all k-points with the tiles stored into the owner ranks through "
                        "NVLink by the projection kernel, owners transform their k-slice, "
                        if pdist.last_path == "frames" else
                        "digit planes stored into every rank through NVLink by the digitise kernel, k-sharded compute, ") +
                       "every rank's spectra copied into one shared pinned host array"}

        # ---------------- multi-GPU parity: a few k columns of the sharded result against a single-GPU compute, bitwise
        if world > 1:
            ok = 1
            if rank == 0:
                res = last["res"]
                mags, vecs, kw, pair = jobs[-1]
                n_k = len(vecs)
                per = [pdist.shard_range(n_k, r, world) for r in range(world)]
                cols = sorted({a for a, b in per if b > a} | {b - 1 for a, b in per if b > a}
                              | {(a + b) // 2 for a, b in per if b > a})
                t_par = time.time()
                full = spec.trajectory(threads=min(16, _host_cpus()))
                single = SEDCalculator(full, *spec.cells, device=dev.index)
                one = single.calculate(np.zeros(len(cols), np.float32), vecs[cols], **kw)
                ok = int(np.array_equal(one.sed, res.sed[:, cols]))
                single.release_device_memory()
                parity = {"parity_checked": bool(ok), "columns": len(cols), "seconds": round(time.time() - t_par, 1),
                          "how": "k columns from every rank's slice of the sharded e2e result vs SEDCalculator.calculate on "
                                 "one GPU with the whole trajectory: numpy.array_equal (bitwise)"}
                del full, one
            flag = torch.tensor([ok], device=dev)
            dist.broadcast(flag, src=0)
        last.clear()

    # ---------------- batched iSED: 64 (k, omega) points x 100 frames on this trajectory (C5 names it; measured on
    # every workload at N = 1 so that the driver's record carries the kernel)
    ised = None
    if not args.no_ised and world == 1:
        n_pts, n_fr = int(cfg.get("ised_points", 64)), int(cfg.get("ised_frames", 100))
        side = max(1, int(round(n_pts ** 0.5)))
        a_lat = float(np.linalg.norm(calc.a1))
        k_max = 2.0 * np.pi / a_lat
        targets = [(k_max * (i + 1) / (side + 1), 1.0 + 14.0 * j / max(1, side - 1)) for i in range(side) for j in range(side)]
        kw = dict(nk_on_path=100, bz_cov_ised=1.0, n_recon_frames=n_fr)
        calc.reconstruct([1, 0, 0], targets[:2], a_lat, keep_on_device=True, **kw)      # warm-up (ingest, plans)
        eng.profile = {}
        torch.cuda.synchronize(dev)
        t_i = time.perf_counter()
        res = calc.reconstruct([1, 0, 0], targets, a_lat, keep_on_device=True, rescale_factor="auto", **kw)
        torch.cuda.synchronize(dev)
        sec_dev = time.perf_counter() - t_i
        iprof = eng.profile_summary()
        eng.profile = None
        ok = bool(all(torch.isfinite(r["frames"]).all().item() for r in res[:2]))
        del res
        out_bytes = 12.0 * len(targets) * n_fr * n_a
        k_ms = iprof.get("psa_ised_frames", {}).get("ms", 0.0)
        ised = {"points": len(targets), "frames": n_fr, "atoms": int(n_a), "frames_GB": out_bytes / 1e9,
                "seconds_whole_call_device_resident": sec_dev, "kernel_ms": {k: v["ms"] for k, v in iprof.items()},
                "path": "SEDCalculator.reconstruct(rescale_factor='auto'): amplitudes from one projection pass per atom "
                        "group, one batched max pass + one batched synthesis pass over all points", "checked": ok}
        if k_ms > 0:
            rooflines["psa_ised_frames"] = roofline_for("psa_ised_frames", k_ms, iprof["psa_ised_frames"]["launches"],
                                                        out_bytes, None)

    # ---------------- CPU baseline (rank 0, N == 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        val, desc, sec = cpu_sample(cfg, traj, max_k=args.cpu_k or CPU_K.get(args.workload, 48))
        cpu = {"value": val, "unit": UNIT, "cores": host_threads(), "kind": "port", "sample": desc, "seconds": sec}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "s8 digit planes -> s32 (exact), f64 FFT", "data": "synthetic",
            "config": config_of(args, cfg),
            "details": {"k_points_total": int(sum(len(j[1]) for j in jobs)), "k_points_this_rank": int(n_k_local),
                        "frames_resident_this_rank": int(rows_local),
                        "l2_policy": "inputs larger than L2 (trajectory %.0f MB per array%s)"
                                     % (n_t * n_a * 12 / 1e6, "" if world == 1 else f", 1/{world} per rank"),
                        "step": "mean positions + digit planes + phase table + tcgen05 projection + FFT/assembly",
                        "step_ms_first_median_last": [round(per_step[0], 3), round(float(np.median(per_step)), 3),
                                                      round(per_step[-1], 3)],
                        "multi_gpu_path": None if world == 1 else pdist.last_path,
                        "collectives_in_value": None if world == 1 else
                        ("N-1 send/recv hops + 1 broadcast of the (n_atoms, 3) running mean; every rank projects its own "
                         "frames for all k-points and the projection kernel stores each tile into the owner rank's buffer "
                         "through NVLink (frames->k all-to-all fused into the tcgen05 kernel); 1-element all-reduces as "
                         "fences (one at the start, one per k-chunk on a side stream); no digit plane leaves its GPU"
                         if pdist.last_path == "frames" else
                         "N-1 send/recv hops + 1 broadcast of the (n_atoms, 3) running mean; digit planes exchanged by the "
                         "digitise kernel's peer stores over NVLink as a ring of N-1 steps on a side stream, each fenced by a "
                         "1-element all-reduce, running under the first k-chunk's projection (which follows the arrival "
                         "order of the frame ranges); nothing else during compute")},
            "roofline": roof, "rooflines": rooflines, "kernels": kernels, "cpu_baseline": cpu, "e2e": e2e, "ised": ised,
            "int8_peak": int8_peak,
            "gpu_launches": int(launches), "gpu_launches_per_step": int(launches) // max(1, args.steps), "clocks": clocks,
        }
        if parity is not None:
            line.update(parity_checked=parity["parity_checked"], parity=parity)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
