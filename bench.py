#!/usr/bin/env python
"""Headline benchmark of the SED hot path (contract: see the task brief / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1..c5] [--impl reference]

One *step* = one full pass of the hot path over one synthetic trajectory of the named BASELINE.json
config: mean positions -> digit planes -> (per k-chunk) phase table -> tensor-core projection -> FFT +
assembly.  ``value`` times that with the raw float32 trajectory already resident in HBM; ``e2e`` times
the public ``SEDCalculator.calculate`` call on HOST (pinned) arrays: H2D of positions + velocities,
the same kernels, D2H of the result.  Units are (k-point, timestep, atom) triples per second.

N > 1 (launched by torchrun, one rank per GPU): ``value`` is weak scaling - every rank holds the
trajectory and projects its own n_k k-points of an N x n_k path, no collective in the data path;
``e2e`` is the real multi-GPU call (rank 0 uploads and ingests, one NCCL broadcast of the digit
planes, k-sharded compute, gather on rank 0, D2H).

``--impl reference`` times the reference's CPU algorithm (the NumPy oracle port - the reference is
pure Python and cannot travel to the GPU box) on all host cores, on a bounded k-subset of the workload.
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

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "k-points\u00b7timesteps\u00b7atoms/sec"      # BASELINE.json's metric, verbatim
UNIT = "k*t*atom/s"
FLOP_PER_UNIT = 12.0          # 3 pol x (2 mul + 2 add): real series x complex phase


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner there) must not
# pollute it: keep a private handle on the real stdout and point fd 1 at stderr for everyone else.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line: dict) -> None:
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


# ------------------------------------------------------------------------------------------------ workload
def build_jobs(cfg, calc, k_mult=1):
    """[(k_mags, k_vecs, kwargs, chiral_pair|None)] for one step of the config; ``k_mult`` densifies
    the k-set (weak scaling over ranks)."""
    jobs = []
    common = dict(basis_atom_types=cfg.get("basis_atom_types"), summation_mode=cfg["summation_mode"])
    if cfg["kind"] in ("kpath", "chiral"):
        for path in cfg["paths"]:
            mags, vecs = calc.get_k_path(path["direction"], cfg["bz_coverage"], path["n_k"] * k_mult)
            pair = (0, 1) if cfg["kind"] == "chiral" else None
            jobs.append((mags, vecs, dict(common), pair))
    else:
        kr = cfg["k_ranges"]
        mags, vecs, shape = calc.get_k_grid(cfg["plane"], kr[:2], kr[2:], cfg["n_kx"] * k_mult, cfg["n_ky"], cfg["k_fixed"])
        jobs.append((mags, vecs, dict(common), None))
    return jobs


def units_of(cfg, types, n_t, jobs_slices):
    """(k, t, atom) triples of a step: sum over jobs and projected groups of n_k * n_t * n_atoms_group."""
    from psa_b200 import groups as grp
    total = 0
    for (mags, vecs, kw, _), (k0, k1) in jobs_slices:
        g = grp.resolve_sed_groups(types, len(types), None, kw["basis_atom_types"], kw["summation_mode"])
        _, proj = grp.plan_sed_groups(g, kw["summation_mode"])
        total += (k1 - k0) * n_t * sum(int(p.size) for p in proj)
    return total


def pinned_trajectory(spec):
    import torch
    shape = (spec.n_frames, spec.n_atoms, 3)
    pos = torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy()
    vel = torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy()
    spec.frames(0, spec.n_frames, out_pos=pos, out_vel=vel, threads=min(16, os.cpu_count() or 8))
    return spec.wrap(pos, vel)


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
def cpu_sample(cfg, traj, max_k=48, repeats=1):
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
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        res = O.calculate(traj.positions, traj.velocities, traj.types, traj.dt_ps, kv,
                          basis_atom_types=cfg.get("basis_atom_types"), summation_mode=cfg["summation_mode"])
        best = min(best, time.perf_counter() - t0)
    n_atoms = sum(int(g.size) for g in res["groups"]) if not res["is_complex"] else \
        int(np.unique(np.concatenate(res["groups"])).size)
    units = len(kv) * traj.n_frames * n_atoms
    desc = (f"{len(kv)} of {len(vecs)} k-points of {cfg['spec'].name} ({traj.n_frames} frames x {n_atoms} atoms), "
            f"full trajectory, one calculate() incl. mean/phase/einsum/FFT")
    return units / best, desc, best


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [lib["num_threads"] for lib in threadpool_info() if lib.get("user_api") == "blas"]
        if n:
            return max(n)
    except Exception:
        pass
    return os.cpu_count() or 1


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    traj = cfg["spec"].trajectory(threads=min(16, os.cpu_count() or 8))
    for _ in range(args.warmup):
        cpu_sample(cfg, traj, max_k=8)
    rates, desc = [], ""
    t_begin = time.perf_counter()
    for _ in range(args.steps):
        val, desc, _sec = cpu_sample(cfg, traj, max_k=args.cpu_k)
        rates.append(val)
    ms = 1e3 * (time.perf_counter() - t_begin) / args.steps
    units_per_s = float(np.median(rates))
    cores = host_threads()
    line = {"impl": "reference", "metric": METRIC, "value": units_per_s, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "desc": cfg["desc"], "inputs": "host memory"},
            "cpu_baseline": {"value": units_per_s, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": units_per_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------ roofline
def peaks():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        p = json.loads(path.read_text())
        return p["hbm_gbs"], p["bf16_tflops"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def roofline_for(kernel, ms_total, calls, work, traffic):
    """``work`` = algorithmic bytes or flops per step for that kernel."""
    hbm, tf, which = peaks()
    per_launch_s = ms_total / 1e3 / max(calls, 1)
    if kernel == "psa_project":
        ach = work / max(calls, 1) / per_launch_s / 1e12
        return {"kernel": kernel, "bound": "tensor", "achieved": ach, "peak": tf, "unit": "TFLOP/s",
                "frac": ach / tf, "traffic": traffic, "peak_source": which,
                # the same launch counted in the int8 operations the tensor cores actually execute
                "executed": {"achieved": 10.0 * ach, "peak": 2.0 * tf, "unit": "int8 TOP/s", "frac": 10.0 * ach / (2.0 * tf),
                             "peak_source": "2 x the measured dense bf16 peak (int8 runs at twice the bf16 rate)"},
                "note": "algorithmic 12 flop/unit; executed: 10 int8 digit products per MAC (120 int8-op/unit), "
                        "so frac <= 2*bf16_peak/10 by construction; see ncu tensor-pipe utilisation in profiles/"}
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
    ap.add_argument("--workload", default=os.environ.get("PSA_BENCH_WORKLOAD", "c2"))
    ap.add_argument("--frames", type=int, default=None, help="override n_frames (smoke runs only)")
    ap.add_argument("--cpu-k", type=int, default=100, help="k-points in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    from psa_b200 import synth
    cfg = synth.baseline_config(args.workload, n_frames=args.frames)
    spec = cfg["spec"]
    cfg["desc"] = (f"{spec.name}: {spec.n_atoms} atoms x {spec.n_frames} frames, {cfg['kind']}, "
                   f"{cfg['summation_mode']}, basis_atom_types={cfg.get('basis_atom_types')}")
    if args.impl == "reference":
        return run_reference(args, cfg)

    import torch
    import torch.distributed as dist
    from psa_b200 import SEDCalculator
    from psa_b200 import dist as pdist

    rank, world, local = pdist.init_from_env()
    if world != args.gpus:
        log(f"[bench] --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    dev = torch.device("cuda", torch.cuda.current_device())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # Every rank needs the raw trajectory resident in HBM for the weak-scaling `value`.  Small workloads are
    # generated by every rank (same seed); large ones are generated once on rank 0 and copied GPU-to-GPU
    # (setup, untimed) so that host memory holds one copy only.
    t_gen = time.time()
    traj_bytes = 2 * spec.n_frames * spec.n_atoms * 12
    share = world > 1 and traj_bytes > (4 << 30)
    resident = None
    if rank == 0 or not share:
        traj = pinned_trajectory(spec)
    else:
        zero = np.broadcast_to(np.zeros(1, np.float32), (spec.n_frames, spec.n_atoms, 3))
        traj = spec.wrap(zero, zero)
    log(f"[bench r{rank}] generated {cfg['desc']} in {time.time() - t_gen:.1f}s (shared={share})")
    calc = SEDCalculator(traj, *spec.cells, device=dev.index)
    eng = calc.engine
    if share:
        from psa_b200.engine import DeviceTrajectory
        shape = (spec.n_frames, spec.n_atoms, 3)
        resident = []
        for arr in (traj.positions, traj.velocities):
            t = torch.empty(shape, dtype=torch.float32, device=dev)
            if rank == 0:
                t.copy_(torch.from_numpy(arr), non_blocking=True)
            dist.broadcast(t, src=0)
            resident.append(t)
        calc._dev_traj = DeviceTrajectory(eng, resident[0], resident[1])
    jobs = build_jobs(cfg, calc, k_mult=world)
    slices = [pdist.shard_range(len(j[1]), rank, world) for j in jobs]

    def step_resident():
        """Hot path from the raw device-resident trajectory to the device-resident result."""
        calc.device_trajectory.reset_derived()
        outs = []
        for (mags, vecs, kw, pair), (k0, k1) in zip(jobs, slices):
            out, _, _ = calc._calculate_device(vecs[k0:k1], None, kw["basis_atom_types"], kw["summation_mode"])
            if pair is not None:
                outs.append(calc._chiral_phase_of_result(out, pair))
            outs.append(out)
        return outs

    # ---------------- value: inputs resident in HBM
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
    for _ in range(args.steps):
        step_resident()
    ev1.record()
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    launches = eng.launches - launches0                         # this rank's kernels inside the timed region

    units_local = units_of(cfg, traj.types, traj.n_frames, list(zip(jobs, slices)))
    units_t = torch.tensor([float(units_local)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(units_t, op=dist.ReduceOp.SUM)
    units = float(units_t.item())
    value = units / (ms_per_step / 1e3)

    # ---------------- per-kernel breakdown (one extra step, events around every C-ABI call)
    eng.profile = {}
    step_resident()
    prof = eng.profile_summary()
    eng.profile = None
    total_prof = sum(v["ms"] for v in prof.values()) or 1.0
    n_t, n_all = traj.n_frames, traj.n_atoms
    n_sel_units = units_local / max(n_t, 1)                      # sum over jobs/groups of n_k * n_atoms_group
    n_k_local = sum(k1 - k0 for k0, k1 in slices)
    from psa_b200 import groups as grp
    g0 = grp.resolve_sed_groups(traj.types, n_all, None, cfg.get("basis_atom_types"), cfg["summation_mode"])
    cplx, proj_groups = grp.plan_sed_groups(g0, cfg["summation_mode"])
    n_sel_sum = sum(int(p.size) for p in proj_groups)
    work = {
        "psa_project": FLOP_PER_UNIT * units_local,                                           # flops
        "psa_mean_positions": 12.0 * n_t * n_all,                                             # bytes
        "psa_digitize": 24.0 * n_t * n_sel_sum,
        "psa_phase_digits": 8.0 * n_sel_units / 1.0,
        "psa_fft_sed": (48.0 if cplx else 24.0 * len(proj_groups) + 4.0) * n_k_local * n_t,
        "psa_chiral_phase": 20.0 * n_k_local * n_t,
    }
    kernels = {k: {"ms": v["ms"], "share": v["ms"] / total_prof, "calls": v["calls"]} for k, v in prof.items()}
    dominant = max(prof, key=lambda k: prof[k]["ms"])
    roof = roofline_for(dominant, prof[dominant]["ms"], prof[dominant]["launches"], work.get(dominant, 0.0),
                        ncu_traffic(dominant, args.workload))
    rooflines = {k: roofline_for(k, prof[k]["ms"], prof[k]["launches"], work[k], ncu_traffic(k, args.workload))
                 for k in prof if k in work}

    # ---------------- e2e: public API on host (pinned) arrays, H2D + compute + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        d2h = [0]
        # N > 1: every rank uploads its own 1/N of the frames (pinned) over its own PCIe link
        ingest, local_rows = "broadcast", None
        if world > 1:
            f0, f1 = pdist.shard_range(spec.n_frames, rank, world)
            if not share or rank == 0:
                ingest = "sliced"
                local_rows = (traj.positions[f0:f1], traj.velocities[f0:f1])
            elif f0 % 256 == 0:                              # synthetic frames are generated in blocks of 256
                ingest = "sliced"
                rows_shape = (f1 - f0, spec.n_atoms, 3)
                lp = torch.empty(rows_shape, dtype=torch.float32, pin_memory=True).numpy()
                lv = torch.empty(rows_shape, dtype=torch.float32, pin_memory=True).numpy()
                spec.frames(f0, f1, out_pos=lp, out_vel=lv, threads=min(16, os.cpu_count() or 8))
                local_rows = (lp, lv)
            flag = torch.tensor([1 if ingest == "sliced" else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                ingest, local_rows = "broadcast", None

        def step_e2e():
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
                    res = pdist.calculate_sharded(calc, mags, vecs, ingest=ingest, local_rows=local_rows, **kw)
                if res is not None:
                    res_bytes += res.sed.nbytes + (res.phase.nbytes if res.phase is not None else 0)
            d2h[0] = res_bytes

        n_e2e = max(1, min(args.steps, 5))
        for _ in range(min(args.warmup, 2)):
            step_e2e()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for _ in range(n_e2e):
            step_e2e()
        e1.record()
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - w0) / n_e2e
        ems = torch.tensor([max(e0.elapsed_time(e1) / n_e2e, wall_ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        jobs_full = [(j, (0, len(j[1]))) for j in jobs]
        units_e2e = units_of(cfg, traj.types, traj.n_frames, jobs_full)
        e2e = {"value": units_e2e / (float(ems.item()) / 1e3), "unit": UNIT, "ms_per_step": float(ems.item()),
               "steps": n_e2e, "h2d_bytes_per_step": int(2 * traj.positions.nbytes),
               "d2h_bytes_per_step": int(d2h[0]),
               "path": "SEDCalculator.calculate on pinned host arrays" if world == 1 else
                       ("psa_b200.dist.calculate_sharded(ingest='sliced'): every rank uploads 1/N of the frames, running-sum "
                        "mean chain + all-gather of the digit planes, k-sharded compute, gather, D2H" if ingest == "sliced" else
                        "psa_b200.dist.calculate_sharded: rank-0 upload+ingest, NCCL broadcast, k-sharded compute, gather, D2H")}

    # ---------------- batched iSED (configs that name it: C5's 64 (k, omega) points, split over the ranks)
    ised = None
    if cfg.get("ised_points"):
        n_pts, n_fr = int(cfg["ised_points"]), int(cfg["ised_frames"])
        side = max(1, int(round(n_pts ** 0.5)))
        a_lat = float(np.linalg.norm(calc.a1))
        k_max = 2.0 * np.pi / a_lat
        targets = [(k_max * (i + 1) / (side + 1), 1.0 + 14.0 * j / max(1, side - 1)) for i in range(side) for j in range(side)]
        mine = targets[rank::world]
        kw = dict(nk_on_path=100, bz_cov_ised=1.0, n_recon_frames=n_fr)
        calc.reconstruct([1, 0, 0], mine[:1], a_lat, **kw)                      # warm-up

        def timed(**extra):
            barrier()
            t0 = time.perf_counter()
            out = calc.reconstruct([1, 0, 0], mine, a_lat, **kw, **extra)
            torch.cuda.synchronize(dev)
            sec = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(sec, op=dist.ReduceOp.MAX)
            return out, float(sec.item())

        res, sec_dev = timed(keep_on_device=True)
        ok = bool(all(torch.isfinite(r["frames"]).all().item() for r in res[:2]))
        del res
        res, sec_host = timed()
        out_bytes = 12.0 * len(targets) * n_fr * traj.n_atoms
        ised = {"points": len(targets), "frames": n_fr, "atoms": int(traj.n_atoms), "frames_GB": out_bytes / 1e9,
                "seconds_device_resident": sec_dev, "GB_per_s_device_resident": out_bytes / 1e9 / sec_dev,
                "seconds_to_host": sec_host, "GB_per_s_to_host": out_bytes / 1e9 / sec_host,
                "path": "SEDCalculator.reconstruct: amplitudes from one projection pass per atom group, frames "
                        "synthesised on the GPU (left there / copied into fresh host arrays); points split over the ranks",
                "checked": ok and bool(all(np.isfinite(r["frames"]).all() for r in res[:2]))}
        del res

    # ---------------- CPU baseline (rank 0, N == 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        val, desc, sec = cpu_sample(cfg, traj, max_k=args.cpu_k)
        cpu = {"value": val, "unit": UNIT, "cores": host_threads(), "kind": "port", "sample": desc, "seconds": sec}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "s8 digit planes -> s32 (exact), f64 FFT", "data": "synthetic",
            "config": {"workload": args.workload, "desc": cfg["desc"], "k_points_per_gpu": n_k_local,
                       "l2_policy": "inputs larger than L2 (trajectory %.0f MB per array)" % (traj.positions.nbytes / 1e6),
                       "step": "mean positions + digit planes + phase table + tcgen05 projection + FFT/assembly"},
            "roofline": roof, "rooflines": rooflines, "kernels": kernels, "cpu_baseline": cpu, "e2e": e2e, "ised": ised,
            "gpu_launches": int(launches), "gpu_launches_per_step": int(launches) // max(1, args.steps), "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
