"""Batch driver over the hot path (N4 of SURVEY.md 8f): the reference's ``psa`` command line
(reference: src/psa/cli.py:25-207) with the same YAML schema and the same output files, minus what is out of scope
here (OVITO loading, matplotlib plots).

    python -m psa_b200.cli --trajectory run.lammpstrj [--config cfg.yaml] [--output-dir out] [--chiral] [--dt 0.002]
                           [--nk 100] [--recalculate-sed] [--device 0]

What differs from the reference, on purpose:

* the trajectory comes from the ``.npy`` cache the reference's loader writes next to the trajectory file
  (``<stem>.positions.npy`` ..., reference: src/psa/io/loader.py:48-76, 363-387) and is streamed to the GPU;
* every direction is computed ONCE.  The reference computes each SED twice when there are several directions - a
  first full pass only to find the global maximum intensity for plot normalisation (cli.py:91-104).  Here the
  maximum of every direction is reduced on the device from the spectra already computed (``psa_intensity`` +
  ``psa_minmax``) and the global maximum is written to ``summary.json`` next to the per-direction bundles;
* results are the reference's own ``SED.save`` bundles (``sed_data_<regular|chiral>_<label>[_basis].*.npy``,
  cli.py:121-156), which its plotter / GUI / a later ``psa`` run load back unchanged.
"""
from __future__ import annotations

import argparse
import copy
import json
import logging
from pathlib import Path
from typing import Any, Dict, List, Optional

import numpy as np

from . import cache
from .sed import SED

logger = logging.getLogger(__name__)

DEFAULT_CONFIG: Dict[str, Any] = {                        # reference: cli.py:38-44
    "general": {"trajectory_file_format": "auto", "use_displacements": False, "save_npy_trajectory": True,
                "save_npy_sed_data": True, "chiral_mode_enabled": False},
    "md_system": {"dt": 0.001, "nx": 1, "ny": 1, "nz": 1, "lattice_parameter": None},
    "sed_calculation": {"directions": [[1, 0, 0]], "n_kpoints": 100, "bz_coverage": 1.0,
                        "polarization_indices_chiral": [0, 1], "basis": {"atom_indices": None, "atom_types": None}},
    "plotting": {"max_freq_2d": None},
    "ised": {"apply": False,
             "k_path": {"direction": "x", "characteristic_length": None, "n_points": 50, "bz_coverage": None},
             "target_point": {"k_value": 6.283, "w_value_thz": 10.0},
             "basis": {"atom_indices": None, "atom_types": None},
             "reconstruction": {"rescaling_factor": "auto", "num_animation_timesteps": 100,
                                "output_dump_filename": "ised_motion.dump"}},
}


def update_dict_recursively(base: dict, update: dict) -> dict:
    """Same merge rule as the reference (src/psa/utils/helpers.py:111-127)."""
    for key, val in update.items():
        if isinstance(val, dict) and isinstance(base.get(key), dict):
            update_dict_recursively(base[key], val)
        else:
            base[key] = val
    return base


def direction_label(spec, index: int) -> str:
    """File-name label of a direction (reference: cli.py:108-113)."""
    if isinstance(spec, (int, float)):
        return f"{spec:.1f}deg"
    if isinstance(spec, str):
        return spec.replace(" ", "_").replace("/", "-")
    if isinstance(spec, (list, tuple, np.ndarray)):
        arr = np.asarray(spec)
        return f"{arr.item():.1f}deg" if arr.size == 1 else "_".join(f"{x:.2f}" for x in arr)
    if isinstance(spec, dict):
        return f"h{spec.get('h', 0)}_k{spec.get('k', 0)}_l{spec.get('l', 0)}"
    return f"dir{index}"


def resolve_basis(types: np.ndarray, n_atoms: int, basis_cfg: dict):
    """``(indices|None, file-name suffix)`` of the main SED basis (reference: cli.py:79-89, 116-119)."""
    idx_spec, type_spec = basis_cfg.get("atom_indices"), basis_cfg.get("atom_types")
    idx, suffix = None, ""
    if idx_spec is not None and len(idx_spec) > 0:
        idx = np.asarray(idx_spec, dtype=int)
        suffix = "_idxbasis"
        if type_spec:
            logger.warning("Main SED: atom_indices and atom_types specified; using atom_indices.")
    elif type_spec is not None and len(type_spec) > 0:
        idx = np.where(np.isin(types, type_spec))[0]
        suffix = "_typebasis" + "_".join(map(str, type_spec))
        if not idx.size:
            logger.warning("Main SED: No atoms for types %s. Using all.", type_spec)
            idx, suffix = None, ""
    if idx is not None and (np.any(idx >= n_atoms) or np.any(idx < 0)):
        raise ValueError("Main SED basis indices out of bounds.")
    return idx, suffix


def run(args: argparse.Namespace) -> Dict[str, Any]:
    import torch

    from . import consumers
    from .calculator import SEDCalculator

    out_dir = Path(args.output_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    config = copy.deepcopy(DEFAULT_CONFIG)
    if args.config:
        import yaml
        with open(args.config) as fh:
            user = yaml.safe_load(fh)
        if user:
            update_dict_recursively(config, user)
        logger.info("Loaded config from %s", args.config)
    if args.dt is not None:
        config["md_system"]["dt"] = args.dt
    if args.nk is not None:
        config["sed_calculation"]["n_kpoints"] = args.nk
    if args.chiral:
        config["general"]["chiral_mode_enabled"] = True
    gen, md, sed_cfg, ised_cfg = config["general"], config["md_system"], config["sed_calculation"], config["ised"]
    if md["dt"] <= 0:
        raise ValueError("Timestep 'dt' must be positive.")
    if not cache.has_npy_cache(args.trajectory):
        raise FileNotFoundError(
            f"{args.trajectory}: no .npy cache next to it ({', '.join(p.name for p in cache.cache_files(args.trajectory).values())}). "
            "Trajectory parsing (OVITO) is the reference's job: load the file once with psa.io.loader.TrajectoryLoader "
            "and save_trajectory_npy(), or write the four arrays with psa_b200.cache.save_npy_cache().")
    traj = cache.load_npy_cache(args.trajectory, dt=md["dt"])
    logger.info("Trajectory: %d frames x %d atoms (dt=%.4f ps)", traj.n_frames, traj.n_atoms, md["dt"])
    calc = SEDCalculator(traj, md["nx"], md["ny"], md["nz"], use_displacements=gen["use_displacements"], device=args.device)

    lat = md.get("lattice_parameter")
    if lat is None or lat <= 1e-6:
        lat = float(np.linalg.norm(calc.a1))
        if not lat > 1e-6:
            raise ValueError("Cannot determine valid effective_lattice_parameter. Specify in config or check box/nx,ny,nz.")
        logger.info("Using |a1| (%.3f A) as effective lattice parameter.", lat)
    md["lattice_parameter"] = lat
    basis_idx, basis_sfx = resolve_basis(np.asarray(traj.types), traj.n_atoms, sed_cfg["basis"])
    chiral = bool(gen["chiral_mode_enabled"])
    kind = "chiral" if chiral else "regular"
    eng = calc.engine

    summary: Dict[str, Any] = {"directions": [], "global_max_intensity": None}
    maxima: List[float] = []
    for i_d, spec in enumerate(sed_cfg["directions"], 1):
        label = direction_label(spec, i_d)
        base = out_dir / f"sed_data_{kind}_{label}{basis_sfx}"
        entry: Dict[str, Any] = {"label": label, "files": base.name + ".*.npy", "loaded_from_cache": False}
        res: Optional[SED] = None
        if gen["save_npy_sed_data"] and not args.recalculate_sed:
            try:
                res = SED.load(base)
                entry["loaded_from_cache"] = True
                logger.info("Loaded SED data for %s.", label)
            except FileNotFoundError:
                logger.info("No pre-calculated SED for %s. Will calculate.", label)
        if res is not None and chiral and res.phase is None:
            logger.info("Recalculating SED for %s (phase data needed).", label)
            res = None
        if res is None:
            k_mags, k_vecs = calc.get_k_path(spec, sed_cfg["bz_coverage"], sed_cfg["n_kpoints"], lat)
            dev_sed, complex_out, groups = calc._calculate_device(k_vecs, basis_idx, None, "coherent")
            with torch.cuda.device(eng.device):
                inten = eng.intensity(dev_sed)                              # reduced on the device: one pass per direction
                _, peak, _ = consumers.nan_range(eng, inten)
                phase = None
                if chiral:
                    pol = sed_cfg["polarization_indices_chiral"]
                    if len(pol) >= 2 and max(pol) < 3:
                        phase = calc._to_host(calc._chiral_phase_of_result(dev_sed, (int(pol[0]), int(pol[1]))))
                    else:
                        logger.error("Chiral mode error for %s: invalid polarization indices %s.", label, pol)
                sed_host = calc._to_host(dev_sed)
            res = SED(sed_host, np.fft.fftfreq(traj.n_frames, d=calc.dt_ps), k_mags, k_vecs, k_grid_shape=None,
                      phase=phase, is_complex=complex_out)
            if gen["save_npy_sed_data"]:
                res.save(base)
            entry["max_intensity"] = float(peak)
        else:
            entry["max_intensity"] = float(np.max(res.intensity)) if res.sed.size else 0.0
        maxima.append(entry["max_intensity"])
        summary["directions"].append(entry)
    if maxima and not chiral:
        summary["global_max_intensity"] = float(max(maxima))
        logger.info("Global max intensity: %.4e", summary["global_max_intensity"])

    if ised_cfg["apply"]:
        kp, tgt, bas, rec = ised_cfg["k_path"], ised_cfg["target_point"], ised_cfg["basis"], ised_cfg["reconstruction"]
        dump = out_dir / rec["output_dump_filename"]
        calc.ised(k_dir_spec=kp["direction"], k_target=float(tgt["k_value"]), w_target=float(tgt["w_value_thz"]),
                  char_len_k_path=float(kp["characteristic_length"] or lat), nk_on_path=int(kp["n_points"]),
                  bz_cov_ised=float(kp["bz_coverage"] or sed_cfg["bz_coverage"]),
                  basis_atom_idx_ised=bas.get("atom_indices"), basis_atom_types_ised=bas.get("atom_types"),
                  rescale_factor=rec["rescaling_factor"], n_recon_frames=int(rec["num_animation_timesteps"]),
                  dump_filepath=str(dump))
        summary["ised_dump"] = dump.name
    (out_dir / "summary.json").write_text(json.dumps(summary, indent=1))
    logger.info("PSA processing completed.")
    return summary


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Phonon Spectral Analysis - B200 batch driver (psa_b200).")
    p.add_argument("--trajectory", type=str, required=True, help="Path of the MD trajectory (its .npy cache is read).")
    p.add_argument("--config", type=str, help="Path to YAML configuration file (the reference's schema).")
    p.add_argument("--output-dir", type=str, default="psa_output", help="Directory for results.")
    p.add_argument("--chiral", action="store_true", help="Enable chiral SED (overrides config).")
    p.add_argument("--dt", type=float, help="Override MD timestep from config (ps).")
    p.add_argument("--nk", type=int, help="Override n_kpoints for SED from config.")
    p.add_argument("--recalculate-sed", action="store_true", help="Force recalculation of SED data.")
    p.add_argument("--device", type=int, default=None, help="CUDA device index.")
    return p


def main(argv=None) -> int:
    logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s", datefmt="%H:%M:%S")
    args = build_parser().parse_args(argv)
    try:
        run(args)
    except FileNotFoundError as exc:
        logger.error("File Error: %s", exc)
        return 1
    except ValueError as exc:
        logger.error("Value Error: %s", exc)
        return 1
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
