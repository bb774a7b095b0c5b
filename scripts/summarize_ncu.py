#!/usr/bin/env python
"""Turn ncu artefacts brought back in gpurun_out/ into the small text/JSON summaries kept under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches_c2.csv  profiles/r01_launches_c2.md
    python scripts/summarize_ncu.py full     gpurun_out/prof_c2.ncu-rep  profiles/r01_ncu_full_c2.md [workload]
"""
import csv
import json
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.avg.per_second"]


def short(name):
    return name.split("(")[0].replace("void ", "").strip()


def launches(src, dst):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    per = defaultdict(list)
    for r in rows[1:]:
        if r[mi] == "gpu__time_duration.sum":
            per[short(r[ki])].append(float(r[vi].replace(",", "")))
    total = sum(sum(v) for v in per.values())
    out = [f"# ncu launch list ({src})", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none` - per-launch times are cold-cache and "
           "serialised; compare SHARES with bench.py's CUDA-event breakdown, not absolutes.", "",
           "| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| {k} | {len(v)} | {sum(v) / 1e3:.1f} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / total:.3f} |")
    Path(dst).write_text("\n".join(out) + "\n")
    print("\n".join(out))


def full(src, dst, workload=None):
    txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {k: hdr.index(k) for k in KEYS if k in hdr}
    ki = hdr.index("Kernel Name")
    out = [f"# ncu --set full summary ({src})", ""]
    traffic = {}
    for r in rows[2:]:
        name = short(r[ki])
        out.append(f"## {name}")
        for k, i in idx.items():
            out.append(f"- {k}: {r[i]} {units[i]}")
        try:
            rd, wr = float(r[idx["dram__bytes_read.sum"]]), float(r[idx["dram__bytes_write.sum"]])
            mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[idx["dram__bytes_read.sum"]]]
            mult_w = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[idx["dram__bytes_write.sum"]]]
            traffic[name] = rd * mult + wr * mult_w
            out.append(f"- dram traffic per launch: {traffic[name] / 1e6:.1f} MB")
        except Exception:
            pass
        out.append("")
    Path(dst).write_text("\n".join(out) + "\n")
    print("\n".join(out))
    if workload:
        import re

        def api_name(kernel):                       # kernel symbol -> the C-ABI call that bench.py times
            base = re.sub(r"<.*>$", "", kernel).split("::")[-1]
            return {"mean_positions_kernel": "psa_mean_positions", "mean_positions_tma_kernel": "psa_mean_positions",
                    "digitize_kernel": "psa_digitize", "phase_digits_kernel": "psa_phase_digits",
                    "project_tc_kernel": "psa_project", "project_tc2_kernel": "psa_project",
                    "fft_sed_kernel": "psa_fft_sed", "fft4_kernel": "psa_fft_sed", "digitize_cluster_kernel": "psa_digitize",
                    "ised_batch_kernel": "psa_ised_frames"}.get(base, kernel)
        api = {k: api_name(k) for k in traffic}
        p = Path(dst).parent / "ncu_traffic.json"
        data = json.loads(p.read_text()) if p.exists() else {}
        data.setdefault(workload, {}).update({api.get(k, k): v for k, v in traffic.items()})
        p.write_text(json.dumps(data, indent=1) + "\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:])
