mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu.py -m gpu -q -x -k "ised" > gpurun_out/r2k_pytest.log 2>&1; tail -3 gpurun_out/r2k_pytest.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2k_bench_c4.json 2>gpurun_out/r2k_bench_c4.err; tail -2 gpurun_out/r2k_bench_c4.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2k_ref_c4.json 2>gpurun_out/r2k_ref_c4.err
for kc in 512 256; do PSA_B200_K_CHUNK=$kc python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-int8-peak --no-e2e --no-ised > gpurun_out/r2k_bench_c4_kc$kc.json 2>/dev/null; done
python bench.py --workload c2 --steps 20 --warmup 5 > gpurun_out/r2k_bench_c2.json 2>gpurun_out/r2k_bench_c2.err
python bench.py --workload c3 --steps 20 --warmup 5 > gpurun_out/r2k_bench_c3.json 2>gpurun_out/r2k_bench_c3.err
python - <<'PY'
import json
for w in ("c4","c4_kc512","c4_kc256","c2","c3"):
    f=f"gpurun_out/r2k_bench_{w}.json"
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(w, "%.3e"%d['value'], round(d['ms_per_step'],3), "e2e", d['e2e'] and round(d['e2e']['ms_per_step'],2), "cpu", d['cpu_baseline'] and "%.3e"%d['cpu_baseline']['value'], d['clocks']); print('  ms', {k:(round(v['ms'],3)) for k,v in d['kernels'].items()}); print('  frac', {k:round(v['frac'],3) for k,v in d['rooflines'].items()});
        if d['ised']: print('  ised', {k:round(v,3) for k,v in d['ised']['kernel_ms'].items()})
    except Exception as e: print(w, "ERR", e)
print(open("gpurun_out/r2k_ref_c4.json").read()[:700])
PY
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-ised --no-int8-peak > gpurun_out/r2k_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_c4.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-ised --no-int8-peak > gpurun_out/r2k_ncu_launches.log 2>&1
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-ised --no-int8-peak > gpurun_out/r2k_plain2.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"project_tc2_kernel|fft4_kernel|digitize|mean_positions" -s 22 -c 6 -f -o gpurun_out/r02_ncu_full_c4 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-ised --no-int8-peak > gpurun_out/r2k_ncu_full.log 2>&1
ls -la gpurun_out/r02_*
