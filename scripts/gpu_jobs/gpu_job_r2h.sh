mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu.py::test_full_size_config5_parity_on_k_subset > gpurun_out/r2h_pytest.log 2>&1; tail -4 gpurun_out/r2h_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench_c4.json 2>gpurun_out/r2h_bench_c4.err
python bench.py --workload c1 --steps 10 --warmup 3 > gpurun_out/r2h_bench_c1.json 2>gpurun_out/r2h_bench_c1.err
python bench.py --workload c5 --steps 3 --warmup 2 --cpu-k 3 > gpurun_out/r2h_bench_c5.json 2>gpurun_out/r2h_bench_c5.err; tail -3 gpurun_out/r2h_bench_c5.err
PSA_TEST_C5=1 timeout 1500 python -m pytest tests/test_gpu.py -m gpu -q -x -s -k config5 > gpurun_out/r2h_pytest_c5.log 2>&1; tail -8 gpurun_out/r2h_pytest_c5.log
python - <<'PY'
import json
for w in ("c4","c1","c5"):
    f=f"gpurun_out/r2h_bench_{w}.json"
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(w, "%.3e"%d['value'], round(d['ms_per_step'],3), "e2e", d['e2e'] and round(d['e2e']['ms_per_step'],2), "cpu", d['cpu_baseline'] and "%.3e"%d['cpu_baseline']['value']); print('  ms', {k:(round(v['ms'],3)) for k,v in d['kernels'].items()}); print('  frac', {k:round(v['frac'],3) for k,v in d['rooflines'].items()}); print('  ised', {k:round(v,3) for k,v in d['ised']['kernel_ms'].items()}, d['ised']['seconds_whole_call_device_resident'])
    except Exception as e: print(w, "ERR", e, open(f.replace('.json','.err')).read()[-1500:])
PY
