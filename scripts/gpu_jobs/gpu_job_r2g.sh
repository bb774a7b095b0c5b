mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu.py::test_full_size_config5_parity_on_k_subset > gpurun_out/r2g_pytest.log 2>&1; tail -6 gpurun_out/r2g_pytest.log
for v in "4 0" "5 0" "4 1" "5 1"; do set -- $v; echo "occ=$1 prefetch=$2"; PSA_FFT4_OCC=$1 PSA_FFT4_PREFETCH=$2 python scripts/fft_tune.py 16384 1024; done 2>&1 | tee gpurun_out/r2g_fft_tune.log
for w in c1 c2 c4; do python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-int8-peak --no-e2e > gpurun_out/r2g_bench_$w.log 2>gpurun_out/r2g_bench_$w.err; done
python - <<'PY'
import json
for w in ("c1","c2","c4"):
    f=f"gpurun_out/r2g_bench_{w}.log"
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(w, "%.3e"%d['value'], round(d['ms_per_step'],3)); print('  ms', {k:(round(v['ms'],3)) for k,v in d['kernels'].items()}); print('  frac', {k:round(v['frac'],3) for k,v in d['rooflines'].items()}); print('  ised', {k:round(v,3) for k,v in d['ised']['kernel_ms'].items()})
    except Exception as e: print(w, "ERR", e, open(f.replace('.log','.err')).read()[-1500:])
PY
