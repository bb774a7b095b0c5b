mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu.py -m gpu -q -x -k "fft or consumers or cli or ised" > gpurun_out/r2f_pytest.log 2>&1; tail -5 gpurun_out/r2f_pytest.log
for v in "128 6" "128 5" "128 4" "256 3" "256 2"; do set -- $v; for s in "16384 1024" "8192 1024"; do echo "threads=$1 occ=$2"; PSA_FFT4_THREADS=$1 PSA_FFT4_OCC=$2 python scripts/fft_tune.py $s; done; done 2>&1 | tee gpurun_out/r2f_fft_tune.log
for v in "3" "2"; do echo "occ=$v"; PSA_FFT4_OCC=$v python scripts/fft_tune.py 32768 256; done 2>&1 | tee -a gpurun_out/r2f_fft_tune.log
python bench.py --workload c1 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-int8-peak > gpurun_out/r2f_bench_c1.log 2>gpurun_out/r2f_bench_c1.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-int8-peak > gpurun_out/r2f_bench_c4.log 2>gpurun_out/r2f_bench_c4.err
python - <<'PY'
import json
for f in ("gpurun_out/r2f_bench_c1.log","gpurun_out/r2f_bench_c4.log"):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e'] and d['e2e']['ms_per_step']); print({k:(round(v['ms'],3)) for k,v in d['kernels'].items()}); print({k:round(v['frac'],3) for k,v in d['rooflines'].items()}); print(d['ised']['kernel_ms'])
PY
