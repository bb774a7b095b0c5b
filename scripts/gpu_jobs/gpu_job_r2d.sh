mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu.py -m gpu -q -x -k "fft or consumers or cli" > gpurun_out/r2d_pytest.log 2>&1; tail -5 gpurun_out/r2d_pytest.log
for s in "16384 200" "16384 1024" "8192 100" "8192 1024" "32768 256"; do python scripts/fft_tune.py $s; PSA_FFT4_THREADS=256 python scripts/fft_tune.py $s; done 2>&1 | tee gpurun_out/r2d_fft_tune.log
python bench.py --workload c1 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-int8-peak > gpurun_out/r2d_plain_bench.log 2>gpurun_out/r2d_plain_bench.err && \
ncu --set full --import-source on --clock-control none -k regex:ised_batch_kernel -s 1 -c 2 -f -o gpurun_out/r2d_ised python bench.py --workload c1 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-int8-peak > gpurun_out/r2d_ncu_ised.log 2>&1
python -c "
import json; d=json.loads(open('gpurun_out/r2d_plain_bench.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step']); print({k:(round(v['ms'],3)) for k,v in d['kernels'].items()}); print({k:round(v['frac'],3) for k,v in d['rooflines'].items()}); print(d['ised']['kernel_ms'])"
