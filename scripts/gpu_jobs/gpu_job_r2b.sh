mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu.py -m gpu -q -x -k "fft" > gpurun_out/r2b_pytest_fft.log 2>&1; tail -15 gpurun_out/r2b_pytest_fft.log
for s in "16384 200" "16384 1024" "8192 100" "32768 256"; do python scripts/fft_tune.py $s; PSA_FFT4=0 python scripts/fft_tune.py $s; done 2>&1 | tee gpurun_out/r2b_fft_tune.log
python scripts/fft_accuracy.py 2>&1 | tee gpurun_out/r2b_fft_accuracy.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu.py::test_full_size_config5_parity_on_k_subset > gpurun_out/r2b_pytest.log 2>&1; tail -25 gpurun_out/r2b_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench_c4.json 2> gpurun_out/r2b_bench_c4.err; tail -3 gpurun_out/r2b_bench_c4.err
python -c "
import json; d=json.loads(open('gpurun_out/r2b_bench_c4.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step']); print({k:(round(v['ms'],3)) for k,v in d['kernels'].items()}); print({k:round(v['frac'],3) for k,v in d['rooflines'].items()}); print(d['ised'])"
