mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu.py -m gpu -q -x -k "consumers or cli" > gpurun_out/r2c_pytest.log 2>&1; tail -8 gpurun_out/r2c_pytest.log
python scripts/fft_tune.py 16384 1024 > gpurun_out/r2c_plain_fft.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:fft4_kernel -s 3 -c 1 -f -o gpurun_out/r2c_fft4 python scripts/fft_tune.py 16384 1024 > gpurun_out/r2c_ncu_fft.log 2>&1
tail -3 gpurun_out/r2c_ncu_fft.log
python bench.py --workload c2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-int8-peak > gpurun_out/r2c_plain_bench.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:ised_batch_kernel -c 2 -f -o gpurun_out/r2c_ised python bench.py --workload c2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-int8-peak > gpurun_out/r2c_ncu_ised.log 2>&1
tail -3 gpurun_out/r2c_ncu_ised.log
ls -la gpurun_out/*.ncu-rep
