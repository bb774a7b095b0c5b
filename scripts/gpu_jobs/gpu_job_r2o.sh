#!/bin/bash
# Round 2, run 15 (1 GPU): the driver's sequence on HEAD - full gpu test suite, smoke(), the default bench line and the reference arm.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2o_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2o_smoke.log
timeout 600 python bench.py --gpus 1 > gpurun_out/r2o_bench_c4_n1.json 2> gpurun_out/r2o_bench_c4_n1.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2o_bench_c4_n1.json; tail -3 gpurun_out/r2o_bench_c4_n1.err
timeout 300 python bench.py --gpus 1 --impl reference --steps 2 --warmup 1 > gpurun_out/r2o_bench_c4_ref.json 2> gpurun_out/r2o_bench_c4_ref.err; echo "ref rc=$?"; cut -c1-800 gpurun_out/r2o_bench_c4_ref.json
