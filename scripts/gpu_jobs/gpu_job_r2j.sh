mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu.py::test_full_size_config5_parity_on_k_subset > gpurun_out/r2j_pytest.log 2>&1; tail -4 gpurun_out/r2j_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu_check.py > gpurun_out/r2j_multigpu_check.log 2>&1; grep MULTIGPU gpurun_out/r2j_multigpu_check.log; tail -2 gpurun_out/r2j_multigpu_check.log | cut -c1-300
for pipe in 1 0; do
PSA_B200_PIPELINE_EXCHANGE=$pipe timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$pipe bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2j_bench_c4_n2_pipe$pipe.json 2> gpurun_out/r2j_bench_c4_n2_pipe$pipe.err; tail -2 gpurun_out/r2j_bench_c4_n2_pipe$pipe.err | cut -c1-200
done
python - <<'PY'
import json
for f in ("gpurun_out/r2j_bench_c4_n2_pipe1.json","gpurun_out/r2j_bench_c4_n2_pipe0.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3e ms %.3f e2e ms %.2f parity %s"%(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("parity_checked")))
        print(" stages", d["e2e"]["stage_ms_rank0"]); print(" kernels", {k:round(v["ms"],3) for k,v in d["kernels"].items()})
    except Exception as e: print(f, "ERR", e)
PY
