mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu_check.py > gpurun_out/r2e_multigpu_check.log 2>&1; tail -6 gpurun_out/r2e_multigpu_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --workload c2 --steps 5 --warmup 3 > gpurun_out/r2e_bench_c2_n2.json 2> gpurun_out/r2e_bench_c2_n2.err; tail -4 gpurun_out/r2e_bench_c2_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2e_bench_c4_n2.json 2> gpurun_out/r2e_bench_c4_n2.err; tail -4 gpurun_out/r2e_bench_c4_n2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2e_bench_c2_n2.json","gpurun_out/r2e_bench_c4_n2.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3e ms %.3f e2e ms %.2f parity %s"%(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("parity_checked")))
        print(" stages", d["e2e"]["stage_ms_rank0"]); print(" kernels", {k:round(v["ms"],3) for k,v in d["kernels"].items()})
    except Exception as e: print(f, "ERR", e)
PY
