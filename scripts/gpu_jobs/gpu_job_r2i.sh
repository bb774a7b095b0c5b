mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu_check.py > gpurun_out/r2i_multigpu_check_n8.log 2>&1; grep MULTIGPU gpurun_out/r2i_multigpu_check_n8.log; tail -3 gpurun_out/r2i_multigpu_check_n8.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2i_bench_c4_n8.json 2> gpurun_out/r2i_bench_c4_n8.err; tail -3 gpurun_out/r2i_bench_c4_n8.err | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2i_bench_c4_n4.json 2> gpurun_out/r2i_bench_c4_n4.err; tail -2 gpurun_out/r2i_bench_c4_n4.err | cut -c1-300
python - <<'PY'
import json
for f in ("gpurun_out/r2i_bench_c4_n8.json","gpurun_out/r2i_bench_c4_n4.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3e ms %.3f e2e ms %.2f parity %s"%(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("parity_checked")))
        print(" stages", d["e2e"]["stage_ms_rank0"]); print(" kernels", {k:round(v["ms"],3) for k,v in d["kernels"].items()}); print(" clocks", d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
