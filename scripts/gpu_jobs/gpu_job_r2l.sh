mkdir -p gpurun_out
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2l_bench_c4_n$n.json 2> gpurun_out/r2l_bench_c4_n$n.err; tail -2 gpurun_out/r2l_bench_c4_n$n.err | cut -c1-200
done
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 8 --workload c5 --steps 5 --warmup 3 > gpurun_out/r2l_bench_c5_n8.json 2> gpurun_out/r2l_bench_c5_n8.err; tail -2 gpurun_out/r2l_bench_c5_n8.err | cut -c1-200
python - <<'PY'
import json
for f in ("gpurun_out/r2l_bench_c4_n8.json","gpurun_out/r2l_bench_c4_n4.json","gpurun_out/r2l_bench_c5_n8.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3e ms %.3f e2e ms %.2f parity %s"%(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("parity_checked")))
        print(" stages", d["e2e"]["stage_ms_rank0"]); print(" kernels", {k:round(v["ms"],3) for k,v in d["kernels"].items()}); print(" clocks", d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
