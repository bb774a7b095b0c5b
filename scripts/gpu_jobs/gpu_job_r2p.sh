#!/bin/bash
# Round 2, run 16 (8 GPUs): C4 strong scaling at N = 8, frame-sharded path (default) against the k-sharded one.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l)
        print(sys.argv[1], 'value %.3e ms %.3f e2e ms %.2f parity %s path %s' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d.get('parity_checked'), d['details'].get('multi_gpu_path')))
        print('  steps', d['details'].get('step_ms_first_median_last'), 'stages', d['e2e'].get('stage_ms_rank0'))
        print('  kernels', {k: round(v['ms'],3) for k,v in d['kernels'].items()}, d['clocks'])
PY
}
timeout 600 $TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2p_bench_c4_n8_frames.json 2> gpurun_out/r2p_bench_c4_n8_frames.err; echo "frames rc=$?"; show gpurun_out/r2p_bench_c4_n8_frames.json; grep -v Warning gpurun_out/r2p_bench_c4_n8_frames.err | tail -3 | cut -c1-300
PSA_B200_SHARD=k timeout 600 $TR --nproc-per-node 8 --master-port 29552 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2p_bench_c4_n8_k.json 2> gpurun_out/r2p_bench_c4_n8_k.err; echo "k rc=$?"; show gpurun_out/r2p_bench_c4_n8_k.json
timeout 900 $TR --nproc-per-node 8 --master-port 29553 bench.py --gpus 8 --workload c5 --steps 5 --warmup 3 > gpurun_out/r2p_bench_c5_n8_frames.json 2> gpurun_out/r2p_bench_c5_n8_frames.err; echo "c5 frames rc=$?"; show gpurun_out/r2p_bench_c5_n8_frames.json; grep -v Warning gpurun_out/r2p_bench_c5_n8_frames.err | tail -3 | cut -c1-300
