#!/bin/bash
# Round 2, run 14 (2 GPUs): frame-sharded multi-GPU path - parity, then C4 A/B against the k-sharded path;
# on GPU 0 alone: PCIe copy rates and the one-GPU e2e timeline.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 2 --master-port 29541 tests/multigpu_check.py > gpurun_out/r2n_multigpu_check.log 2>&1; echo "multigpu_check rc=$?"; grep MULTIGPU_CHECK gpurun_out/r2n_multigpu_check.log | cut -c1-400; tail -5 gpurun_out/r2n_multigpu_check.log | cut -c1-300
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
timeout 900 $TR --nproc-per-node 2 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2n_bench_c4_n2_frames.json 2> gpurun_out/r2n_bench_c4_n2_frames.err; echo "frames rc=$?"; show gpurun_out/r2n_bench_c4_n2_frames.json; tail -3 gpurun_out/r2n_bench_c4_n2_frames.err | cut -c1-300
PSA_B200_SHARD=k timeout 900 $TR --nproc-per-node 2 --master-port 29543 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2n_bench_c4_n2_k.json 2> gpurun_out/r2n_bench_c4_n2_k.err; echo "k rc=$?"; show gpurun_out/r2n_bench_c4_n2_k.json


