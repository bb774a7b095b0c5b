#!/bin/bash
# Round 2, run 18 (2 GPUs): routed frame-sharded projection (one launch per chunk for all owners) - bitwise check, C4 bench.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29561 tests/multigpu_check.py > gpurun_out/r2r_multigpu_check.log 2>&1; echo "multigpu_check rc=$?"; grep MULTIGPU_CHECK gpurun_out/r2r_multigpu_check.log | head -1 | cut -c1-400; grep -v Warning gpurun_out/r2r_multigpu_check.log | tail -4 | cut -c1-300
timeout 600 $TR --nproc-per-node 2 --master-port 29562 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2r_bench_c4_n2.json 2> gpurun_out/r2r_bench_c4_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r2r_bench_c4_n2.json') if l.startswith('{')][0]
print('value %.3e ms %.3f e2e ms %.2f parity %s path %s' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d.get('parity_checked'), d['details'].get('multi_gpu_path')))
print('  steps', d['details'].get('step_ms_first_median_last'), 'stages', d['e2e'].get('stage_ms_rank0'))
print('  kernels', {k: (round(v['ms'],3), v['calls']) for k,v in d['kernels'].items()}, d['clocks'])
PY
grep -v Warning gpurun_out/r2r_bench_c4_n2.err | tail -3 | cut -c1-300
