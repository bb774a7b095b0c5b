#!/bin/bash
# Round 2, run 17 (1 GPU): full gpu suite with the routed projection + displacement-moment tests (-s keeps the parity tables), default bench line.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -s -m gpu > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2q_pytest.log | cut -c1-300; grep -n "0.0001" gpurun_out/r2q_pytest.log | cut -c1-400 | head -12
timeout 600 python bench.py --gpus 1 > gpurun_out/r2q_bench_c4_n1.json 2> gpurun_out/r2q_bench_c4_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r2q_bench_c4_n1.json') if l.startswith('{')][0]
print('value %.3e ms %.2f e2e ms %.2f' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step']), d['clocks'])
print('ised', d['ised']['kernel_ms'], d['ised']['seconds_whole_call_device_resident'], d['ised']['checked'])
print({k: round(v['frac'],3) for k,v in d['rooflines'].items()})
PY
tail -2 gpurun_out/r2q_bench_c4_n1.err | cut -c1-300
