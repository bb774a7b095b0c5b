#!/bin/bash
# Round 2, run 13: streamed ingest (one GPU) - tests, then e2e on C4 / C5 / C2 with and without it.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "streamed or arrival or frame_ranges or npy_cache or gold" > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2m_pytest.log
timeout 900 python -m pytest tests -x -q -s -m gpu -k "full_size" > gpurun_out/r2m_pytest_full.log 2>&1; echo "pytest full rc=$?"; grep -n "1e-06\|passed\|failed" gpurun_out/r2m_pytest_full.log | tail -8
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l)
        print(sys.argv[1], 'value %.3e ms %.3f e2e ms %.2f' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step']), d['details'].get('step_ms_first_median_last'), d['clocks'])
PY
}
for w in c4 c5 c2 c1; do
  st=20; [ $w = c5 ] && st=5
  timeout 1200 python bench.py --workload $w --steps $st --warmup 5 --no-cpu-baseline --no-ised --no-int8-peak > gpurun_out/r2m_bench_$w.json 2> gpurun_out/r2m_bench_$w.err; echo "$w rc=$?"; show gpurun_out/r2m_bench_$w.json
  PSA_B200_STREAM_INGEST=0 timeout 1200 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-ised --no-int8-peak > gpurun_out/r2m_bench_${w}_nostream.json 2> gpurun_out/r2m_bench_${w}_nostream.err; echo "$w nostream rc=$?"; show gpurun_out/r2m_bench_${w}_nostream.json
done
