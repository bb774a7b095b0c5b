#!/bin/bash
# Round 2, run 19 (1 GPU): full gpu suite on the final tree (-s keeps the parity tables of the full-size tests), smoke().
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -s -m gpu > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2s_pytest.log | cut -c1-300; grep -n "0.0001" gpurun_out/r2s_pytest.log | cut -c1-400 | head -12
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2s_smoke.log
