#!/bin/bash
# Round 2, run 20 (4 GPUs): bitwise multi-GPU check on 4 ranks with the routed frame-sharded projection.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 4 --master-port 29571 tests/multigpu_check.py > gpurun_out/r2t_multigpu_check_n4.log 2>&1; echo "multigpu_check rc=$?"; grep MULTIGPU_CHECK gpurun_out/r2t_multigpu_check_n4.log | head -1 | cut -c1-400; grep -v Warning gpurun_out/r2t_multigpu_check_n4.log | tail -3 | cut -c1-300
