#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tests/multi_gpu_worker.py > gpurun_out/mw.log 2>&1; echo "exit $?" >> gpurun_out/mw.log
grep -v "^W0\|OMP_NUM" gpurun_out/mw.log | tail -40
