#!/bin/bash
# round-2 GPU call 21 (1 GPU): software-pipelined sparse pass (RRI_SP_PIPE 1..4) against the default
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for v in 1 2 4; do
  RRI_SP_PIPE=$v timeout 300 python -m pytest tests/test_gpu_sparse.py -m gpu -x -q -k long_factor > gpurun_out/c21_sp_pytest_p$v.log 2>&1; echo "pipe $v pytest rc=$?"
done
for v in 0 1 2 3 4; do
  RRI_SP_PIPE=$v timeout 300 python tools/bench_sparse.py 100000 rri 8 8 > gpurun_out/c21_sp_p$v.log 2>&1
done
grep -H '^{' gpurun_out/c21_sp_p*.log | cut -c1-400
