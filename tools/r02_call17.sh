#!/bin/bash
# round-2 GPU call 17 (1 GPU): streamlined sparse pass kernel (sp_pass_stream_kernel), RRI_SP_VARIANT 4..8 against 0
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for v in 4 5 7; do
  RRI_SP_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_sparse.py -m gpu -x -q > gpurun_out/c17_sp_pytest_v$v.log 2>&1; echo "v$v sparse pytest rc=$?"
done
for v in 0 4 5 6 7 8; do
  RRI_SP_VARIANT=$v timeout 300 python tools/bench_sparse.py 100000 rri 8 > gpurun_out/c17_sp_v$v.log 2>&1
done
for v in 4 5 7; do tail -1 gpurun_out/c17_sp_pytest_v$v.log; done
for v in 0 4 5 6 7 8; do grep -H '^{' gpurun_out/c17_sp_v$v.log | cut -c1-420; done
