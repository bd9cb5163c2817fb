#!/bin/bash
# round-2 GPU call 24 (1 GPU): sparse pass with the branch-free remainder (gathers in groups of 4 / 8)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sparse.py -m gpu -x -q > gpurun_out/c24_sp_pytest.log 2>&1; echo "sparse pytest rc=$?"
RRI_SP_GRP=8 timeout 300 python -m pytest tests/test_gpu_sparse.py -m gpu -x -q -k long_factor > gpurun_out/c24_sp_pytest_g8.log 2>&1; echo "g8 pytest rc=$?"
for g in 4 8; do
  RRI_SP_GRP=$g timeout 300 python tools/bench_sparse.py 100000 rri 8 > gpurun_out/c24_sp_g$g.log 2>&1
  RRI_SP_GRP=$g timeout 300 python tools/bench_sparse.py 100000 rri 8 8 > gpurun_out/c24_sp_g${g}_r8.log 2>&1
done
RRI_SP_STREAM=0 timeout 300 python tools/bench_sparse.py 100000 rri 8 8 > gpurun_out/c24_sp_old_r8.log 2>&1
grep -H '^{' gpurun_out/c24_sp_*.log | cut -c1-400
