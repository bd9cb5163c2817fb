#!/bin/bash
# round-2 GPU call 20 (1 GPU): streamlined sparse pass as the default + larger staging blocks (longer sub-segments)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sparse.py -m gpu -x -q > gpurun_out/c20_sp_pytest.log 2>&1; echo "sparse pytest rc=$?"
for kb in 128 160 200 216; do
  RRI_SP_BLOCK_KB=$kb timeout 300 python tools/bench_sparse.py 100000 rri 8 > gpurun_out/c20_sp_kb$kb.log 2>&1
  RRI_SP_BLOCK_KB=$kb timeout 300 python tools/bench_sparse.py 100000 rri 8 8 > gpurun_out/c20_sp_kb${kb}_r8.log 2>&1
done
RRI_SP_BLOCK_KB=200 timeout 600 python -m pytest tests/test_gpu_sparse.py -m gpu -x -q > gpurun_out/c20_sp_pytest_kb200.log 2>&1; echo "sparse pytest kb200 rc=$?"
tail -2 gpurun_out/c20_sp_pytest.log gpurun_out/c20_sp_pytest_kb200.log
grep -H '^{' gpurun_out/c20_sp_kb*.log | cut -c1-400
