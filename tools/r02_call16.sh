#!/bin/bash
# round-2 GPU call 16 (1 GPU): full GPU suite on the current tree (per-topic sums once per call), default bench,
# A/B of the sparse pass kernel shapes (RRI_SP_VARIANT 0..3) with the sparse parity tests under each
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c16_pytest.log
timeout 600 python bench.py > gpurun_out/c16_bench.log 2> gpurun_out/c16_bench.err; echo "bench rc=$?"
for v in 0 1 2 3; do
  RRI_SP_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_sparse.py -m gpu -x -q > gpurun_out/c16_sp_pytest_v$v.log 2>&1; echo "v$v sparse pytest rc=$?"
  RRI_SP_VARIANT=$v timeout 300 python tools/bench_sparse.py 100000 rri 8 > gpurun_out/c16_sp_v$v.log 2>&1
  RRI_SP_VARIANT=$v timeout 300 python tools/bench_sparse.py 100000 rri 8 8 > gpurun_out/c16_sp_v${v}_r8.log 2>&1
done
tail -3 gpurun_out/c16_pytest.log
grep '^{' gpurun_out/c16_bench.log | cut -c1-400
for v in 0 1 2 3; do tail -1 gpurun_out/c16_sp_pytest_v$v.log; grep '^{' gpurun_out/c16_sp_v$v.log gpurun_out/c16_sp_v${v}_r8.log | cut -c1-420; done
