#!/bin/bash
# round-2 GPU call 19 (1 GPU): upper bound of a bank-aware entry order -- the pass kernel with a conflict-free gather
# (RRI_SP_FAKE_IDX=1: timing only) on the unchanged random matrix
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for v in 0 5; do
  for f in 0 1; do
    RRI_SP_FAKE_IDX=$f RRI_SP_VARIANT=$v timeout 300 python tools/bench_sparse.py 100000 rri 8 8 > gpurun_out/c19_sp_v${v}_fake$f.log 2>&1
  done
done
grep -H '^{' gpurun_out/c19_sp_*.log | cut -c1-420
