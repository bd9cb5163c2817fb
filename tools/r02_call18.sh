#!/bin/bash
# round-2 GPU call 18 (1 GPU): is the sparse pass bound by shared-memory bank conflicts of the gather?
# regular patterns: spacing 21 (conflict-free quarter-warps), 19 (conflict-free), 24 (8-way conflicts), random 5 %
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for v in 0 5; do
  for pat in 21 19 24 20; do
    RRI_SP_VARIANT=$v timeout 300 python tools/bench_sparse.py 100000 rri 8 --pattern=$pat > gpurun_out/c18_sp_v${v}_p$pat.log 2>&1
  done
done
grep -h '^{' gpurun_out/c18_sp_*.log | cut -c1-420
