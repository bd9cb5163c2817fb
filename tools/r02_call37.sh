#!/bin/bash
# round-2 GPU call 37 (1 GPU): config-4 bench lines (observed entries) with the final sparse pass kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 400 python bench.py --config cfg4 --masked sparse --steps 8 --warmup 3 --no-cpu > gpurun_out/c37_cfg4_sparse.log 2> gpurun_out/c37_cfg4_sparse.err; echo "rc=$?"
timeout 400 python bench.py --config cfg4 --masked sparse --steps 8 --warmup 3 --no-cpu --no-e2e --refresh-every 8 > gpurun_out/c37_cfg4_sparse_r8.log 2> gpurun_out/c37_cfg4_sparse_r8.err; echo "rc=$?"
for f in c37_cfg4_sparse c37_cfg4_sparse_r8; do grep '^{' gpurun_out/$f.log | cut -c1-900; done
