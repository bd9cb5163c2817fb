#!/bin/bash
# round-2 GPU call 34 (1 GPU): ncu --set full of the contraction on a 25 000-row shard of config 3 (what a rank of the
# 8-GPU run computes): where do the 5-7 % against the full-size launch go?
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --rows 25000 --steps 2 --warmup 1 --no-e2e --no-cpu --no-rri"
timeout 300 $CMD > gpurun_out/c34_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tf32_gemm -s 6 -c 2 -f -o gpurun_out/r02_tf32_gemm_25k $CMD > gpurun_out/c34_ncu.log 2>&1
echo "rc=$?"; grep '^{' gpurun_out/c34_plain.log | cut -c1-200
