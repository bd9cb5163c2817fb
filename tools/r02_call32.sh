#!/bin/bash
# round-2 GPU call 32 (4 GPUs): final code at 4 ranks -- parity (peer exchange) and the config-3 bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 tests/multi_gpu_check.py > gpurun_out/c32_mg4.log 2>&1; echo "mg4 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/c32_bench4.log 2> gpurun_out/c32_bench4.err; echo "bench4 rc=$?"
grep -v "Warning\|warn" gpurun_out/c32_mg4.log | tail -10
grep '^{' gpurun_out/c32_bench4.log | cut -c1-300
