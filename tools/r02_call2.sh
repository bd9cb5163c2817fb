#!/bin/bash
# round-2 GPU call 2 (2 GPUs): tests incl. 2-rank parity, bench N=1 / N=2, cfg5 shard
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c2_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-pageable > gpurun_out/c2_bench1.log 2> gpurun_out/c2_bench1.err; echo "rc=$?" >> gpurun_out/c2_bench1.err
timeout 600 python bench.py --config cfg5 --rows 125000 --steps 20 --warmup 5 --no-cpu --no-pageable --no-e2e > gpurun_out/c2_cfg5shard.log 2> gpurun_out/c2_cfg5shard.err; echo "rc=$?" >> gpurun_out/c2_cfg5shard.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-pageable > gpurun_out/c2_bench2.log 2> gpurun_out/c2_bench2.err; echo "rc=$?" >> gpurun_out/c2_bench2.err
RRI_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 --no-pageable --no-rri --no-e2e > gpurun_out/c2_bench2_nccl.log 2> gpurun_out/c2_bench2_nccl.err; echo "rc=$?" >> gpurun_out/c2_bench2_nccl.err
tail -5 gpurun_out/c2_pytest.log; for f in c2_bench1 c2_cfg5shard c2_bench2 c2_bench2_nccl; do echo "== $f"; tail -2 gpurun_out/$f.err | cut -c1-300; cut -c1-400 gpurun_out/$f.log; done
