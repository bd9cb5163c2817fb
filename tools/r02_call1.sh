#!/bin/bash
# round-2 GPU call 1 (1 GPU): tests, headline bench (both arms), cfg5 shard, sparse pass A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/c1_gpu.txt 2>&1
nproc > gpurun_out/c1_host.txt; free -g >> gpurun_out/c1_host.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c1_bench.log 2> gpurun_out/c1_bench.err; echo "rc=$?" >> gpurun_out/c1_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c1_ref.log 2> gpurun_out/c1_ref.err; echo "rc=$?" >> gpurun_out/c1_ref.err
timeout 600 python bench.py --config cfg5 --rows 125000 --steps 20 --warmup 5 --no-cpu --no-pageable > gpurun_out/c1_cfg5shard.log 2> gpurun_out/c1_cfg5shard.err; echo "rc=$?" >> gpurun_out/c1_cfg5shard.err
for v in "" "RRI_SP_PASS_V2=1" "RRI_SP_BATCHED_SUMS=1" "RRI_SP_PASS_V2=1 RRI_SP_BATCHED_SUMS=1"; do
  echo "== $v" >> gpurun_out/c1_sparse.log
  env $v timeout 300 python tools/bench_sparse.py 100000 rri 5 >> gpurun_out/c1_sparse.log 2>&1
done
tail -3 gpurun_out/c1_pytest.log; cat gpurun_out/c1_bench.log | cut -c1-600; cat gpurun_out/c1_sparse.log | grep -v Warn | cut -c1-300
