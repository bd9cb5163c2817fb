#!/bin/bash
# round-2 GPU call 6 (2 GPUs): all tests (incl. 2-rank parity), masked kernel v3, init timing, bench N=1/N=2 with half-step breakdown
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c6_pytest.log
timeout 300 python tools/bench_masked.py 100000 rri tf32 > gpurun_out/c6_masked.log 2>&1
timeout 300 python tools/bench_init.py 200000 > gpurun_out/c6_init.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-pageable --no-rri > gpurun_out/c6_bench2.log 2> gpurun_out/c6_bench2.err; echo "rc=$?" >> gpurun_out/c6_bench2.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-rri --no-cpu > gpurun_out/c6_bench1.log 2> gpurun_out/c6_bench1.err; echo "rc=$?" >> gpurun_out/c6_bench1.err
tail -5 gpurun_out/c6_pytest.log; tail -2 gpurun_out/c6_masked.log; cat gpurun_out/c6_init.log | tail -4; for f in c6_bench2 c6_bench1; do tail -2 gpurun_out/$f.err | cut -c1-300; cut -c1-300 gpurun_out/$f.log; done
