#!/bin/bash
# round-2 GPU call 7 (2 GPUs): 4-threads-per-row update kernels: all tests, bench N=1 / N=2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c7_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-pageable --no-rri --no-e2e > gpurun_out/c7_bench2.log 2> gpurun_out/c7_bench2.err; echo "rc=$?" >> gpurun_out/c7_bench2.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-rri --no-cpu --no-e2e > gpurun_out/c7_bench1.log 2> gpurun_out/c7_bench1.err; echo "rc=$?" >> gpurun_out/c7_bench1.err
timeout 600 python bench.py --config cfg5 --rows 125000 --steps 20 --warmup 5 --no-cpu --no-e2e --no-rri > gpurun_out/c7_cfg5shard.log 2> gpurun_out/c7_cfg5shard.err; echo "rc=$?" >> gpurun_out/c7_cfg5shard.err
tail -5 gpurun_out/c7_pytest.log; for f in c7_bench2 c7_bench1 c7_cfg5shard; do tail -2 gpurun_out/$f.err | cut -c1-300; python - <<PY
import json
j=json.loads([l for l in open('gpurun_out/$f.log') if l.startswith('{')][-1])
print('$f', j['value'], j['ms_per_step'], j['roofline']['half_steps_ms'], j['roofline']['frac'])
PY
done
