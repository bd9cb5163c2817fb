#!/bin/bash
# round-2 GPU call 9 (2 GPUs): Gram pre-kernel + 32-row exchange blocks: parity, bench N=2 (cfg3, cfg5-shape k=128)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c9_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-pageable --no-rri --no-e2e > gpurun_out/c9_bench2.log 2> gpurun_out/c9_bench2.err; echo "rc=$?" >> gpurun_out/c9_bench2.err
timeout 600 $TR --nproc-per-node 2 --master-port 29534 bench.py --gpus 2 --config cfg5 --rows 250000 --steps 20 --warmup 5 --no-pageable --no-rri --no-e2e > gpurun_out/c9_cfg5_2.log 2> gpurun_out/c9_cfg5_2.err; echo "rc=$?" >> gpurun_out/c9_cfg5_2.err
tail -5 gpurun_out/c9_pytest.log; for f in c9_bench2 c9_cfg5_2; do tail -2 gpurun_out/$f.err | cut -c1-300; python - <<PY
import json
j=json.loads([l for l in open('gpurun_out/$f.log') if l.startswith('{')][-1])
print('$f', j['value'], j['ms_per_step'], j['gpu_launches'], j['roofline']['half_steps_ms'], j['roofline']['frac'], j['config']['final_rel_error'])
PY
done
