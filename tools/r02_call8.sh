#!/bin/bash
# round-2 GPU call 8 (8 GPUs): multi-GPU parity at 8 and 4 ranks (peer exchange and NCCL), config-3 scaling N = 4, 8,
# config 5 (1M x 20k, k = 128) on 8 GPUs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c8_gpu.txt 2>&1; nproc >> gpurun_out/c8_gpu.txt; free -g >> gpurun_out/c8_gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
RRI_P2P=1 timeout 400 $TR --nproc-per-node 8 --master-port 29581 tests/multi_gpu_check.py > gpurun_out/c8_mg8_p2p.log 2>&1; echo "rc=$?" >> gpurun_out/c8_mg8_p2p.log
RRI_P2P=0 timeout 400 $TR --nproc-per-node 8 --master-port 29582 tests/multi_gpu_check.py > gpurun_out/c8_mg8_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/c8_mg8_nccl.log
RRI_P2P=1 timeout 400 $TR --nproc-per-node 4 --master-port 29583 tests/multi_gpu_check.py > gpurun_out/c8_mg4_p2p.log 2>&1; echo "rc=$?" >> gpurun_out/c8_mg4_p2p.log
timeout 500 $TR --nproc-per-node 8 --master-port 29584 bench.py --gpus 8 --steps 20 --warmup 5 --no-pageable > gpurun_out/c8_bench8.log 2> gpurun_out/c8_bench8.err; echo "rc=$?" >> gpurun_out/c8_bench8.err
timeout 500 $TR --nproc-per-node 4 --master-port 29585 bench.py --gpus 4 --steps 20 --warmup 5 --no-pageable > gpurun_out/c8_bench4.log 2> gpurun_out/c8_bench4.err; echo "rc=$?" >> gpurun_out/c8_bench4.err
timeout 600 $TR --nproc-per-node 8 --master-port 29586 bench.py --gpus 8 --config cfg5 --steps 20 --warmup 5 --no-pageable > gpurun_out/c8_cfg5_8.log 2> gpurun_out/c8_cfg5_8.err; echo "rc=$?" >> gpurun_out/c8_cfg5_8.err
RRI_P2P=0 timeout 500 $TR --nproc-per-node 8 --master-port 29587 bench.py --gpus 8 --steps 20 --warmup 5 --no-pageable --no-e2e --no-rri > gpurun_out/c8_bench8_nccl.log 2> gpurun_out/c8_bench8_nccl.err; echo "rc=$?" >> gpurun_out/c8_bench8_nccl.err
for f in c8_mg8_p2p c8_mg8_nccl c8_mg4_p2p; do echo "== $f"; grep -E "ok|FAIL|PARITY|rc=|Error|error" gpurun_out/$f.log | tail -14; done
for f in c8_bench8 c8_bench4 c8_cfg5_8 c8_bench8_nccl; do echo "== $f"; tail -3 gpurun_out/$f.err | cut -c1-300; python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/$f.log') if l.startswith('{')][-1])
    print(j['value'], j['ms_per_step'], j['gpu_launches'], j['config']['final_rel_error'], j['roofline']['frac'], j['roofline'].get('half_steps_ms'), (j.get('e2e') or {}).get('value'))
except Exception as e:
    print('no line', e)
PY
done
