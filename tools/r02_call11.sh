#!/bin/bash
# round-2 GPU call 11 (8 GPUs, final code): parity at 8 and 4 ranks with the two-kernel exchange, config-3 scaling N = 8, 4,
# config 5 on 8 GPUs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
RRI_P2P=1 timeout 400 $TR --nproc-per-node 8 --master-port 29581 tests/multi_gpu_check.py > gpurun_out/c11_mg8_p2p.log 2>&1; echo "rc=$?" >> gpurun_out/c11_mg8_p2p.log
RRI_P2P=1 timeout 400 $TR --nproc-per-node 4 --master-port 29583 tests/multi_gpu_check.py > gpurun_out/c11_mg4_p2p.log 2>&1; echo "rc=$?" >> gpurun_out/c11_mg4_p2p.log
timeout 500 $TR --nproc-per-node 8 --master-port 29584 bench.py --gpus 8 --steps 20 --warmup 5 --no-pageable > gpurun_out/c11_bench8.log 2> gpurun_out/c11_bench8.err; echo "rc=$?" >> gpurun_out/c11_bench8.err
timeout 500 $TR --nproc-per-node 4 --master-port 29585 bench.py --gpus 4 --steps 20 --warmup 5 --no-pageable > gpurun_out/c11_bench4.log 2> gpurun_out/c11_bench4.err; echo "rc=$?" >> gpurun_out/c11_bench4.err
timeout 600 $TR --nproc-per-node 8 --master-port 29586 bench.py --gpus 8 --config cfg5 --steps 20 --warmup 5 --no-pageable > gpurun_out/c11_cfg5_8.log 2> gpurun_out/c11_cfg5_8.err; echo "rc=$?" >> gpurun_out/c11_cfg5_8.err
timeout 500 $TR --nproc-per-node 8 --master-port 29587 bench.py --gpus 8 --steps 100 --warmup 5 --no-pageable --no-e2e --no-rri > gpurun_out/c11_bench8_100.log 2> gpurun_out/c11_bench8_100.err; echo "rc=$?" >> gpurun_out/c11_bench8_100.err
for f in c11_mg8_p2p c11_mg4_p2p; do echo "== $f"; grep -E "ok|FAIL|PARITY|rc=|Error|error" gpurun_out/$f.log | grep -v Warn | tail -12; done
for f in c11_bench8 c11_bench4 c11_cfg5_8 c11_bench8_100; do echo "== $f"; tail -2 gpurun_out/$f.err | cut -c1-300; python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/$f.log') if l.startswith('{')][-1])
    print(j['value'], j['ms_per_step'], j['gpu_launches'], j['config']['final_rel_error'], j['roofline']['frac'], j['roofline'].get('half_steps_ms'), (j.get('e2e') or {}).get('value'))
except Exception as e:
    print('no line', e)
PY
done
