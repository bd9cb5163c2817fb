#!/bin/bash
# round-2 GPU call 36 (1 GPU): epilogue warps of the contraction parked on the barrier (suspend-time hint) instead of
# polling: power / clocks / sustained sweep time over 100 sweeps
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
RRI_GEMM_PARKED_WAIT=1 timeout 300 python -m pytest tests -m gpu -x -q -k "gemm or cfg3 or full_size" > gpurun_out/c36_pytest.log 2>&1; echo "pytest rc=$?"
for v in 1 0 1 0; do
  RRI_GEMM_PARKED_WAIT=$v timeout 300 python bench.py --steps 150 --warmup 5 --no-rri --no-cpu --no-e2e >> gpurun_out/c36_p$v.log 2>> gpurun_out/c36_p$v.err
done
tail -1 gpurun_out/c36_pytest.log
for v in 1 0; do python - <<PY
import json
for l in open('gpurun_out/c36_p$v.log'):
    if l.startswith('{'):
        j=json.loads(l); h=j['roofline'].get('half_steps_ms')
        print('parked=$v', round(j['value'],2), round(j['ms_per_step'],4), 'gemm_t %.4f gemm_w %.4f t_half %.4f w_half %.4f' % (h['gemm_t'], h['gemm_w'], h['t_half'], h['w_half']), j['clocks'])
PY
done
