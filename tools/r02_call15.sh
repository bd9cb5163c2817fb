#!/bin/bash
# round-2 GPU call 15 (1 GPU): A/B of the reversed K order of split-tile tails in the contraction
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
RRI_GEMM_REVERSE_TAIL=1 timeout 600 python -m pytest tests -m gpu -x -q -k "gemm or hals or cfg3 or cfg5 or full_size or tf32" > gpurun_out/c15_pytest_rev.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c15_pytest_rev.log
for v in 1 0; do
  RRI_GEMM_REVERSE_TAIL=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-rri --no-cpu --no-e2e > gpurun_out/c15_cfg3_rev$v.log 2> gpurun_out/c15_cfg3_rev$v.err
  RRI_GEMM_REVERSE_TAIL=$v timeout 600 python bench.py --config cfg5 --rows 125000 --steps 20 --warmup 5 --no-cpu --no-e2e --no-rri > gpurun_out/c15_cfg5_rev$v.log 2> gpurun_out/c15_cfg5_rev$v.err
  RRI_GEMM_REVERSE_TAIL=$v timeout 600 python bench.py --rows 25000 --steps 50 --warmup 5 --no-rri --no-cpu --no-e2e > gpurun_out/c15_25k_rev$v.log 2> gpurun_out/c15_25k_rev$v.err
done
tail -3 gpurun_out/c15_pytest_rev.log; for f in c15_cfg3_rev1 c15_cfg3_rev0 c15_cfg5_rev1 c15_cfg5_rev0 c15_25k_rev1 c15_25k_rev0; do python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/$f.log') if l.startswith('{')][-1])
    h=j['roofline'].get('half_steps_ms')
    print('$f', round(j['value'],2), round(j['ms_per_step'],4), j['config']['final_rel_error'], 'gemm_t %.4f gemm_w %.4f t_half %.4f w_half %.4f' % (h['gemm_t'], h['gemm_w'], h['t_half'], h['w_half']))
except Exception as e:
    print('$f no line', e)
PY
done
