#!/bin/bash
# round-2 GPU call 33 (1 GPU): fp32 row update with eight accumulators (RRI_UPDATE_ACC8=1) vs four (default)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
RRI_UPDATE_ACC8=1 timeout 600 python -m pytest tests -m gpu -x -q -k "hals or cfg3 or cfg5 or full_size or tf32" > gpurun_out/c33_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c33_pytest.log
for v in 1 0; do
  RRI_UPDATE_ACC8=$v timeout 600 python bench.py --rows 25000 --steps 100 --warmup 5 --no-rri --no-cpu --no-e2e > gpurun_out/c33_25k_a$v.log 2> gpurun_out/c33_25k_a$v.err
  RRI_UPDATE_ACC8=$v timeout 600 python bench.py --steps 40 --warmup 5 --no-rri --no-cpu --no-e2e > gpurun_out/c33_cfg3_a$v.log 2> gpurun_out/c33_cfg3_a$v.err
done
tail -3 gpurun_out/c33_pytest.log; for f in c33_25k_a1 c33_25k_a0 c33_cfg3_a1 c33_cfg3_a0; do python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/$f.log') if l.startswith('{')][-1])
    h=j['roofline'].get('half_steps_ms')
    print('$f', round(j['value'],2), round(j['ms_per_step'],4), j['config']['final_rel_error'], 'gemm_t %.4f gemm_w %.4f t_half %.4f w_half %.4f' % (h['gemm_t'], h['gemm_w'], h['t_half'], h['w_half']))
except Exception as e:
    print('$f no line', e)
PY
done
