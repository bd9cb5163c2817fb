#!/bin/bash
# round-2 GPU call 27 (1 GPU): contraction tile height (RRI_GEMM_MT=1: 128-row tiles, default 2: 256-row super-tiles)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for mt in 2 1; do
  RRI_GEMM_MT=$mt timeout 600 python bench.py --steps 40 --warmup 5 --no-rri --no-cpu --no-e2e > gpurun_out/c27_cfg3_mt$mt.log 2> gpurun_out/c27_cfg3_mt$mt.err
  RRI_GEMM_MT=$mt timeout 600 python bench.py --rows 50000 --steps 60 --warmup 5 --no-rri --no-cpu --no-e2e > gpurun_out/c27_50k_mt$mt.log 2> gpurun_out/c27_50k_mt$mt.err
  RRI_GEMM_MT=$mt timeout 600 python bench.py --rows 25000 --steps 100 --warmup 5 --no-rri --no-cpu --no-e2e > gpurun_out/c27_25k_mt$mt.log 2> gpurun_out/c27_25k_mt$mt.err
  RRI_GEMM_MT=$mt timeout 600 python bench.py --config cfg5 --rows 125000 --steps 20 --warmup 5 --no-cpu --no-e2e --no-rri > gpurun_out/c27_cfg5_mt$mt.log 2> gpurun_out/c27_cfg5_mt$mt.err
done
for f in c27_cfg3_mt2 c27_cfg3_mt1 c27_50k_mt2 c27_50k_mt1 c27_25k_mt2 c27_25k_mt1 c27_cfg5_mt2 c27_cfg5_mt1; do python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/$f.log') if l.startswith('{')][-1])
    h=j['roofline'].get('half_steps_ms')
    print('$f', round(j['value'],2), round(j['ms_per_step'],4), j['config']['final_rel_error'], 'gemm_t %.4f gemm_w %.4f t_half %.4f w_half %.4f' % (h['gemm_t'], h['gemm_w'], h['t_half'], h['w_half']))
except Exception as e:
    print('$f no line', e)
PY
done
