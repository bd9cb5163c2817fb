#!/bin/bash
# round-2 GPU call 26 (1 GPU): row update with vectorised, prefetched contraction entries (RRI_UPDATE_PF=1, default) vs
# scalar loads (=0); contraction tile height on a 25 000-row shard (RRI_GEMM_MT=1 vs default 2)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "hals or cfg3 or cfg5 or cfg2 or full_size or tf32" > gpurun_out/c26_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c26_pytest.log
for v in 1 0; do
  RRI_UPDATE_PF=$v timeout 600 python bench.py --rows 25000 --steps 100 --warmup 5 --no-rri --no-cpu --no-e2e > gpurun_out/c26_25k_pf$v.log 2> gpurun_out/c26_25k_pf$v.err
  RRI_UPDATE_PF=$v timeout 600 python bench.py --steps 40 --warmup 5 --no-rri --no-cpu --no-e2e > gpurun_out/c26_cfg3_pf$v.log 2> gpurun_out/c26_cfg3_pf$v.err
  RRI_UPDATE_PF=$v timeout 600 python bench.py --config cfg5 --rows 125000 --steps 20 --warmup 5 --no-cpu --no-e2e --no-rri > gpurun_out/c26_cfg5_pf$v.log 2> gpurun_out/c26_cfg5_pf$v.err
  RRI_UPDATE_PF=$v timeout 600 python bench.py --config cfg2 --steps 40 --warmup 5 --no-cpu --no-e2e --no-rri > gpurun_out/c26_cfg2_pf$v.log 2> gpurun_out/c26_cfg2_pf$v.err
done
RRI_GEMM_MT=1 timeout 600 python bench.py --rows 25000 --steps 100 --warmup 5 --no-rri --no-cpu --no-e2e > gpurun_out/c26_25k_mt1.log 2> gpurun_out/c26_25k_mt1.err
tail -3 gpurun_out/c26_pytest.log; for f in c26_25k_pf1 c26_25k_pf0 c26_25k_mt1 c26_cfg3_pf1 c26_cfg3_pf0 c26_cfg5_pf1 c26_cfg5_pf0 c26_cfg2_pf1 c26_cfg2_pf0; do python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/$f.log') if l.startswith('{')][-1])
    h=j['roofline'].get('half_steps_ms')
    print('$f', round(j['value'],2), round(j['ms_per_step'],4), j['config']['final_rel_error'], 'gemm_t %.4f gemm_w %.4f t_half %.4f w_half %.4f' % (h['gemm_t'], h['gemm_w'], h['t_half'], h['w_half']))
except Exception as e:
    print('$f no line', e)
PY
done
