#!/bin/bash
# round-2 GPU call 13 (1 GPU): register-resident row update for k = 128: tests, cfg5-shard bench A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c13_pytest.log
timeout 600 python bench.py --config cfg5 --rows 125000 --steps 20 --warmup 5 --no-cpu --no-e2e --no-rri > gpurun_out/c13_cfg5_reg.log 2> gpurun_out/c13_cfg5_reg.err; echo "rc=$?" >> gpurun_out/c13_cfg5_reg.err
RRI_UPDATE_REG=0 timeout 600 python bench.py --config cfg5 --rows 125000 --steps 20 --warmup 5 --no-cpu --no-e2e --no-rri > gpurun_out/c13_cfg5_tpr.log 2> gpurun_out/c13_cfg5_tpr.err; echo "rc=$?" >> gpurun_out/c13_cfg5_tpr.err
tail -4 gpurun_out/c13_pytest.log; for f in c13_cfg5_reg c13_cfg5_tpr; do tail -1 gpurun_out/$f.err | cut -c1-300; python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/$f.log') if l.startswith('{')][-1])
    print('$f', j['value'], j['ms_per_step'], j['gpu_launches'], j['config']['final_rel_error'], j['roofline'].get('half_steps_ms'))
except Exception as e:
    print('$f no line', e)
PY
done
