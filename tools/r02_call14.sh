#!/bin/bash
# round-2 GPU call 14 (1 GPU): register-resident row update at every rank: tests, cfg3 / cfg5-shard / small-shard bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c14_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c14_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-rri --no-cpu --no-e2e > gpurun_out/c14_bench.log 2> gpurun_out/c14_bench.err; echo "rc=$?" >> gpurun_out/c14_bench.err
RRI_UPDATE_REG=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-rri --no-cpu --no-e2e > gpurun_out/c14_bench_tpr.log 2> gpurun_out/c14_bench_tpr.err; echo "rc=$?" >> gpurun_out/c14_bench_tpr.err
timeout 600 python bench.py --rows 25000 --steps 50 --warmup 5 --no-rri --no-cpu --no-e2e > gpurun_out/c14_bench25k.log 2> gpurun_out/c14_bench25k.err
timeout 600 python bench.py --config cfg5 --rows 125000 --steps 20 --warmup 5 --no-cpu --no-e2e --no-rri > gpurun_out/c14_cfg5.log 2> gpurun_out/c14_cfg5.err
timeout 300 python bench.py --config cfg2 --order hals --steps 20 --warmup 3 --no-rri --no-cpu --no-e2e > gpurun_out/c14_cfg2.log 2> gpurun_out/c14_cfg2.err
tail -4 gpurun_out/c14_pytest.log; for f in c14_bench c14_bench_tpr c14_bench25k c14_cfg5 c14_cfg2; do tail -1 gpurun_out/$f.err | cut -c1-300; python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/$f.log') if l.startswith('{')][-1])
    print('$f', j['value'], j['ms_per_step'], j['gpu_launches'], j['config']['final_rel_error'], j['roofline'].get('half_steps_ms'))
except Exception as e:
    print('$f no line', e)
PY
done
