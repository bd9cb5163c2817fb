#!/bin/bash
# round-2 GPU call 29 (1 GPU): default bench after restricting the 128-row-tile rule to small factors; 25k / 50k shards
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/c29_bench.log 2> gpurun_out/c29_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --rows 25000 --steps 100 --warmup 5 --no-rri --no-cpu --no-e2e > gpurun_out/c29_25k.log 2> gpurun_out/c29_25k.err
timeout 600 python bench.py --rows 50000 --steps 60 --warmup 5 --no-rri --no-cpu --no-e2e > gpurun_out/c29_50k.log 2> gpurun_out/c29_50k.err
timeout 600 python -m pytest tests -m gpu -x -q -k "gemm or hals or cfg3 or full_size" > gpurun_out/c29_pytest.log 2>&1; echo "pytest rc=$?"
for f in c29_bench c29_25k c29_50k; do python - <<PY
import json
j=json.loads([l for l in open('gpurun_out/$f.log') if l.startswith('{')][-1])
h=j['roofline'].get('half_steps_ms')
print('$f', round(j['value'],2), round(j['ms_per_step'],4), j['config']['final_rel_error'], 'gemm_t %.4f gemm_w %.4f t_half %.4f w_half %.4f' % (h['gemm_t'], h['gemm_w'], h['t_half'], h['w_half']), j.get('clocks'))
PY
done; tail -1 gpurun_out/c29_pytest.log
