#!/bin/bash
# round-2 GPU call 12 (1 GPU): fused W half-step (RRI_FUSE_UPDATE=1): tests, bench A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
RRI_FUSE_UPDATE=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c12_pytest_fused.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c12_pytest_fused.log
RRI_FUSE_UPDATE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-rri --no-cpu --no-e2e > gpurun_out/c12_bench_fused.log 2> gpurun_out/c12_bench_fused.err; echo "rc=$?" >> gpurun_out/c12_bench_fused.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-rri --no-cpu --no-e2e > gpurun_out/c12_bench_plain.log 2> gpurun_out/c12_bench_plain.err; echo "rc=$?" >> gpurun_out/c12_bench_plain.err
RRI_FUSE_UPDATE=1 timeout 600 python bench.py --rows 25000 --steps 50 --warmup 5 --no-rri --no-cpu --no-e2e > gpurun_out/c12_bench25k_fused.log 2> gpurun_out/c12_bench25k_fused.err
timeout 600 python bench.py --rows 25000 --steps 50 --warmup 5 --no-rri --no-cpu --no-e2e > gpurun_out/c12_bench25k_plain.log 2> gpurun_out/c12_bench25k_plain.err
tail -4 gpurun_out/c12_pytest_fused.log; for f in c12_bench_fused c12_bench_plain c12_bench25k_fused c12_bench25k_plain; do tail -1 gpurun_out/$f.err | cut -c1-300; python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/$f.log') if l.startswith('{')][-1])
    print('$f', j['value'], j['ms_per_step'], j['gpu_launches'], j['config']['final_rel_error'], j['roofline'].get('half_steps_ms'), (j.get('objective_per_sweep') or {}).get('rel_diff'))
except Exception as e:
    print('$f no line', e)
PY
done
