#!/bin/bash
# round-2 GPU call 3 (1 GPU): new masked kernel (tests + cfg4 timing), half-step breakdown, launch lists, H2D probe
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
timeout 300 python tools/bench_masked.py 100000 rri tf32 > gpurun_out/c3_masked.log 2>&1
timeout 300 python tools/bench_masked.py 100000 hals tf32 >> gpurun_out/c3_masked.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-pageable --no-cpu --no-rri > gpurun_out/c3_bench1.log 2> gpurun_out/c3_bench1.err; echo "rc=$?" >> gpurun_out/c3_bench1.err
timeout 300 python tools/h2d_probe.py 8 > gpurun_out/c3_h2d.log 2>&1
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-rri > gpurun_out/c3_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c3_hals_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-rri > gpurun_out/c3_ncu.log 2>&1
tail -4 gpurun_out/c3_pytest.log; cat gpurun_out/c3_masked.log | tail -4; cut -c1-300 gpurun_out/c3_bench1.log; cat gpurun_out/c3_h2d.log
