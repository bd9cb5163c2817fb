#!/bin/bash
# round-2 GPU call 35 (1 GPU): the driver's round-end sequence on the final, cleanly rebuilt tree
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/c35_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c35_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/c35_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/c35_bench.log 2> gpurun_out/c35_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/c35_pytest.log; tail -1 gpurun_out/c35_smoke.log; grep '^{' gpurun_out/c35_bench.log | cut -c1-250
