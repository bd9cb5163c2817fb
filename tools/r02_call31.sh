#!/bin/bash
# round-2 GPU call 31 (2 GPUs): verification of the final tree -- full GPU suite (incl. the 2-rank tests), smoke, default
# bench, 2-GPU bench, ncu launch list of the bench command, ncu --set full of the row-update kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/c31_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c31_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/c31_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/c31_bench.log 2> gpurun_out/c31_bench.err; echo "bench rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c31_bench2.log 2> gpurun_out/c31_bench2.err; echo "bench2 rc=$?"
timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-rri > gpurun_out/c31_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c31_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-rri > gpurun_out/c31_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:update_rows_reg -s 4 -c 2 -f -o gpurun_out/r02_update_rows_x2 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-rri > gpurun_out/c31_ncu_full.log 2>&1
tail -3 gpurun_out/c31_pytest.log; tail -2 gpurun_out/c31_smoke.log
for f in c31_bench c31_bench2; do grep '^{' gpurun_out/$f.log | cut -c1-200; done
