#!/bin/bash
# round-2 GPU call 28 (2 GPUs): verification of the final tree -- full GPU suite (incl. the 2-rank tests), smoke, default
# bench, reference arm, 2-GPU bench, ncu launch list of the default bench command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/c28_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c28_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/c28_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/c28_bench.log 2> gpurun_out/c28_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/c28_ref.log 2> gpurun_out/c28_ref.err; echo "ref rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c28_bench2.log 2> gpurun_out/c28_bench2.err; echo "bench2 rc=$?"
timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-rri > gpurun_out/c28_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c28_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-rri > gpurun_out/c28_ncu.log 2>&1
tail -3 gpurun_out/c28_pytest.log; tail -2 gpurun_out/c28_smoke.log
for f in c28_bench c28_ref c28_bench2; do grep '^{' gpurun_out/$f.log | cut -c1-330; done
