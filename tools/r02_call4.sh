#!/bin/bash
# round-2 GPU call 4 (1 GPU): estimator tests, H2D probe, TF32 flush sweep, ncu full captures (masked kernel, W update)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
timeout 300 python tools/h2d_probe.py 8 > gpurun_out/c4_h2d.log 2>&1
for f in 1 2 4 8 32; do
  echo "== RRI_GEMM_FLUSH=$f" >> gpurun_out/c4_flush.log
  RRI_GEMM_FLUSH=$f timeout 300 python tools/tf32_bias.py 2>&1 | grep -E "K=200000|K=20000 " >> gpurun_out/c4_flush.log
  RRI_GEMM_FLUSH=$f timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-rri 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('sweeps/s', j['value'], 'kernel_ms', j['roofline']['kernel_ms'])" >> gpurun_out/c4_flush.log
done
python tools/bench_masked.py 20000 rri tf32 > gpurun_out/c4_m_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wrri_tc_tma -s 20 -c 4 -o gpurun_out/r02_wrri_tma python tools/bench_masked.py 20000 rri tf32 > gpurun_out/c4_m_ncu.log 2>&1
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-rri > gpurun_out/c4_u_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:update_rows_tpr -s 2 -c 2 -o gpurun_out/r02_update_rows python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-rri > gpurun_out/c4_u_ncu.log 2>&1
tail -4 gpurun_out/c4_pytest.log; cat gpurun_out/c4_h2d.log; cat gpurun_out/c4_flush.log; ls -la gpurun_out/*.ncu-rep
