#!/bin/bash
# round-2 GPU call 10 (1 GPU): full tests, smoke, config-4 bench lines (dense / sparse), final launch list and
# ncu --set full of the contraction and of the final masked kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c10_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c10_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/c10_smoke.log
timeout 900 python bench.py --config cfg4 --masked dense --steps 5 --warmup 1 > gpurun_out/c10_cfg4_dense.log 2> gpurun_out/c10_cfg4_dense.err; echo "rc=$?" >> gpurun_out/c10_cfg4_dense.err
timeout 900 python bench.py --config cfg4 --masked sparse --steps 8 --warmup 1 --no-cpu > gpurun_out/c10_cfg4_sparse.log 2> gpurun_out/c10_cfg4_sparse.err; echo "rc=$?" >> gpurun_out/c10_cfg4_sparse.err
timeout 900 python bench.py --config cfg4 --masked sparse --steps 8 --warmup 1 --no-cpu --no-e2e --refresh-every 8 > gpurun_out/c10_cfg4_sparse_r8.log 2> gpurun_out/c10_cfg4_sparse_r8.err; echo "rc=$?" >> gpurun_out/c10_cfg4_sparse_r8.err
HALS="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-rri"
$HALS > gpurun_out/c10_hals_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_hals_launches_final.csv $HALS > gpurun_out/c10_hals_ncu.log 2>&1
$HALS > gpurun_out/c10_hals_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tf32_gemm_kernel -s 2 -c 2 -f -o gpurun_out/r02_tf32_gemm $HALS > gpurun_out/c10_hals_ncu_full.log 2>&1
python tools/bench_masked.py 20000 rri tf32 > gpurun_out/c10_m_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wrri_tc_tma -s 20 -c 2 -f -o gpurun_out/r02_wrri_tma_v3 python tools/bench_masked.py 20000 rri tf32 > gpurun_out/c10_m_ncu.log 2>&1
tail -4 gpurun_out/c10_pytest.log; tail -2 gpurun_out/c10_smoke.log; for f in c10_cfg4_dense c10_cfg4_sparse c10_cfg4_sparse_r8; do tail -2 gpurun_out/$f.err | cut -c1-300; cut -c1-250 gpurun_out/$f.log; done
