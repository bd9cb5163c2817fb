#!/bin/bash
# round-2 GPU call 5 (1 GPU): masked kernel v2, sparse variants, objective through the contraction, ncu captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c5_pytest.log
timeout 300 python tools/bench_masked.py 100000 rri tf32 > gpurun_out/c5_masked.log 2>&1
timeout 300 python tools/bench_masked.py 100000 hals tf32 >> gpurun_out/c5_masked.log 2>&1
echo "== default (128 KB blocks, prefetched bounds)" > gpurun_out/c5_sparse.log
timeout 300 python tools/bench_sparse.py 100000 rri 8 >> gpurun_out/c5_sparse.log 2>&1
echo "== RRI_SP_BLOCK_KB=64" >> gpurun_out/c5_sparse.log
RRI_SP_BLOCK_KB=64 timeout 300 python tools/bench_sparse.py 100000 rri 8 >> gpurun_out/c5_sparse.log 2>&1
echo "== RRI_SP_BLOCK_KB=32" >> gpurun_out/c5_sparse.log
RRI_SP_BLOCK_KB=32 timeout 300 python tools/bench_sparse.py 100000 rri 8 >> gpurun_out/c5_sparse.log 2>&1
echo "== refresh every 8 sweeps" >> gpurun_out/c5_sparse.log
timeout 300 python tools/bench_sparse.py 100000 rri 8 8 >> gpurun_out/c5_sparse.log 2>&1
echo "== RRI_SP_BLOCK_KB=64, refresh every 8 sweeps" >> gpurun_out/c5_sparse.log
RRI_SP_BLOCK_KB=64 timeout 300 python tools/bench_sparse.py 100000 rri 8 8 >> gpurun_out/c5_sparse.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c5_bench1.log 2> gpurun_out/c5_bench1.err; echo "rc=$?" >> gpurun_out/c5_bench1.err
python tools/bench_masked.py 20000 rri tf32 > gpurun_out/c5_m_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wrri_tc_tma -s 20 -c 2 -o gpurun_out/r02_wrri_tma_v2 python tools/bench_masked.py 20000 rri tf32 > gpurun_out/c5_m_ncu.log 2>&1
python tools/bench_sparse.py 100000 rri 1 > gpurun_out/c5_s_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sp_pass_blocked -s 40 -c 2 -o gpurun_out/r02_sp_pass python tools/bench_sparse.py 100000 rri 1 > gpurun_out/c5_s_ncu.log 2>&1
tail -4 gpurun_out/c5_pytest.log; cat gpurun_out/c5_masked.log | tail -3; grep -v "Warn\|Xs = " gpurun_out/c5_sparse.log | cut -c1-330; cut -c1-200 gpurun_out/c5_bench1.log
