#!/bin/bash
# launch list of one hals/tf32 sweep at the per-GPU size of an 8-GPU config-3 run (25000 rows)
KREGEX='regex:(tf32_gemm|rri_|update_rows|gram_|reduce_|colsum|transpose|simt_gemm|wrri|objective|norms|project|finalize|flag_|partials)'
CMD="python bench.py --rows 25000 --steps 3 --warmup 2 --no-e2e --no-cpu"
$CMD > gpurun_out/small_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 300 --csv --log-file gpurun_out/small_launches.csv $CMD > gpurun_out/small_ncu.log 2>&1
tail -1 gpurun_out/small_plain.log | cut -c1-600
