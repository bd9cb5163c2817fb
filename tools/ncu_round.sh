#!/bin/bash
# ncu evidence for one round (run under gpurun, 1 GPU).  Usage: tools/ncu_round.sh r01
# Each ncu run is preceded, in the same call, by the identical command without ncu (must exit 0).
R=${1:-r01}
KREGEX='regex:(tf32_gemm|rri_|update_rows|gram_|reduce_|colsum|transpose|simt_gemm|wrri|objective|norms|project|finalize|flag_|partials)'
HALS="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-rri"
RRI="python bench.py --order rri --steps 1 --warmup 1 --no-e2e --no-cpu"
mkdir -p gpurun_out
$HALS > gpurun_out/${R}_hals_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 400 --csv --log-file gpurun_out/${R}_hals_launches.csv $HALS > gpurun_out/${R}_hals_ncu.log 2>&1
$HALS > gpurun_out/${R}_hals_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tf32_gemm_kernel -s 2 -c 2 -f -o gpurun_out/${R}_tf32_gemm $HALS > gpurun_out/${R}_hals_ncu_full.log 2>&1
$RRI > gpurun_out/${R}_rri_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 600 --csv --log-file gpurun_out/${R}_rri_launches.csv $RRI > gpurun_out/${R}_rri_ncu.log 2>&1
$RRI > gpurun_out/${R}_rri_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rri_pass_kernel -s 3 -c 2 -f -o gpurun_out/${R}_rri_pass $RRI > gpurun_out/${R}_rri_ncu_full.log 2>&1
ls -la gpurun_out | tail -20
