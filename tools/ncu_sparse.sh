#!/bin/bash
# ncu evidence for the observed-entries (sparse) path at config-4 shape (run under gpurun, 1 GPU, ONE ncu per call).
#   tools/ncu_sparse.sh list r02    launch list (gpu__time_duration.sum) of one rri-order sweep
#   tools/ncu_sparse.sh full r02    ncu --set full capture of two launches of the pass kernel (sp_pass_*_kernel)
# The ncu run is preceded, in the same call, by the identical command without ncu (must exit 0).
# Read the results here with:  python tools/launch_summary.py gpurun_out/<r>_sparse_launches.csv rri::
#                              ncu -i gpurun_out/<r>_sp_pass.ncu-rep --page raw --csv
MODE=${1:-list}
R=${2:-r02}
CMD="python tools/bench_sparse.py 100000 rri 1"
mkdir -p gpurun_out
if [ "$MODE" = list ]; then
  $CMD > gpurun_out/${R}_sparse_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/${R}_sparse_launches.csv \
      $CMD > gpurun_out/${R}_sparse_ncu.log 2>&1
else
  $CMD > gpurun_out/${R}_sparse_plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:sp_pass_ -s 110 -c 2 -f \
      -o gpurun_out/${R}_sp_pass $CMD > gpurun_out/${R}_sparse_ncu_full.log 2>&1
fi
tail -1 gpurun_out/${R}_sparse_plain*.log | cut -c1-400
