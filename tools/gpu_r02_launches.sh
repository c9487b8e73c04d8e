#!/bin/bash
# ncu launch list (gpu__time_duration only, no clock control) of two eager training steps of the final code: every kernel of the
# step. The plain run fills the GEMM tune cache, so the profiled run launches no tuning kernels.
mkdir -p gpurun_out
export VY_GEMM_TUNE_CACHE=gpurun_out/tune_cache_launches.json
rm -f $VY_GEMM_TUNE_CACHE
CMD="python bench.py --no-graph --steps 1 --warmup 3 --no-decode --no-slots --no-configs-1-2 --no-config5 --no-cpu-baseline"
$CMD > gpurun_out/launches_plain.json 2> gpurun_out/launches_plain.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1400 --launch-count 860 --csv --log-file gpurun_out/r02_launches_n1_nograph.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_launches.log | cut -c1-200; wc -l gpurun_out/r02_launches_n1_nograph.csv
