#!/bin/bash
# DRAM traffic of vy_gemm over TWO WHOLE training steps (350 consecutive launches = every shape of the step twice), eager launches
mkdir -p gpurun_out
CMD="python bench.py --no-graph --steps 10 --warmup 3 --no-decode --no-slots --no-configs-1-2 --no-config5 --no-cpu-baseline"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_kernel --launch-skip 3000 --launch-count 350 --csv --log-file gpurun_out/r02_gemm_traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_traffic.log | cut -c1-300; wc -l gpurun_out/r02_gemm_traffic.csv
