#!/bin/bash
# ncu --set full of the vy_gemm launches of one gemm_bench case: bash tools/ncu_gemm.sh "<case substring>" <out name>
mkdir -p gpurun_out
export VY_GEMM_AUTOTUNE=0  # profile the library's own tiling choice, without the tuner's trial launches
CMD="python tools/gemm_bench.py --iters 2 --only $1"
$CMD > gpurun_out/plain_$2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 2 -o gpurun_out/prof_$2 $CMD > gpurun_out/ncu_$2.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$2.log
