#!/bin/bash
mkdir -p gpurun_out
export VY_DECODE_FUSED=1
CMD="python tools/decode_bench.py --attn gqa --decode 16 --trace"
$CMD > gpurun_out/plain_fused.log 2> gpurun_out/plain_fused.err && \
ncu --set full --clock-control none --import-source on -k regex:decode_step_kernel -s 6 -c 1 -o gpurun_out/r02_prof_decode_step $CMD > gpurun_out/ncu_fused.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_fused.log; cat gpurun_out/plain_fused.log; tail -24 gpurun_out/plain_fused.err
