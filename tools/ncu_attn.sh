#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/attn_bench.py"
$CMD > gpurun_out/plain_attn.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_fused -s 3 -c 1 -o gpurun_out/prof_attnbwd $CMD > gpurun_out/ncu_attn.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_attn.log
