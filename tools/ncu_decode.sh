#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/decode_bench.py --attn gqa --decode 24"
$CMD > gpurun_out/plain_dec.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_decode -s 40 -c 1 -o gpurun_out/prof_decode $CMD > gpurun_out/ncu_dec.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_dec.log
