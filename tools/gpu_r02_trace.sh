#!/bin/bash
mkdir -p gpurun_out
VY_DECODE_FUSED=1 timeout 300 python tools/decode_bench.py --trace --attn gqa --decode 32 > gpurun_out/trace.json 2> gpurun_out/trace.err; echo "rc=$?"; tail -30 gpurun_out/trace.err
nvidia-smi --query-gpu=clocks.sm,clocks.mem,clocks.max.sm,clocks.max.mem --format=csv
