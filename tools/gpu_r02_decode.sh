#!/bin/bash
# fused decode step: parity tests, then config 3 decode timing with the one-kernel step (stage trace on stderr) and with the per-op graph
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode_step.py tests/test_gpu_real_shapes.py -m gpu -q -x > gpurun_out/pytest_decode.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_decode.log
for attn in gqa mha; do
  VY_DECODE_FUSED=1 timeout 300 python tools/decode_bench.py --trace --attn $attn > gpurun_out/decode_fused_$attn.json 2> gpurun_out/decode_fused_$attn.err; echo "fused $attn rc=$?"; cat gpurun_out/decode_fused_$attn.json
  if [ "$1" = "both" ]; then
  VY_DECODE_FUSED=0 timeout 300 python tools/decode_bench.py --attn $attn > gpurun_out/decode_perop_$attn.json 2> gpurun_out/decode_perop_$attn.err; echo "per-op $attn rc=$?"; cat gpurun_out/decode_perop_$attn.json
  fi
done
