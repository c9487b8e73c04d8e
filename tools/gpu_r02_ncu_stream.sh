#!/bin/bash
# ncu --set full of the streaming kernels of the training step that had no capture yet: AdamW, the gradient-norm pass, one
# bias-gradient column sum (FFN-in, 8256 x 3072) — achieved DRAM GB/s against the measured copy peak.
mkdir -p gpurun_out
export VY_GEMM_TUNE_CACHE=gpurun_out/tune_cache_launches.json
CMD="python bench.py --no-graph --steps 1 --warmup 3 --no-decode --no-slots --no-configs-1-2 --no-config5 --no-cpu-baseline"
$CMD > gpurun_out/stream_plain.json 2> gpurun_out/stream_plain.err || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"adamw_kernel|sqnorm_kernel" -s 6 -c 2 -f -o gpurun_out/r02_prof_adamw $CMD > gpurun_out/ncu_adamw.log 2>&1
echo "ncu adamw rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:colsum_partial -s 100 -c 3 -f -o gpurun_out/r02_prof_colsum $CMD > gpurun_out/ncu_colsum.log 2>&1
echo "ncu colsum rc=$?"
