#!/bin/bash
# Round-2 evidence pass: ncu --set full of the attention / decode / norm kernels at the bench shapes
# (each only after the same command exited 0 without ncu), plus the per-kernel sequence of one decode-graph replay.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
CMD="python tools/attn_bench.py"
$CMD > gpurun_out/plain_attn.log 2>&1 && {
  ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 3 -c 1 -o gpurun_out/r02_prof_attnfwd_dec $CMD > gpurun_out/ncu_attn1.log 2>&1; echo "ncu attnfwd dec rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 30 -c 1 -o gpurun_out/r02_prof_attnfwd_vit $CMD > gpurun_out/ncu_attn2.log 2>&1; echo "ncu attnfwd vit rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:attn_bwd_fused -s 3 -c 1 -o gpurun_out/r02_prof_attnbwd_dec $CMD > gpurun_out/ncu_attn3.log 2>&1; echo "ncu attnbwd rc=$?"
}
cat gpurun_out/plain_attn.log
CMD="python tools/decode_bench.py --attn gqa --decode 24"
VY_PROFILE_SEQUENCE=1 $CMD > gpurun_out/plain_dec.log 2> gpurun_out/decode_sequence.txt && \
ncu --set full --clock-control none --import-source on -k regex:attn_decode -s 40 -c 1 -o gpurun_out/r02_prof_decode $CMD > gpurun_out/ncu_dec.log 2>&1
echo "ncu decode rc=$?"; cat gpurun_out/plain_dec.log
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
export VY_GEMM_TUNE_CACHE=gpurun_out/tune_cache.json
$CMD > gpurun_out/plain_b.log 2>&1 && {
  ncu --set full --clock-control none --import-source on -k regex:add_layernorm_fwd -s 60 -c 1 -o gpurun_out/r02_prof_lnfwd $CMD > gpurun_out/ncu_ln1.log 2>&1; echo "ncu lnfwd rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:add_layernorm_bwd_kernel -s 60 -c 1 -o gpurun_out/r02_prof_lnbwd $CMD > gpurun_out/ncu_ln2.log 2>&1; echo "ncu lnbwd rc=$?"
}
ls -la gpurun_out
