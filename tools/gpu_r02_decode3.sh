#!/bin/bash
# small-batch GEMM: parity, then the config-3 decode bench A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_paged.py tests/test_gpu_decode_step.py tests/test_gpu_models.py -x -q -m gpu > gpurun_out/pytest_decode3.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_decode3.log
OUT=gpurun_out/decode_gemm_ab.txt
: > $OUT
run() { # label env...
  label=$1; shift
  for attn in gqa mha; do
    env "$@" timeout 300 python tools/decode_bench.py --attn $attn 2>gpurun_out/decode3.err | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$label', '$attn', round(d['graph_decode_us_per_step'],1), 'us/step', round(d['graph_decode_hbm_frac_of_measured'],4), 'ids_match', d.get('ids_match_generate'), 'eager', round(d['decode_us_per_step'],1), flush=True)" | tee -a $OUT
    tail -2 gpurun_out/decode3.err | grep -i "error" | tee -a $OUT
  done
}
run "tcgen05 gemm, PDL" VY_GEMM_SKINNY=0
run "skinny gemm, PDL" X=1
run "skinny gemm, no PDL" VY_DECODE_PDL=0
run "tcgen05 gemm, no PDL" VY_GEMM_SKINNY=0 VY_DECODE_PDL=0
cat $OUT
VY_PROFILE_SEQUENCE=1 timeout 300 python tools/decode_bench.py --attn gqa > /dev/null 2> gpurun_out/decode_sequence2.txt; grep -v Warn gpurun_out/decode_sequence2.txt | head -50
