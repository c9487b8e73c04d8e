#!/bin/bash
# decode attention through the copy engine: parity, then the config-3 decode bench A/B (LDG kernel, split / stage choices, PDL)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_paged.py tests/test_gpu_decode_step.py -x -q -m gpu -k "decode or paged" > gpurun_out/pytest_decode2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_decode2.log
OUT=gpurun_out/decode_attn_ab.txt
: > $OUT
run() { # label env...
  label=$1; shift
  for attn in gqa mha; do
    env "$@" timeout 300 python tools/decode_bench.py --attn $attn 2>gpurun_out/decode2.err | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$label', '$attn', round(d['graph_decode_us_per_step'],1), 'us/step', round(d['graph_decode_hbm_frac_of_measured'],4), 'ids_match', d.get('ids_match_generate'), flush=True)" | tee -a $OUT
  done
}
run "ldg-kernel" VY_DECODE_ATTN_LDG=1
run "tma default" X=1
run "tma splits1 stages6" VY_DECODE_TMA_SPLITS=1 VY_DECODE_TMA_STAGES=6
run "tma splits1 stages3" VY_DECODE_TMA_SPLITS=1 VY_DECODE_TMA_STAGES=3
run "tma splits2 stages3" VY_DECODE_TMA_SPLITS=2 VY_DECODE_TMA_STAGES=3
run "tma splits3 stages2" VY_DECODE_TMA_SPLITS=3 VY_DECODE_TMA_STAGES=2
run "tma default + PDL" VY_PDL=1
run "ldg + PDL" VY_PDL=1 VY_DECODE_ATTN_LDG=1
cat $OUT
