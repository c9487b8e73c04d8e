#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/decode_skinny_sweep.txt
: > $OUT
run() { # label env...
  label=$1; shift
  for attn in gqa mha; do
    env "$@" timeout 300 python tools/decode_bench.py --attn $attn 2>gpurun_out/decode4.err | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$label', '$attn', round(d['graph_decode_us_per_step'],1), 'us/step', round(d['graph_decode_hbm_frac_of_measured'],4), 'ids_match', d.get('ids_match_generate'), flush=True)" | tee -a $OUT
  done
}
run "ctas 48" VY_SKINNY_CTAS=48
run "ctas 96" VY_SKINNY_CTAS=96
run "ctas 148" VY_SKINNY_CTAS=148
run "ctas 200" VY_SKINNY_CTAS=200
