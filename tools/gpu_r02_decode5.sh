#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_paged.py tests/test_gpu_models.py -x -q -m gpu -k "decode or paged or generate" > gpurun_out/pytest_decode5.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_decode5.log
for attn in gqa mha; do
  timeout 300 python tools/decode_bench.py --attn $attn 2>gpurun_out/decode5.err | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$attn', round(d['graph_decode_us_per_step'],1), 'us/step', round(d['graph_decode_hbm_frac_of_measured'],4), 'ids_match', d.get('ids_match_generate'), flush=True)"
done
