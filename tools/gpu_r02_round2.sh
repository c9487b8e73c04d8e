#!/bin/bash
# Round-2 second GPU pass: full parity suite, the bench line (incl. decode + notebook-II), ncu of the new decode kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?"; tail -22 gpurun_out/pytest_gpu3.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench2.json; tail -5 gpurun_out/bench2.err
CMD="python tools/decode_bench.py --attn gqa --decode 16"
$CMD > gpurun_out/plain_dec2.log 2> gpurun_out/plain_dec2.err && \
ncu --set full --clock-control none --import-source on -k regex:attn_decode_tma_kernel -s 8 -c 1 -o gpurun_out/r02_prof_attn_decode_tma $CMD > gpurun_out/ncu_dec_tma.log 2>&1
echo "ncu tma rc=$?"; tail -2 gpurun_out/ncu_dec_tma.log
ncu --set full --clock-control none --import-source on -k regex:gemm_skinny_kernel -s 40 -c 4 -o gpurun_out/r02_prof_gemm_skinny $CMD > gpurun_out/ncu_skinny.log 2>&1
echo "ncu skinny rc=$?"; tail -2 gpurun_out/ncu_skinny.log
