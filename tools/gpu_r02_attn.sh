#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attn_fwd or attn_bwd" > gpurun_out/pytest_attn.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_attn.log
python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1; cat gpurun_out/attn_bench.log
