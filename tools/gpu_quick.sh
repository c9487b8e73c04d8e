#!/bin/bash
# quick GPU iteration: kernel self-tests of the named groups, parity tests, per-shape GEMM timing
mkdir -p gpurun_out
python tools/gpu_selftest.py $SELFTEST_ARGS > gpurun_out/selftest.log 2>&1; echo "selftest rc=$?"
grep -c PASS gpurun_out/selftest.log; grep FAIL gpurun_out/selftest.log | head -20
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python tools/gemm_bench.py --json gpurun_out/gemm_bench.json > gpurun_out/gemm_bench.log 2>&1; echo "gemm_bench rc=$?"
cat gpurun_out/gemm_bench.log
