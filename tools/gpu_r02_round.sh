#!/bin/bash
# Round-2 GPU pass: full parity suite (incl. the reference's own test files), the bench line, the same-box GPU baseline.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
if [ "$1" != "testsonly" ]; then
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python tools/ref_gpu_bench.py --steps 20 > gpurun_out/ref_gpu.log 2>&1; echo "ref_gpu rc=$?"; tail -3 gpurun_out/ref_gpu.log
python tools/ref_gpu_bench.py --steps 20 --dropout 0.1 > gpurun_out/ref_gpu_p01.log 2>&1; echo "ref_gpu p0.1 rc=$?"; tail -3 gpurun_out/ref_gpu_p01.log
fi
