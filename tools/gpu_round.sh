#!/bin/bash
# One gpurun call: GPU parity tests, per-shape GEMM timing, the bench line, then the ncu launch list and a full
# capture of the dominant kernel (each ncu pass only after the same command exited 0 without ncu).
mkdir -p gpurun_out
if [ "$1" != "ncuonly" ]; then
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
python tools/gemm_bench.py --json gpurun_out/gemm_bench.json > gpurun_out/gemm_bench.log 2>&1; echo "gemm_bench rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.log
fi
if [ "$1" = "ncu" ] || [ "$1" = "ncuonly" ]; then
  export VY_GEMM_TUNE_CACHE=gpurun_out/tune_cache.json  # the plain run tunes and fills it; the profiler passes launch no tuning kernels
  rm -f $VY_GEMM_TUNE_CACHE
  CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
  $CMD > gpurun_out/plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
  echo "ncu list rc=$?"
  $CMD > gpurun_out/plain2.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 460 -c 6 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"
fi
