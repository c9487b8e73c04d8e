#!/bin/bash
# CTA-pair GEMM bring-up: probe -> self-tests -> per-shape timing (pair vs single) -> training step A/B
mkdir -p gpurun_out
VY_GEMM_PAIR=1 timeout 120 python tools/pair_probe.py > gpurun_out/pair_probe.log 2>&1; rc=$?; echo "pair_probe rc=$rc"; cat gpurun_out/pair_probe.log | tail -30
if [ $rc -ne 0 ]; then exit 0; fi
VY_GEMM_PAIR=1 timeout 600 python tools/gpu_selftest.py --only gemm_bf16,qkv_rope > gpurun_out/selftest_pair.log 2>&1; echo "selftest(pair) rc=$?"
grep -c PASS gpurun_out/selftest_pair.log; grep -E "FAIL|TIMEOUT|SUMMARY" gpurun_out/selftest_pair.log | head -20
VY_GEMM_PAIR=1 timeout 300 python tools/gemm_bench.py --json gpurun_out/gemm_bench_pair.json > gpurun_out/gemm_bench_pair.log 2>&1; echo "gemm_bench(pair) rc=$?"
cat gpurun_out/gemm_bench_pair.log
timeout 300 python tools/gemm_bench.py --json gpurun_out/gemm_bench.json > gpurun_out/gemm_bench.log 2>&1; echo "gemm_bench rc=$?"
cat gpurun_out/gemm_bench.log
VY_GEMM_PAIR=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err; echo "bench(pair) rc=$?"; cut -c1-400 gpurun_out/bench_pair.json
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err; echo "bench(single) rc=$?"; cut -c1-400 gpurun_out/bench_single.json
