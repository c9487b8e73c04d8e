#!/bin/bash
# Final validation of the round: whole -m gpu suite, smoke(), the default bench line, the xent microbenchmark + its ncu capture.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_gpu_final3.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final3.log 2>&1
python bench.py > gpurun_out/bench_final4.json 2> gpurun_out/bench_final4.err
python tools/xent_bench.py > gpurun_out/xent_bench.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:xent_colsum -s 3 -c 1 -o gpurun_out/r02_prof_xent_colsum -f python tools/xent_bench.py 8256 50265 3 > gpurun_out/ncu_xent.log 2>&1
tail -3 gpurun_out/pytest_gpu_final3.log; tail -2 gpurun_out/smoke_final3.log; cat gpurun_out/xent_bench.txt; tail -c 300 gpurun_out/bench_final4.err
