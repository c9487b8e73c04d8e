#!/bin/bash
mkdir -p gpurun_out
python tools/xent_bench.py > gpurun_out/xent_bench.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:xent_colsum -s 3 -c 1 -o gpurun_out/r02_prof_xent_colsum -f python tools/xent_bench.py 8256 50265 3 > gpurun_out/ncu_xent.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_xent.log; cat gpurun_out/xent_bench.txt
