#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_paligemma.py -x -q -m gpu -s > gpurun_out/pytest_pg.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_pg.log
timeout 120 python tools/fa_probe.py 2>&1 | tail -2
