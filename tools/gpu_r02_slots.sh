#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_slot_captioner.py -x -q -m gpu -s > gpurun_out/pytest_slots.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_slots.log
