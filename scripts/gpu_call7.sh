#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_force.py -m gpu -x -q > gpurun_out/pytest_force.log 2>&1; echo "pytest force rc=$?"; tail -3 gpurun_out/pytest_force.log
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -8 gpurun_out/pytest_multi.log
