#!/usr/bin/env bash
mkdir -p gpurun_out; rm -f gpurun_out/r02s_sigma.jsonl
timeout 200 python scripts/bench_sigma.py >> gpurun_out/r02s_sigma.jsonl 2>> gpurun_out/r02s_sigma.err
cat gpurun_out/r02s_sigma.jsonl
timeout 600 python -m pytest tests/test_gpu_stout.py tests/test_gpu_callers.py tests/test_gpu_force.py -m gpu -q > gpurun_out/r02s_stout_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02s_stout_tests.log
