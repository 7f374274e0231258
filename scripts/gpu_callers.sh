#!/usr/bin/env bash
# Runs on the B200 box (under gpurun) with a small GPU-minute budget: the new callers' parity tests first, then the
# whole GPU suite, smoke and a short bench; every step tees into gpurun_out/ so that a cut-off call still leaves results.
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_callers.py -x -q -s 2>&1 | tail -40 > gpurun_out/r01d_callers.log; tail -15 gpurun_out/r01d_callers.log
timeout 400 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_callers.py 2>&1 | tail -15 > gpurun_out/r01d_pytest_gpu.log; tail -5 gpurun_out/r01d_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01d_smoke.log 2>&1; tail -2 gpurun_out/r01d_smoke.log
timeout 300 python bench.py > gpurun_out/r01d_bench_n1.json 2> gpurun_out/r01d_bench.err; cat gpurun_out/r01d_bench_n1.json; tail -2 gpurun_out/r01d_bench.err
