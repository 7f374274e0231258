#!/usr/bin/env bash
# round 2, one GPU, final: smoke, parity suite, driver-contract bench (both arms)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02q_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02q_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02q_bench_n1.json 2> gpurun_out/r02q_bench_n1.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02q_bench_n1.json
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02q_bench_reference_n1.json 2> gpurun_out/r02q_bench_ref.err; echo "bench ref rc=$?"; cut -c1-200 gpurun_out/r02q_bench_reference_n1.json
