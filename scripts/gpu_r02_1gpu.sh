#!/usr/bin/env bash
# round 2, one GPU: parity suite, driver-contract bench (both arms, CG-M iteration fixture written), launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02a_pytest_gpu.log
timeout 900 python bench.py --write-fixture > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err; echo "bench rc=$?"; cut -c1-3000 gpurun_out/r02a_bench_n1.json; tail -5 gpurun_out/r02a_bench_n1.err
cp tests/golden/bench_cgm_iterations.json gpurun_out/ 2>/dev/null
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02a_bench_reference_n1.json 2> gpurun_out/r02a_bench_ref.err; echo "bench ref rc=$?"; cut -c1-900 gpurun_out/r02a_bench_reference_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:staple:: -s 20 -c 400 --csv \
    --log-file gpurun_out/r02a_launches_bench_config1.csv python bench.py --sections config1 --steps 4 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dslash_kernel -s 12 -c 1 -f -o gpurun_out/r02a_prof_dslash \
    python bench.py --sections config1 --steps 4 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_full_dslash.log 2>&1; echo "ncu dslash rc=$?"
ls -la gpurun_out/
