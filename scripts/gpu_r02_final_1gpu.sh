#!/usr/bin/env bash
# round 2, one GPU, final: parity suite, driver-contract bench (both arms), launch list and full ncu capture of the dominant kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02k_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02k_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02k_bench_n1.json 2> gpurun_out/r02k_bench_n1.err; echo "bench rc=$?"; cut -c1-1200 gpurun_out/r02k_bench_n1.json
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02k_bench_reference_n1.json 2> gpurun_out/r02k_bench_ref.err; echo "bench ref rc=$?"; cut -c1-300 gpurun_out/r02k_bench_reference_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:staple:: -c 300 --csv \
    --log-file gpurun_out/r02k_launches_bench_headline.csv python bench.py --sections headline --steps 4 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dslash_kernel -s 8 -c 1 -f -o gpurun_out/r02k_prof_dslash_64x3x128 \
    python bench.py --sections headline --steps 4 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_full_dslash.log 2>&1; echo "ncu dslash rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cgm_fused_kernel -s 30 -c 1 -f -o gpurun_out/r02k_prof_cgm_fused_48x3x96 \
    python bench.py --sections config3 --no-cpu-baseline --no-parity > gpurun_out/ncu_full_cgm.log 2>&1; echo "ncu cgm rc=$?"
ls -la gpurun_out/*.ncu-rep
