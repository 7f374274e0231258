#!/usr/bin/env bash
# final 1-GPU validation of the round: parity suite, driver-contract bench (both arms), launch list, full ncu
# captures of the dominant kernels, BASELINE config benchmarks incl. force and stout
mkdir -p gpurun_out; rm -f gpurun_out/bench_configs.jsonl
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cut -c1-600 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:staple:: -s 40 -c 600 --csv \
    --log-file gpurun_out/launches_bench.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dslash_kernel -s 12 -c 1 -f -o gpurun_out/prof_dslash \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-solver > gpurun_out/ncu_full_dslash.log 2>&1; echo "ncu dslash rc=$?"
for kern in cgm_fused_kernel force_outer_kernel stout_staples_kernel stout_exp_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$kern -s 2 -c 1 -f -o gpurun_out/prof_$kern \
    python scripts/bench_configs.py --global-lattice 32x32x32x32 --order 19 --skip-fp32 --reps 3 --max-cg 40 > /dev/null 2> gpurun_out/ncu_full_$kern.log; echo "ncu $kern rc=$?"
done
rm -f gpurun_out/bench_configs.jsonl
timeout 600 python scripts/bench_configs.py --global-lattice 32x32x32x32 --order 19 --mass 0.0507 > /dev/null 2> gpurun_out/cfg_32.err; tail -2 gpurun_out/cfg_32.err
timeout 900 python scripts/bench_configs.py --global-lattice 48x48x48x96 --order 19 --mass 0.0507 > /dev/null 2> gpurun_out/cfg_48.err; tail -2 gpurun_out/cfg_48.err
cat gpurun_out/bench_configs.jsonl
ls -la gpurun_out/*.ncu-rep
