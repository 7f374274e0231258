#!/usr/bin/env bash
# 1-GPU: parity, bench (driver contract), launch list of bench incl. solver, config-3 style solver benchmarks
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:staple:: -s 40 -c 400 --csv \
    --log-file gpurun_out/launches_solver.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
rm -f gpurun_out/bench_configs.jsonl
timeout 900 python scripts/bench_configs.py --global-lattice 32x32x32x32 --order 19 --mass 0.0507 > /dev/null 2> gpurun_out/cfg_32.err; tail -2 gpurun_out/cfg_32.err
timeout 1200 python scripts/bench_configs.py --global-lattice 48x48x48x96 --order 19 --mass 0.0507 > /dev/null 2> gpurun_out/cfg_48.err; tail -2 gpurun_out/cfg_48.err
timeout 1200 python scripts/bench_configs.py --global-lattice 48x48x48x96 --order 19 --mass 0.0018 --skip-fp32 > /dev/null 2> gpurun_out/cfg_48l.err; tail -2 gpurun_out/cfg_48l.err
cat gpurun_out/bench_configs.jsonl
