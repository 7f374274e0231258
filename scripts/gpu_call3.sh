#!/usr/bin/env bash
# 1-GPU: parity suite, bench (driver contract), chunk sweep of the graph-replayed host round trip, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
for c in 1 2 4; do
  timeout 300 python bench.py --no-solver --no-cpu-baseline --steps 60 --stream-chunk $c 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); e=d['e2e']; print('chunk $c e2e ms', round(e['ms_per_step'],4), 'GF', round(e['value']), 'plain ms', round(e['unpipelined_ms_per_step'],4))"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:staple:: -s 40 -c 500 --csv \
    --log-file gpurun_out/launches_solver.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
rm -f gpurun_out/bench_configs.jsonl
timeout 900 python scripts/bench_configs.py --global-lattice 32x32x32x32 --order 19 --mass 0.0507 --skip-fp32 > /dev/null 2> gpurun_out/cfg_32.err; tail -2 gpurun_out/cfg_32.err
timeout 1200 python scripts/bench_configs.py --global-lattice 48x48x48x96 --order 19 --mass 0.0507 --skip-fp32 > /dev/null 2> gpurun_out/cfg_48.err; tail -2 gpurun_out/cfg_48.err
cat gpurun_out/bench_configs.jsonl
