#!/usr/bin/env bash
# 1-GPU: full parity suite (incl. the streamed host round trip), bench (driver contract), chunk sweep of the
# pipelined host round trip, CG-M shifted-pass tuning sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
for c in 1 2 4 8; do
  timeout 300 python bench.py --no-solver --no-cpu-baseline --steps 60 --stream-chunk $c 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); e=d['e2e']; print('chunk $c e2e ms', round(e['ms_per_step'],4), 'GF', round(e['value']), 'plain ms', round(e['unpipelined_ms_per_step'],4))"
done
timeout 600 python scripts/tune_cgm.py run 32x32x32x32 200
timeout 900 python scripts/tune_cgm.py run 48x48x48x96 50
