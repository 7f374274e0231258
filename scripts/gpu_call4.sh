#!/usr/bin/env bash
mkdir -p gpurun_out
python scripts/pcie_probe.py
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "streamed or deo_doe" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for m in 1 0; do for c in 0 2 4; do
  timeout 300 python bench.py --no-solver --no-cpu-baseline --steps 60 --stream-chunk $c --stream-mode $m 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); e=d['e2e']; print('mode $m chunk $c e2e ms', round(e['ms_per_step'],4), 'GF', round(e['value']), 'plain ms', round(e['unpipelined_ms_per_step'],4), 'value', round(d['value']))"
done; done
