#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02h_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02h_pytest_gpu.log
timeout 900 python bench.py --sections config1,config3 > gpurun_out/r02h_bench_c13.json 2> gpurun_out/r02h_bench_c13.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02h_bench_c13.json').read().strip().splitlines()[-1])
print(json.dumps(d['secondary'],indent=1)[:5000]); print(d['parity']['ok'], d['parity']['failures'])
PY
tail -3 gpurun_out/r02h_bench_c13.err
