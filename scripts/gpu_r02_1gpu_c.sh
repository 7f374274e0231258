#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02j_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02j_pytest_gpu.log
timeout 900 python bench.py --sections config1,config3 --no-cpu-baseline > gpurun_out/r02j_bench_c13.json 2> gpurun_out/r02j_bench_c13.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02j_bench_c13.json').read().strip().splitlines()[-1])
s=d['secondary']
print(json.dumps(s['config1_32x32x32x32']['single_system_solvers'],indent=1))
for k in ('fp64','fp32','fp32_accelerated_fp64_refined'): print(k, json.dumps(s['config3_48x48x48x96'][k]))
print(d['parity']['ok'], d['parity']['failures'])
PY
tail -3 gpurun_out/r02j_bench_c13.err
