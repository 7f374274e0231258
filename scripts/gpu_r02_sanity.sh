#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --sections headline,config1 --steps 20 --warmup 3 --cpu-seconds 3 > gpurun_out/r02v_bench_sanity.json 2> gpurun_out/r02v_bench_sanity.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02v_bench_sanity.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline'], d['parity']['ok'], list(d['secondary']))
PY
