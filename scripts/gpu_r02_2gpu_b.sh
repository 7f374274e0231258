#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multirank.py -m gpu -q -k "two_gpus" > gpurun_out/r02i_multirank_2gpu.log 2>&1; echo "multirank rc=$?"; tail -5 gpurun_out/r02i_multirank_2gpu.log
timeout 900 python -m pytest tests/test_gpu_zz_reference_host_multirank.py tests/test_gpu_zz_reference_rhmc.py -m gpu -q -rxXs > gpurun_out/r02i_hostprograms_2gpu.log 2>&1; echo "host programs rc=$?"; tail -8 gpurun_out/r02i_hostprograms_2gpu.log
timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py -m gpu -q > gpurun_out/r02i_baseline_shapes.log 2>&1; echo "baseline shapes rc=$?"; tail -8 gpurun_out/r02i_baseline_shapes.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 > gpurun_out/r02i_bench_n2.json 2> gpurun_out/r02i_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02i_bench_n2.json').read().strip().splitlines()[-1])
print('ms_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['unpipelined_ms_per_step'])
print('cgm', d['multishift']['fp64'])
for k,v in d['secondary'].items(): print(k, json.dumps(v)[:900])
print(d['parity']['ok'], d['parity']['failures'], d['parity']['max_rel_err'], d['parity']['cg_iters'])
PY
