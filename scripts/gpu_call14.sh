#!/usr/bin/env bash
mkdir -p gpurun_out; rm -f gpurun_out/bench_configs.jsonl
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29561 --nproc-per-node 4"
timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -k "four" > gpurun_out/pytest_multi4.log 2>&1; echo "pytest four rc=$?"; tail -4 gpurun_out/pytest_multi4.log
timeout 200 python -m pytest tests/test_gpu_stout.py -m gpu -x -q > gpurun_out/pytest_stout.log 2>&1; echo "pytest stout rc=$?"; tail -3 gpurun_out/pytest_stout.log
timeout 150 $TR bench.py --gpus 4 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/final_bench_n4.json 2> gpurun_out/final_bench_n4.err; echo "bench n=4 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/final_bench_n4.json').read().strip().split('\n')[-1])
print(d['n_gpus'], 'ms/step', round(d['ms_per_step'],5), 'GF', round(d['value']), 'cgm ms/iter', round(d['multishift']['ms_per_iteration'],5), 'iters', d['multishift']['iterations'], 'mdagm', d['mdagm']['ms'])"
timeout 240 $TR scripts/bench_configs.py --global-lattice 64x64x64x16 --order 19 --skip-fp32 --reps 30 > /dev/null 2> gpurun_out/cfg5_n4.err; echo "cfg5 n=4 rc=$?"
timeout 300 python scripts/bench_configs.py --global-lattice 32x32x32x32 --order 19 --skip-fp32 --max-cg 60 > /dev/null 2> gpurun_out/cfg_32.err; tail -2 gpurun_out/cfg_32.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_configs.jsonl'):
    d=json.loads(l)
    print(d['global_lattice'], d['n_gpus'], 'pair ms', round(d['deo_doe_fp64']['ms_per_pair'],4), 'cgm ms/it', round(d['multishift_fp64']['ms_per_iteration'],4), 'force ms', round(d['fermion_force']['ms'],3), 'stout', d['stout_isotropic'], 'mdagm', d['mdagm_fp64']['ms'])
PY
