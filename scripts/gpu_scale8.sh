#!/usr/bin/env bash
# 8-GPU box, most important first: driver-contract bench at 8, BASELINE configs 4 and 5 at 8, multi-rank parity at
# 4/8 ranks, then the smaller rank counts of the sweeps.
mkdir -p gpurun_out; rm -f gpurun_out/bench_configs.jsonl
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29541"
run_bench() { n=$1; if [ $n = 1 ]; then L="python"; else L="$TR --nproc-per-node $n"; fi
  timeout 150 $L bench.py --gpus $n --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/scale_bench_n$n.json 2> gpurun_out/scale_bench_n$n.err
  echo "bench n=$n rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/scale_bench_n$n.json').read().strip().split('\n')[-1])
print(d['n_gpus'], 'ms/step', round(d['ms_per_step'],5), 'GF', round(d['value']), 'cgm ms/iter', round(d['multishift']['ms_per_iteration'],5), 'iters', d['multishift']['iterations'], d['config'].get('halo'))"; }
run_cfg() { n=$1; lat=$2; reps=$3; if [ $n = 1 ]; then L="python"; else L="$TR --nproc-per-node $n"; fi
  timeout 240 $L scripts/bench_configs.py --global-lattice $lat --order 19 --skip-fp32 --reps $reps > /dev/null 2> gpurun_out/cfg_${lat}_n$n.err; echo "cfg $lat n=$n rc=$?"; }
run_bench 8
run_cfg 8 64x64x64x128 20
run_cfg 8 64x64x64x16 30
timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -k "four or eight" > gpurun_out/pytest_multi8.log 2>&1; echo "pytest_multi8 rc=$?"; tail -4 gpurun_out/pytest_multi8.log
run_cfg 1 64x64x64x128 10
run_bench 1; run_bench 2; run_bench 4
run_cfg 1 64x64x64x16 30; run_cfg 2 64x64x64x16 30; run_cfg 4 64x64x64x16 30
python - <<'PY'
import json
for l in open('gpurun_out/bench_configs.jsonl'):
    d=json.loads(l)
    print(d['global_lattice'], d['n_gpus'], d['halo'], 'pair ms', round(d['deo_doe_fp64']['ms_per_pair'],4), 'GF', round(d['deo_doe_fp64']['gflops']), 'cgm s', round(d['multishift_fp64']['s_per_solve'],4), 'ms/it', round(d['multishift_fp64']['ms_per_iteration'],4), 'it', d['multishift_fp64']['iterations'], 'force ms', round(d['fermion_force']['ms'],3))
PY
