#!/usr/bin/env bash
# 8-GPU box: multi-rank parity at 4 and 8 ranks, bench.py scaling sweep, configs 4 and 5.
mkdir -p gpurun_out; rm -f gpurun_out/bench_configs.jsonl
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29541"
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -k "four or eight" > gpurun_out/pytest_multi8.log 2>&1; echo "pytest_multi8 rc=$?"; tail -15 gpurun_out/pytest_multi8.log
for n in 1 2 4 8; do
  if [ $n = 1 ]; then L="python"; else L="$TR --nproc-per-node $n"; fi
  timeout 600 $L bench.py --gpus $n --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/scale_bench_n$n.json 2> gpurun_out/scale_bench_n$n.err
  echo "bench n=$n rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/scale_bench_n$n.json').read().strip().split('\n')[-1])
print(d['n_gpus'], 'ms/step', d['ms_per_step'], 'GF', d['value'], 'cgm ms/iter', d['multishift']['ms_per_iteration'], 'iters', d['multishift']['iterations'], d['config'].get('halo'))"
done
STAPLE_P2P=0 timeout 600 $TR --nproc-per-node 8 bench.py --gpus 8 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/scale_bench_n8_nccl.json 2> gpurun_out/scale_bench_n8_nccl.err; echo "bench n=8 nccl rc=$?"
# config 5: 64^3 x 16 strong scaling; config 4: 64^3 x 128 on 8 (and on 1 for the parallel efficiency)
for n in 1 2 4 8; do
  if [ $n = 1 ]; then L="python"; else L="$TR --nproc-per-node $n"; fi
  timeout 600 $L scripts/bench_configs.py --global-lattice 64x64x64x16 --order 19 --skip-fp32 --reps 30 > /dev/null 2> gpurun_out/cfg5_n$n.err; echo "cfg5 n=$n rc=$?"
done
timeout 900 $TR --nproc-per-node 8 scripts/bench_configs.py --global-lattice 64x64x64x128 --order 19 --skip-fp32 --reps 20 > /dev/null 2> gpurun_out/cfg4_n8.err; echo "cfg4 n=8 rc=$?"
timeout 900 python scripts/bench_configs.py --global-lattice 64x64x64x128 --order 19 --skip-fp32 --reps 10 > /dev/null 2> gpurun_out/cfg4_n1.err; echo "cfg4 n=1 rc=$?"
STAPLE_P2P=0 timeout 600 $TR --nproc-per-node 8 scripts/bench_configs.py --global-lattice 64x64x64x16 --order 19 --skip-fp32 --reps 30 > /dev/null 2> gpurun_out/cfg5_n8_nccl.err; echo "cfg5 n=8 nccl rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_configs.jsonl'):
    d=json.loads(l)
    print(d['global_lattice'], d['n_gpus'], d['halo'], 'pair ms', round(d['deo_doe_fp64']['ms_per_pair'],4), 'GF', round(d['deo_doe_fp64']['gflops']), 'cgm s', round(d['multishift_fp64']['s_per_solve'],4), 'ms/it', round(d['multishift_fp64']['ms_per_iteration'],4), 'it', d['multishift_fp64']['iterations'])
PY
