#!/usr/bin/env bash
# round 2, two GPUs: multi-rank parity tests (all transports), driver-contract bench at N=2, small-volume A/B of the halo protocols
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm --format=csv,noheader
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -k "two_gpus and (loc0 or staged or eager)" > gpurun_out/r02b_multirank_2gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r02b_multirank_2gpu.log; fi
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:2}"; }
run 29511 > gpurun_out/r02b_bench_n2.json 2> gpurun_out/r02b_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-2500 gpurun_out/r02b_bench_n2.json; tail -3 gpurun_out/r02b_bench_n2.err
for mode in 1 4 3 2 0; do
  STAPLE_P2P=$mode run $((29520+mode)) --lattice 64x64x64x4 --sections headline --steps 400 --warmup 20 --no-cpu-baseline > gpurun_out/r02b_small_p2p$mode.json 2> gpurun_out/r02b_small_p2p$mode.err; echo "small-volume p2p=$mode rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02b_small_p2p$mode.json').read().strip().splitlines()[-1])
    print('p2p=$mode ms_per_step', d['ms_per_step'], 'gflops', d['value'], 'cgm ms/it', d['multishift']['fp64']['ms_per_iteration'], 'parity', d['parity']['max_rel_err'], d['parity']['ok'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['unpipelined_ms_per_step'])
except Exception as e:
    print('p2p=$mode failed', e)
PY
done
ls gpurun_out/
