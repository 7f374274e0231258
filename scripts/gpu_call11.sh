#!/usr/bin/env bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29551 --nproc-per-node 2"
show() { python -c "
import json,sys; d=json.loads(open('$1').read().strip().split('\n')[-1])
print('$2', 'ms/step', round(d['ms_per_step'],5), 'kernel alone us', round(d['roofline']['us_per_launch'],2), 'mdagm ms', round(d['mdagm']['ms'],5), 'cgm ms/it', round(d['multishift']['ms_per_iteration'],5) if d.get('multishift') else None)"; }
timeout 600 python -m pytest tests/test_gpu_stout.py tests/test_gpu_force.py -m gpu -x -q > gpurun_out/pytest_stout.log 2>&1; echo "pytest stout+force rc=$?"; tail -12 gpurun_out/pytest_stout.log
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -8 gpurun_out/pytest_multi.log
STAPLE_P2P=1 timeout 200 $TR bench.py --gpus 2 --lattice 64x64x64x2 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/sv3_p2p1.json 2> gpurun_out/sv3_p2p1.err; show gpurun_out/sv3_p2p1.json "64^3x2 p2p=1"
STAPLE_P2P=1 timeout 200 $TR bench.py --gpus 2 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/sv3_weak.json 2> gpurun_out/sv3_weak.err; show gpurun_out/sv3_weak.json "weak 32^4 p2p=1"
