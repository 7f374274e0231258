#!/usr/bin/env bash
# 2 GPUs, local lattice 64^3 x 2 (the per-GPU shape of 64^3 x 16 on 8 GPUs: every site on a face): where does the time go?
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29551 --nproc-per-node 2"
show() { python -c "
import json,sys; d=json.loads(open('$1').read().strip().split('\n')[-1])
print('$2', 'ms/step', round(d['ms_per_step'],5), 'kernel alone us', round(d['roofline']['us_per_launch'],2), 'mdagm ms', round(d['mdagm']['ms'],5), 'cgm ms/it', round(d['multishift']['ms_per_iteration'],5) if d.get('multishift') else None)"; }
for mode in 1 3 2 0; do
  STAPLE_P2P=$mode timeout 200 $TR bench.py --gpus 2 --lattice 64x64x64x2 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/sv_p2p$mode.json 2> gpurun_out/sv_p2p$mode.err; show gpurun_out/sv_p2p$mode.json "p2p=$mode"
done
for dbg in 1 2 3; do
  STAPLE_DEBUG_HALO=$dbg STAPLE_P2P=1 timeout 200 $TR bench.py --gpus 2 --lattice 64x64x64x2 --steps 200 --warmup 10 --no-cpu-baseline --no-solver > gpurun_out/sv_dbg$dbg.json 2> gpurun_out/sv_dbg$dbg.err; show gpurun_out/sv_dbg$dbg.json "p2p=1 dbg=$dbg"
done
timeout 200 python bench.py --lattice 64x64x64x2 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/sv_n1.json 2> gpurun_out/sv_n1.err; show gpurun_out/sv_n1.json "single rank 64^3x2 periodic"
