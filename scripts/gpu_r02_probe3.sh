#!/usr/bin/env bash
mkdir -p gpurun_out; rm -f gpurun_out/r02u_probe.jsonl
timeout 300 python -m pytest tests/test_gpu_loopback.py -m gpu -q -x > gpurun_out/r02u_loopback_tests.log 2>&1; echo "loopback tests rc=$?"; tail -3 gpurun_out/r02u_loopback_tests.log
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x -k "two_gpus and loc0 and (staged or eager)" > gpurun_out/r02u_multirank_2gpu.log 2>&1; echo "multirank tests rc=$?"; tail -3 gpurun_out/r02u_multirank_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scripts/halo_probe.py --loc3 8,16,64 --modes 1 --no-cgm >> gpurun_out/r02u_probe.jsonl 2> gpurun_out/r02u_probe.err; echo "probe rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r02u_probe.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print("ranks %d loc3 %2d mode %d  unsafe %7.1f/%7.1f  eager %7.1f (+%5.1f)  mdagm %7.1f (2x unsafe %7.1f)" % (d['ranks'], d['loc3'], d['mode'], d['unsafe_us'], d.get('unsafe_again_us',0), d['eager_us'], d['eager_us']-d['unsafe_us'], d['mdagm_us'], 2*d['unsafe_us']))
PY
