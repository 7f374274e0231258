#!/usr/bin/env bash
mkdir -p gpurun_out; rm -f gpurun_out/r02n_probe.jsonl
timeout 600 python -m pytest tests/test_gpu_loopback.py -m gpu -q -x > gpurun_out/r02n_loopback_tests.log 2>&1; echo "loopback tests rc=$?"; tail -4 gpurun_out/r02n_loopback_tests.log
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x -k "two_gpus and (staged or eager or three or unpack)" > gpurun_out/r02n_multirank_2gpu.log 2>&1; echo "multirank tests rc=$?"; tail -4 gpurun_out/r02n_multirank_2gpu.log
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 scripts/halo_probe.py "${@:2}"; }
run 29611 --loc3 2,8,16 --modes 1,4 >> gpurun_out/r02n_probe.jsonl 2> gpurun_out/r02n_probe.err; echo "probe rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r02n_probe.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print("%-12s ranks %d loc3 %2d mode %d  unsafe %7.1f/%7.1f  eager %7.1f (+%5.1f)  mdagm %7.1f (2x unsafe %7.1f)  cgm/it %7.1f" % (d['tag'], d['ranks'], d['loc3'], d['mode'], d['unsafe_us'], d.get('unsafe_again_us',0), d['eager_us'], d['eager_us']-d['unsafe_us'], d['mdagm_us'], 2*d['unsafe_us'], d.get('cgm_us_per_iteration', 0)))
PY
grep -v "^WARNING\|^{" gpurun_out/r02n_probe.err | tail -3
nvidia-smi nvlink -gt d -i 0 2>&1 | head -8
