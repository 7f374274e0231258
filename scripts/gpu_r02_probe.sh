#!/usr/bin/env bash
# two GPUs: loopback + multi-rank tests, halo probe over NVLink for the transports and timing-only library variants, loopback probe
mkdir -p gpurun_out; rm -f gpurun_out/r02e_probe.jsonl
timeout 600 python -m pytest tests/test_gpu_loopback.py -m gpu -q -x > gpurun_out/r02e_loopback_tests.log 2>&1; echo "loopback tests rc=$?"; tail -5 gpurun_out/r02e_loopback_tests.log
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x -k "two_gpus and (staged or eager or three or unpack)" > gpurun_out/r02e_multirank_2gpu.log 2>&1; echo "multirank tests rc=$?"; tail -5 gpurun_out/r02e_multirank_2gpu.log
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 scripts/halo_probe.py "${@:2}"; }
run 29611 --loc3 2,8,16,64 --modes 1,4,2 >> gpurun_out/r02e_probe.jsonl 2> gpurun_out/r02e_probe.err; echo "probe rc=$?"
STAPLE_LIB=$PWD/build/lib_mr7.so run 29612 --loc3 2,8,16 --modes 1 --tag mr7 >> gpurun_out/r02e_probe.jsonl 2>> gpurun_out/r02e_probe.err; echo "probe mr7 rc=$?"
STAPLE_LIB=$PWD/build/lib_nopeer.so run 29613 --loc3 2,8 --modes 4 --tag nopeer >> gpurun_out/r02e_probe.jsonl 2>> gpurun_out/r02e_probe.err; echo "probe nopeer rc=$?"
timeout 600 python scripts/halo_probe.py --loopback --loc3 2,8,16 --modes 1,4 --tag loopback >> gpurun_out/r02e_probe.jsonl 2>> gpurun_out/r02e_probe.err; echo "probe loopback rc=$?"
STAPLE_LIB=$PWD/build/lib_mr7.so timeout 600 python scripts/halo_probe.py --loopback --loc3 2,8,16 --modes 1 --tag loopback-mr7 >> gpurun_out/r02e_probe.jsonl 2>> gpurun_out/r02e_probe.err; echo "probe loopback mr7 rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r02e_probe.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print("%-12s ranks %d loc3 %2d mode %d  unsafe %7.1f/%7.1f  eager %7.1f (+%5.1f)  mdagm %7.1f (2x unsafe %7.1f)  cgm/it %7.1f" % (d['tag'], d['ranks'], d['loc3'], d['mode'], d['unsafe_us'], d.get('unsafe_again_us',0), d['eager_us'], d['eager_us']-d['unsafe_us'], d['mdagm_us'], 2*d['unsafe_us'], d.get('cgm_us_per_iteration', 0)))
PY
grep -v "^WARNING\|^{" gpurun_out/r02e_probe.err | tail -5
