#!/usr/bin/env bash
mkdir -p gpurun_out; rm -f gpurun_out/r02o_probe_loopback.jsonl
timeout 600 python -m pytest tests/test_gpu_loopback.py -m gpu -q -x > gpurun_out/r02o_loopback_tests.log 2>&1; echo "loopback tests rc=$?"; tail -3 gpurun_out/r02o_loopback_tests.log
timeout 600 python scripts/halo_probe.py --loopback --loc3 2,8,16 --modes 1 --tag loopback >> gpurun_out/r02o_probe_loopback.jsonl 2> gpurun_out/r02o_probe_loopback.err; echo "probe rc=$?"
STAPLE_LIB=$PWD/build/lib_faces6.so timeout 600 python scripts/halo_probe.py --loopback --loc3 2 --modes 1 --tag loopback-faces6 >> gpurun_out/r02o_probe_loopback.jsonl 2>> gpurun_out/r02o_probe_loopback.err; echo "probe faces6 rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r02o_probe_loopback.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print("%-16s ranks %d loc3 %2d mode %d  unsafe %7.1f/%7.1f  eager %7.1f (+%5.1f)  mdagm %7.1f (2x unsafe %7.1f)  cgm/it %7.1f" % (d['tag'], d['ranks'], d['loc3'], d['mode'], d['unsafe_us'], d.get('unsafe_again_us',0), d['eager_us'], d['eager_us']-d['unsafe_us'], d['mdagm_us'], 2*d['unsafe_us'], d.get('cgm_us_per_iteration', 0)))
PY
