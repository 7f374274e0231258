#!/usr/bin/env bash
mkdir -p gpurun_out; rm -f gpurun_out/bench_configs.jsonl
timeout 600 python -m pytest tests/test_gpu_stout.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_stout.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_stout.log
timeout 600 python scripts/bench_configs.py --global-lattice 32x32x32x32 --order 19 --skip-fp32 --max-cg 60 > /dev/null 2> gpurun_out/cfg_32.err; tail -2 gpurun_out/cfg_32.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_configs.jsonl').read().strip().split('\n')[-1]); print('stout', d['stout_isotropic'], 'mdagm', d['mdagm_fp64'], 'deo_doe', d['deo_doe_fp64']['ms_per_pair'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stout_staples_kernel -s 2 -c 1 -f -o gpurun_out/prof_stout_staples_kernel_v2 \
    python scripts/bench_configs.py --global-lattice 32x32x32x32 --order 19 --skip-fp32 --reps 3 --max-cg 40 > /dev/null 2> gpurun_out/ncu_full_stout2.log; echo "ncu rc=$?"
