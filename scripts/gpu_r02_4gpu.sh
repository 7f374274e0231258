#!/usr/bin/env bash
# round 2, four GPUs: driver-contract bench at N = 4 with the final code
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 4 > gpurun_out/r02t_bench_n4.json 2> gpurun_out/r02t_bench_n4.err; echo "bench n4 rc=$?"; cut -c1-300 gpurun_out/r02t_bench_n4.json; grep -i "parity" gpurun_out/r02t_bench_n4.err | head -3
