#!/usr/bin/env bash
# round 2, eight GPUs, final: 4- and 8-rank parity tests (8 ranks: LOC_N3 = 2, every site on a face), the reference's two-rank
# programs against the library, and the driver-contract bench at N = 8 (N = 4 with WITH_N4=1)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x -k "eight_gpus or four_gpus" > gpurun_out/r02p_multirank_8gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02p_multirank_8gpu.log
timeout 600 python -m pytest tests/test_gpu_zz_reference_host_multirank.py tests/test_gpu_zz_reference_rhmc.py -m gpu -q > gpurun_out/r02p_hostprograms.log 2>&1; echo "host programs rc=$?"; tail -3 gpurun_out/r02p_hostprograms.log
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 "${@:3}"; }
run 8 29711 > gpurun_out/r02p_bench_n8.json 2> gpurun_out/r02p_bench_n8.err; echo "bench n8 rc=$?"; cut -c1-400 gpurun_out/r02p_bench_n8.json; grep -i "parity\|error" gpurun_out/r02p_bench_n8.err | head -5
if [ -n "$WITH_N4" ]; then run 4 29712 > gpurun_out/r02p_bench_n4.json 2> gpurun_out/r02p_bench_n4.err; echo "bench n4 rc=$?"; cut -c1-400 gpurun_out/r02p_bench_n4.json; fi
