#!/usr/bin/env bash
# Runs on an N-GPU B200 box (gpurun --gpus N): single-GPU parity, multi-GPU parity, 1..N bench sweep.
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_multi.txt 2>&1
nvidia-smi topo -m >> gpurun_out/smi_multi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest1 rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 1500 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest_multi rc=$?"; tail -25 gpurun_out/pytest_multi.log
for n in 1 $N; do
  for p2p in 0 1; do
    if [ $n = 1 ] && [ $p2p = 1 ]; then continue; fi
    if [ $n = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531"; fi
    STAPLE_P2P=$p2p timeout 600 $L bench.py --gpus $n --steps 100 --warmup 5 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/bench_n${n}_p2p${p2p}.json 2> gpurun_out/bench_n${n}_p2p${p2p}.err
    echo "bench n=$n p2p=$p2p rc=$?"; cut -c1-600 gpurun_out/bench_n${n}_p2p${p2p}.json; tail -3 gpurun_out/bench_n${n}_p2p${p2p}.err
  done
done
