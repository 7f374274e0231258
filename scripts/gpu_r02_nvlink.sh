#!/usr/bin/env bash
# two GPUs: NVLink bytes of the fused operator+halo kernel from the link counters; host topology probe; ncu of the segmented kernel in loopback
mkdir -p gpurun_out
echo "--- host topology"; ls /sys/devices/system/node/ 2>/dev/null | tr '\n' ' '; echo; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" ; cat /proc/self/status | grep -i "allowed_list"; nvidia-smi topo -m | head -5
for bdf in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader | head -2); do b=$(echo $bdf | tr 'A-Z' 'a-z' | sed 's/^0000//'); echo "$bdf numa_node: $(cat /sys/bus/pci/devices/$b/numa_node 2>/dev/null)"; done
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 scripts/nvlink_bytes.py "${@:2}"; }
run 29911 --lattice 64x64x64x16 --launches 2000 > gpurun_out/r02m_nvlink_bytes.jsonl 2> gpurun_out/r02m_nvlink.err; echo "nvlink 16 rc=$?"
run 29912 --lattice 64x64x64x2 --launches 4000 >> gpurun_out/r02m_nvlink_bytes.jsonl 2>> gpurun_out/r02m_nvlink.err; echo "nvlink 2 rc=$?"
cat gpurun_out/r02m_nvlink_bytes.jsonl; tail -3 gpurun_out/r02m_nvlink.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dslash_kernel -s 40 -c 1 -f -o gpurun_out/r02m_prof_dslash_segmented_loopback \
    python scripts/halo_probe.py --loopback --loc3 8 --modes 4 --no-cgm --reps 20 > gpurun_out/ncu_mr.log 2>&1; echo "ncu segmented rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -2
