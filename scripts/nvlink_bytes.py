#!/usr/bin/env python
"""NVLink bytes of the fused operator + halo kernel, from the driver's own link counters (NVML field values
NVLINK_THROUGHPUT_DATA_TX / _RX, KiB, summed over links): K launches of acc_Deo on D3 slabs between two readings, against the
algorithmic 2 faces x 3 colours x vol3h x 16 B per launch and direction.  torchrun, 2+ ranks."""
import argparse
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def nvlink_kib(index):
    import pynvml as N
    N.nvmlInit()
    h = N.nvmlDeviceGetHandleByIndex(index)
    out = {}
    for name, fid in (("tx", N.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX), ("rx", N.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX)):
        try:
            v = N.nvmlDeviceGetFieldValues(h, [(fid, 0xFFFFFFFF)])[0]          # scopeId UINT_MAX: sum over all links
        except Exception:
            v = N.nvmlDeviceGetFieldValues(h, [fid])[0]
        out[name] = int(v.value.ullVal)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lattice", default="64x64x64x16", help="LOCAL lattice per GPU")
    ap.add_argument("--launches", type=int, default=2000)
    ap.add_argument("--mode", type=int, default=1)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import openstaple_b200 as osb
    world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    real_stdout = os.dup(1); os.dup2(2, 1)
    torch.cuda.set_stream(torch.cuda.Stream(device=torch.device("cuda", lr)))
    loc = tuple(int(x) for x in args.lattice.split("x"))
    lat = osb.Lattice(loc, nranks_d3=world, device=lr)
    lat.init_multidev(dist, async_comm_fermion=1, p2p=args.mode)
    u, v = bench.make_fields(torch, lat, rank)
    ph = bench.staggered_phases(lat, rank, torch)
    b = lat.new_vec()
    for _ in range(20):
        lat.acc_Deo(u, b, v, ph)
    dist.barrier(); torch.cuda.synchronize()
    c0 = nvlink_kib(lr)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.launches):
        lat.acc_Deo(u, b, v, ph)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    c1 = nvlink_kib(lr)
    algorithmic = 2 * 3 * lat.vol3h * 16                         # bytes per launch and direction (both faces)
    row = {"rank": rank, "ranks": world, "local_lattice": args.lattice, "mode": args.mode, "launches": args.launches,
           "us_per_launch": e0.elapsed_time(e1) / args.launches * 1e3,
           "nvlink_tx_bytes_per_launch": (c1["tx"] - c0["tx"]) * 1024 / args.launches,
           "nvlink_rx_bytes_per_launch": (c1["rx"] - c0["rx"]) * 1024 / args.launches,
           "algorithmic_bytes_per_launch_and_direction": algorithmic}
    row["tx_over_algorithmic"] = row["nvlink_tx_bytes_per_launch"] / algorithmic
    row["rx_over_algorithmic"] = row["nvlink_rx_bytes_per_launch"] / algorithmic
    rows = [None] * world
    dist.all_gather_object(rows, row)
    lat.shutdown_multidev()
    if rank == 0:
        os.dup2(real_stdout, 1)
        for r in rows:
            print(json.dumps(r), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
