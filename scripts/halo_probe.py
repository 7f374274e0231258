#!/usr/bin/env python
"""Where does the time of acc_Deo/acc_Doe on D3 slabs go?  Per-launch CUDA-event timings of
   unsafe   the operator on the local interior, no exchange (the plain single-GPU kernel)
   eager    acc_Deo with its exchange, halos of `out` valid on return (the API's contract)
   mdagm    fermion_matrix_multiplication (API: both halos unpacked)
   cgm      CG-M iteration (staged halos inside when the transport allows), 19 shifts of the shipped approximation
for one or more LOCAL slab thicknesses, on N ranks (torchrun) or on one GPU in loopback (--loopback).
STAPLE_LIB selects a library variant (timing experiments: -DSTAPLE_DEBUG_NO_PEER_STORES, -DSTAPLE_DEBUG_NO_SIGNAL_FENCE)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--space", default="64x64x64")
    ap.add_argument("--loc3", default="2,8,16")
    ap.add_argument("--modes", default="1,4,3,2,0")
    ap.add_argument("--loopback", action="store_true")
    ap.add_argument("--reps", type=int, default=200)
    ap.add_argument("--no-cgm", action="store_true")
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import openstaple_b200 as osb
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    real_stdout = os.dup(1); os.dup2(2, 1)
    torch.cuda.set_stream(torch.cuda.Stream(device=torch.device("cuda", lr)))
    sp = tuple(int(x) for x in args.space.split("x"))
    rows = []
    for loc3 in [int(x) for x in args.loc3.split(",")]:
        for mode in [int(x) for x in args.modes.split(",")]:
            nr = world if world > 1 else 2
            lat = osb.Lattice(sp + (loc3,), nranks_d3=nr, device=lr)
            if world > 1:
                lat.init_multidev(dist, async_comm_fermion=1, p2p=mode)
            else:
                lat.init_loopback(mode)
            u, v = bench.make_fields(torch, lat, rank)
            ph = bench.staggered_phases(lat, rank, torch)
            a, b, c = v, lat.new_vec(), lat.new_vec()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

            def barrier():
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()

            def timeit(fn, reps):
                for _ in range(10):
                    fn()
                barrier(); e0.record()
                for _ in range(reps):
                    fn()
                e1.record(); barrier()
                ms = e0.elapsed_time(e1) / reps
                if world > 1:
                    t = torch.tensor([ms], device=lat.device, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
                return ms * 1e3
            row = {"space": args.space, "loc3": loc3, "mode": mode, "ranks": world, "loopback": world == 1, "tag": args.tag,
                   "unsafe_us": timeit(lambda: lat.acc_Deo_unsafe(u, b, a, ph), args.reps),
                   "eager_us": timeit(lambda: lat.acc_Deo(u, b, a, ph), args.reps)}
            row["unsafe_again_us"] = timeit(lambda: lat.acc_Deo_unsafe(u, b, a, ph), args.reps)      # clocks drift under the power cap
            pars = lat.ferm_param(0.0507, ph)
            row["mdagm_us"] = timeit(lambda: lat.fermion_matrix_multiplication(u, c, a, b, pars), args.reps // 2)
            if not args.no_cgm:
                approx = bench.shipped_order19(osb, 7.07, 0.0507)
                n = approx.approx_order
                sol, ps = lat.new_vec(n), lat.new_vec(n)
                r, h, s, p = (lat.new_vec() for _ in range(4))
                lat.multishift_invert(u, pars, approx, sol, v, 1e-8, r, h, s, p, ps, 24)
                barrier()
                lat.multishift_invert(u, pars, approx, sol, v, 1e-8, r, h, s, p, ps, 200)
                it, act, loop_ms = lat.last_solve_stats()
                row["cgm_us_per_iteration"] = loop_ms / max(it, 1) * 1e3
                row["cgm_mean_active"] = act / max(it, 1)
                del sol, ps
            rows.append(row)
            if rank == 0:
                sys.stderr.write(json.dumps(row) + "\n")
            lat.shutdown_multidev()
            del u, v, a, b, c, lat
            torch.cuda.empty_cache()
    if rank == 0:
        os.dup2(real_stdout, 1)
        for r in rows:
            print(json.dumps(r), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
