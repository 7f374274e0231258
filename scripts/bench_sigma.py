#!/usr/bin/env python
"""Sigma' -> Sigma chain of the stouted force on one GPU (stouting.c:171-1305): compute_lambda, compute_sigma and the whole
compute_sigma_from_sigma_prime_backinto_sigma_prime, per level.  STAPLE_LIB selects a library variant."""
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import openstaple_b200 as osb
    loc = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "32x32x32x32").split("x"))
    torch.cuda.set_device(0)
    torch.cuda.set_stream(torch.cuda.Stream(device=torch.device("cuda", 0)))
    lat = osb.Lattice(loc, device=0)
    u, _ = bench.make_fields(torch, lat, 0)
    g = torch.Generator(device=lat.device); g.manual_seed(5)
    sp = torch.complex(torch.randn(u.shape, generator=g, device=lat.device, dtype=torch.float64), torch.randn(u.shape, generator=g, device=lat.device, dtype=torch.float64))
    lat.set_stout(0.15, 1)
    lam, qa, tmp = lat.new_tamat(), lat.new_tamat(), lat.new_conf()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timeit(fn, reps=10):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    sg = sp.clone()
    lat.compute_sigma_from_sigma_prime_backinto_sigma_prime(sg, lam, qa, u, tmp, 0)       # leaves Q, Lambda consistent
    out = {"lattice": "x".join(map(str, loc)), "lib": os.environ.get("STAPLE_LIB", "default")}
    out["compute_lambda_ms"] = timeit(lambda: lat.compute_lambda(lam, sp, u, qa, tmp))
    out["compute_sigma_ms"] = timeit(lambda: lat.compute_sigma(lam, u, sg, qa, tmp, 0))
    out["whole_chain_ms"] = timeit(lambda: lat.compute_sigma_from_sigma_prime_backinto_sigma_prime(sg, lam, qa, u, tmp, 0))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
