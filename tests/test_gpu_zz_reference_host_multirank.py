"""Multi-GPU drop-in, executed (needs 2 GPUs; skipped otherwise): the reference's deo_doe_test built for NRANKS_D3 = 2 and
linked against libstaple_b200.so runs as two processes, one per GPU, under oracle/mpi_mini -- the reference's own MPI code
scatters the configuration and gathers the results through rank 0, openstaple_b200/host/memory_wrapper_staple.c hands the NCCL id around with MPI_Bcast
and calls staple_init_multidev1D, the operator and its halo exchange run in the library (NCCL over
NVLink).  The global result files are compared with the SINGLE-rank pure-reference build reading the configuration and source
the two GPU ranks saved: FP64 relative 1e-13.

Written at the end of round 1 after the GPU budget was spent (the CPU half, tests/test_reference_host_multirank_cpu.py, runs the
pure-reference two-rank program through the same launcher).  First run on 2 x B200 in round 2: green (profiles/r02i_hostprograms_2gpu.log)."""
import os

import numpy as np
import pytest

from test_reference_host_multirank_cpu import FILES, run_single_rank_on_saved_inputs, run_two_ranks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["staple", "staplemd"])
def test_two_gpu_reference_deo_doe_program(tmp_path, kind):
    """kind staplemd: src/Mpi/multidev.c is left out as well and openstaple_b200/host/multidev_staple.c provides devinfo,
    pre_init_multidev1D, init_multidev1D (which joins the library's rank layer) and shutdown_multidev"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = ":".join(x for x in (env.get("LD_LIBRARY_PATH", ""), "/usr/local/cuda/lib64") if x)
    td = str(tmp_path)
    two = run_two_ranks(kind, td, env=env)
    err0 = open(os.path.join(td, "stderr.0")).read()
    assert "hot path served by staple_b200" in err0
    assert ("joined in init_multidev1D" in err0) == (kind == "staplemd")
    one = run_single_rank_on_saved_inputs(td)
    for f in FILES:
        e = float(np.abs(two[f] - one[f]).max() / np.abs(one[f]).max())
        print("%s: %.1e" % (f, e))
        assert e < 1e-13, (f, e)
