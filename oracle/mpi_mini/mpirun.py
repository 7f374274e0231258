#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- launcher for oracle/mpi_mini: `mpirun.py -n N [--cwd DIR] prog args...` starts N copies of prog
with MINI_MPI_RANK / MINI_MPI_SIZE / MINI_MPI_DIR set, prefixes nothing, returns the largest exit code.  Rank r's stdout goes to
<cwd>/stdout.r (rank 0's also to this process' stdout)."""
import os
import subprocess
import sys
import tempfile


def launch(n, argv, cwd=None, env=None, timeout=3600):
    cwd = cwd or os.getcwd()
    with tempfile.TemporaryDirectory(prefix="mpimini") as d:
        procs = []
        for r in range(n):
            e = dict(env or os.environ); e.update(MINI_MPI_RANK=str(r), MINI_MPI_SIZE=str(n), MINI_MPI_DIR=d)
            procs.append(subprocess.Popen(argv, cwd=cwd, env=e, stdout=open(os.path.join(cwd, "stdout.%d" % r), "w"),
                                          stderr=open(os.path.join(cwd, "stderr.%d" % r), "w")))
        rc = 0
        for p in procs:
            try:
                rc = max(rc, abs(p.wait(timeout=timeout)))
            except subprocess.TimeoutExpired:
                for q in procs:
                    q.kill()
                rc = max(rc, 124)
        return rc


if __name__ == "__main__":
    a = sys.argv[1:]
    n = 1
    if a[:1] == ["-n"]:
        n = int(a[1]); a = a[2:]
    rc = launch(n, a)
    sys.stdout.write(open("stdout.0").read())
    sys.exit(rc)
