"""Worker of tests/test_sharding_gloo.py: one rank of a world_size-N gloo group on CPU.  Exercises the host side of
the multi-GPU path (row partition by nnz, shard extraction, id broadcast, sum of partial opb products ==
the full product, M-side vector slices concatenating to the full vector).  The device kernels and NCCL are
covered on the GPU box (tests/test_gpu_multi.py)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import svdlibrs_b200  # noqa: E402
from svdlibrs_b200 import sharding, synthetic  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = {}
    for name, fmt in (("tall_csr", "csr"), ("wide_csc", "csc")):
        A, _ = synthetic.make("cfg2" if fmt == "csr" else "cfg3", scale=0.004)
        B, transposed = sharding.orient(A)
        assert transposed == (fmt == "csc")
        bounds = sharding.partition_rows_by_nnz(B.indptr, world)
        Bl, lo, hi = sharding.local_shard(B, bounds, rank)
        # id broadcast plumbing (a stand-in id: NCCL itself is not exercised on the CPU box)
        svdlibrs_b200.comm_unique_id = lambda: bytes(range(128))
        sharding.comm_unique_id = svdlibrs_b200.comm_unique_id
        cid = sharding.broadcast_comm_id(dist, rank)
        assert cid == bytes(range(128))
        # y = sum_g B_g'(B_g x): the one exchange step of the path
        x = np.random.default_rng(5).standard_normal(B.shape[1])
        t_local = Bl @ x
        y = torch.from_numpy(Bl.T @ t_local)
        dist.all_reduce(y)
        y_full = B.T @ (B @ x)
        err = float(np.linalg.norm(y.numpy() - y_full) / np.linalg.norm(y_full))
        # M-side slices concatenate without a collective
        parts = [None] * world
        dist.all_gather_object(parts, (lo, hi, t_local))
        t_cat = np.concatenate([p[2] for p in sorted(parts, key=lambda p: p[0])])
        nnz = [None] * world
        dist.all_gather_object(nnz, int(Bl.nnz))
        out[name] = dict(err=err, cat_ok=bool(np.allclose(t_cat, B @ x)), bounds=[int(b) for b in bounds], nnz=nnz,
                         total=int(B.nnz), lo=lo, hi=hi)
    if rank == 0:
        print("RESULT " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
