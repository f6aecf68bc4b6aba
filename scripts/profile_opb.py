"""ncu target: upload one BASELINE workload and apply opb a few times (3 warm-up + reps timed applications).
   python scripts/profile_opb.py [--workload cfg2] [--scale 1.0] [--reps 3] [--variant -1]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svdlibrs_b200 import synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--variant", type=int, default=-1)
ap.add_argument("--flush", type=int, default=0)
ap.add_argument("--hot", type=int, default=-1)
ap.add_argument("--shard", type=int, default=1, help="profile the first of this many row shards (multi-GPU shard shape)")
a = ap.parse_args()
try:
    cores = len(os.sched_getaffinity(0))
except Exception:
    cores = os.cpu_count() or 1
A, dims = synthetic.make(a.workload, scale=a.scale, workers=min(32, cores))  # forked generator workers: before CUDA starts
if a.shard > 1:
    A = A[: A.shape[0] // a.shard].tocsr()
import svdlibrs_b200 as las2
with las2.Operator(A) as op:
    op.set("spmv_variant", a.variant)
    op.set("spmv_hot", a.hot)
    t = op.time_opb(reps=a.reps, flush_l2=bool(a.flush))
    t["GBs_opb"] = t["bytes_opb"] / t["ms_per_opb"] / 1e6
    t["GBs_b"] = t["bytes_spmv_b"] / t["ms_spmv_b"] / 1e6
    t["GBs_bt"] = t["bytes_spmv_bt"] / t["ms_spmv_bt"] / 1e6
    t.update(workload=a.workload, scale=a.scale, shape=list(A.shape), nnz=int(A.nnz), m=op.m, n=op.n, variant=a.variant)
    print(json.dumps(t))
